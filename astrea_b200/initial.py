"""Initial conditions: the test-problem table and the grid initialiser (host side, runs once per run).

Follows static/tests.py:7-328 (problem table) and functions/constructor.py:11-109 (``initialise``):
pointwise primitive initial data on cell centres -> conservative variables -> cell averages via
``+ laplacian/24`` (functions/fv.py:67-85).  This is input preparation, not the per-step hot path
(SURVEY.md §8f rank 2), so it stays in numpy on the host; the arrays it returns are what
``astrea_upload`` sends to the device.  ``extent`` generalises the reference's square ``N x N`` grid to
an ``Nx x Ny`` window of a larger periodic box for the slab-decomposed weak-scaling runs.
"""
import numpy as np

_A = np.array


def _ll(index):
    """Lax & Liu (1998) quadrant states, static/tests.py:217-315.  ``coeff = -1**index`` is -1 for 5 and 6 (Q8)."""
    t = {
        1: ([.5197, -.7259, 0, 0, .4], [1, 0, 0, 0, 1], [.1072, -.7259, -1.4045, 0, .0439], [.2579, 0, -1.4045, 0, .15]),
        2: ([.5197, -.7259, 0, 0, .4], [1, 0, 0, 0, 1], [1, -.7259, -.7259, 0, 1], [.5197, 0, -.7259, 0, .4]),
        3: ([1.5, 0, 0, 0, 1.5], [.5323, 1.206, 0, 0, .3], [.5323, 0, 1.206, 0, .3], [.138, 1.206, 1.206, 0, .029]),
        4: ([.5065, .8939, 0, 0, .35], [1.1, 0, 0, 0, 1.1], [1.1, .8939, .8939, 0, 1.1], [.5065, 0, .8939, 0, .35]),
        5: ([2, -.75, .5, 0, 1], [1, -.75, -.5, 0, 1], [1, .75, .5, 0, 1], [3, .75, -.5, 0, 1]),
        7: ([.5197, -.6259, .1, 0, .4], [1, .1, .1, 0, 1], [.8, .1, .1, 0, .4], [.5197, .1, -.6259, 0, .4]),
        8: ([1, -.6259, .1, 0, 1], [.5197, .1, .1, 0, .4], [.8, .1, .1, 0, 1], [1, .1, -.6259, 0, 1]),
        9: ([2, 0, -.3, 0, 1], [1, 0, .3, 0, 1], [1.039, 0, -.8133, 0, .4], [.5197, 0, -.4259, 0, .4]),
        10: ([.5, 0, .6076, 0, 1], [1, 0, .4297, 0, 1], [.2281, 0, -.6076, 0, .3333], [.4562, 0, -.4297, 0, .3333]),
        11: ([.5313, .8276, 0, 0, .4], [1, .1, 0, 0, 1], [.8, .1, 0, 0, .4], [.5313, .1, .7276, 0, .4]),
        12: ([1, .7276, 0, 0, 1], [.5313, 0, 0, 0, .4], [.8, 0, 0, 0, 1], [1, 0, .7276, 0, 1]),
        13: ([2, .3, 0, 0, 1], [1, 0, -.3, 0, 1], [1.0625, 0, .8145, 0, .4], [.5313, 0, .4276, 0, .4]),
        14: ([1, 0, -1.2172, 0, 8], [2, 0, -.5606, 0, 8], [.4736, 0, 1.2172, 0, 2.6667], [.9474, 0, 1.1606, 0, 2.6667]),
        15: ([.5197, -.6259, -.3, 0, .4], [1, .1, -.3, 0, 1], [.8, .1, -.3, 0, .4], [.5313, .1, .4276, 0, .4]),
        16: ([1.0222, -.6179, .1, 0, 1], [.5313, .1, .1, 0, .4], [.8, .1, .1, 0, 1], [1, .1, .8276, 0, 1]),
    }
    t[6] = t[5]
    for idx, (v1, v4) in {17: (-.4, -1.1259), 18: (1, .2741), 19: (.3, -.4259)}.items():
        t[idx] = ([2, 0, -.3, 0, 1], [1, 0, v1, 0, 1], [1.0625, 0, .2145, 0, .4], [.5197, 0, v4, 0, .4])
    left, right, bl, br = (_A(list(s) + [0, 0, 0], dtype=float) for s in t[index])
    return left, right, {"bottom_left": bl, "bottom_right": br}


def problem(config, cells, gamma):
    """static/tests.py:7-328 -> dict(start_pos, end_pos, shock_pos, t_end, boundary, misc, initial_left, initial_right, dx, dy)."""
    c = config.lower()
    pi = np.pi
    misc = None
    lo, hi, shock, t_end, bc = 0, 1, .5, .2, "edge"
    left, right = _A([1, 0, 0, 0, 1, 0, 0, 0.]), _A([.125, 0, 0, 0, .1, 0, 0, 0])
    if "sod" in c:
        pass
    elif "sedov" in c or c == "blast":
        lo, hi, shock, t_end, bc = -10, 10, .5, .6, "wrap"
        left, right = _A([1, 0, 0, 0, 100, 0, 0, 0.]), _A([1, 0, 0, 0, 1, 0, 0, 0.])
    elif "shu" in c or "osher" in c or c == "so":
        lo, hi, shock, t_end, bc = -1, 1, -.8, .47, "edge"
        left, right = _A([3.857143, 2.629369, 0, 0, 10.3333, 0, 0, 0]), _A([0, 0, 0, 0, 1, 0, 0, 0.])
        misc = {"freq": 5, "ampl": .2, "y_offset": 1}
    elif c.startswith("sin"):
        lo, hi, shock, t_end, bc = 0, 1, 1, 1, "wrap"
        left = right = _A([0, 1, 1, 1, 1, 0, 0, 0.])
        misc = {"freq": 2, "ampl": .1, "y_offset": 2}
    elif c.startswith("gauss"):
        lo, hi, shock, t_end, bc = -1, 1, 1, 2, "wrap"
        left = right = _A([0, 1, 1, 1, 1e-6, 0, 0, 0])
        misc = {"peak_pos": 0, "ampl": .75, "fwhm": .08, "y_offset": 1}
    elif c.startswith("lin"):
        lo, hi, shock, t_end, bc = 0, 1, 1, 2 * pi, "wrap"
        left, right = _A([1, 1, 1, 1, 1 / gamma, 0, 0, 0]), _A([1, 1, 1, 1, 1 / gamma, 0, 0, 0])
        if "mhd" in c:
            left[5:] = _A([1, np.sqrt(2), .5]) * np.sqrt(4 * pi)
            right[5:] = _A([1, np.sqrt(2), .5]) * np.sqrt(4 * pi)
        misc = {"freq": 2, "ampl": 1e-6}
    elif "slow" in c:
        t_end = .08
        left, right = _A([5.6698, -1.5336, 0, 0, 100, 0, 0, 0]), _A([1, -10.5636, 0, 0, 1, 0, 0, 0])
    elif c.startswith("sq"):
        lo, hi, shock, t_end, bc = -1, 1, 1 / 3, .05, "wrap"
        left, right = _A([1, 1, 0, 0, 1, 0, 0, 0.]), _A([.01, 1, 0, 0, 1, 0, 0, 0])
    elif "ryu" in c or "jones" in c or c == "rj":
        lo, hi, shock, t_end = -.5, .5, 0, .15
        s = np.sqrt(pi)
        left, right = _A([1.08, 1.2, .01, .5, .95, 1 / s, 1.8 / s, 1 / s]), _A([1, 0, 0, 0, 1, 1 / s, 2 / s, 1 / s])
    elif "brio" in c or "wu" in c or c == "bw":
        lo, hi, shock, t_end = -.5, .5, 0, .1
        left, right = _A([1, 0, 0, 0, 1, .75, 1, 0]), _A([.125, 0, 0, 0, .1, .75, -1, 0])
    elif "kelvin" in c or "helmholtz" in c or c == "khi":
        lo, hi, shock, t_end, bc = -1, 1, 0, 4, "wrap"
        left, right = _A([2, -.5, 0, 0, 1, 0, 0, 0]), _A([1, .5, 0, 0, 1, 0, 0, 0])
        misc = {"perturb_ampl": .5, "freq": 4}
    elif "isentropic" in c or "vortex" in c or c == "ivc":
        lo, hi, shock, t_end, bc = 0, 10, 5, 1, "wrap"
        left = right = _A([1, 0, 0, 0, 1, 0, 0, 0.])
        misc = {"vortex_str": 5, "freq": 2}
    elif "orszag" in c or "tang" in c or c == "ot":
        lo, hi, shock, t_end, bc = -.5, .5, .5, pi / 2, "wrap"
        left = right = _A([25 / (36 * pi), 0, 0, 0, 5 / (12 * pi), 0, 0, 0])
        misc = {"ampl": 1 / np.sqrt(4 * pi)}
    elif "rotor" in c:
        lo, hi, shock, t_end, bc = -.5, .5, .1, .15, "wrap"
        left, right = _A([10, 0, 0, 0, 1, 0, 0, 0.]), _A([1, 0, 0, 0, 1, 0, 0, 0.])
        misc = {"omega": 20}
    elif "blast" in c and c.startswith("mhd"):
        lo, hi, shock, t_end, bc = -.5, .5, .1, .2, "wrap"
        r = 1 / np.sqrt(2)
        left, right = _A([1, 0, 0, 0, 10, r, r, 0]), _A([1, 0, 0, 0, .1, r, r, 0])
    elif "toro" in c:
        if "2" in c:
            shock, t_end = .5, .14
            left, right = _A([1, -2, 0, 0, .4, 0, 0, 0]), _A([1, 2, 0, 0, .4, 0, 0, 0])
        elif "3" in c:
            shock, t_end = .5, .012
            left, right = _A([1, 0, 0, 0, 1000, 0, 0, 0.]), _A([1, 0, 0, 0, .01, 0, 0, 0])
        elif "4" in c:
            shock, t_end = .3, .05
            left, right = _A([5.99924, 19.5975, 0, 0, 460.894, 0, 0, 0]), _A([5.99242, -6.19633, 0, 0, 46.095, 0, 0, 0])
        elif "5" in c:
            shock, t_end = .8, .012
            left, right = _A([1, -19.59745, 0, 0, 1000, 0, 0, 0]), _A([1, -19.59745, 0, 0, .01, 0, 0, 0])
        else:
            shock, t_end = .3, .2
            left, right = _A([1, .75, 0, 0, 1, 0, 0, 0]), _A([.125, 0, 0, 0, .1, 0, 0, 0])
    elif ("lax" in c or "liu" in c) or "ll" in c:
        lo, hi, shock, t_end, bc = 0, 1, .5, 2, "wrap"
        key = "ll" if "ll" in c else "liu"
        left, right, misc = _ll(int(c.replace(" ", "").split(key)[-1]))
    return {"start_pos": lo, "end_pos": hi, "shock_pos": shock, "t_end": t_end, "boundary": bc, "misc": misc,
            "initial_left": np.asarray(left, dtype=float), "initial_right": np.asarray(right, dtype=float),
            "dx": abs(hi - lo) / cells, "dy": abs(hi - lo) / cells}


# ------------------------------------------------------------------ host-side conversions used only for the ICs
def _norm3(v):
    return np.sqrt((v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]) + v[..., 2] * v[..., 2])


def _point_cons(w, gamma):
    q = np.copy(w)
    q[..., 4] = w[..., 4] / (gamma - 1) + .5 * (w[..., 0] * _norm3(w[..., 1:4]) ** 2 + _norm3(w[..., 5:8]) ** 2)
    q[..., 1:4] = w[..., 1:4] * w[..., 0][..., None]
    return q


def _d2(a, bc, axis):
    n = a.shape[axis]
    idx = np.arange(n)
    up = (idx + 1) % n if bc == "wrap" else np.minimum(idx + 1, n - 1)
    dn = (idx - 1) % n if bc == "wrap" else np.maximum(idx - 1, 0)
    return (np.take(a, up, axis=axis) - a) - (a - np.take(a, dn, axis=axis))


def _cons_from_point_prim(w, gamma, bc, high_order, ndim):
    """sim_variables.convert_primitive (generic.py:250-255) followed by fv.high_order_convert('cntr') (constructor.py:105-109)."""
    if high_order:
        w_acc, q_acc = np.copy(w), np.zeros_like(w)
        for ax in range(ndim):
            w_acc -= 1 / 24 * _d2(w, bc, ax)
            q_acc += 1 / 24 * _d2(_point_cons(w, gamma), bc, ax)
        q = _point_cons(w_acc, gamma) + q_acc
    else:
        q = _point_cons(w, gamma)
    out = np.copy(q)
    for ax in range(ndim):
        out += 1 / 24 * _d2(q, bc, ax)
    return out


def cell_centres(lo, hi, n):
    """constructor.py:13-16."""
    half = abs(hi - lo) / n / 2
    return np.linspace(lo - half, hi + half, n + 2)[1:-1]


def primitive_points(config, cells, dimension, gamma, prob=None):
    """constructor.py:18-103 — pointwise primitive initial data on the N[xN] cell centres."""
    c = config.lower()
    prob = prob or problem(c, cells, gamma)
    lo, hi, shock, par = prob["start_pos"], prob["end_pos"], prob["shock_pos"], prob["misc"]
    left, right = prob["initial_left"], prob["initial_right"]
    g = np.zeros((cells,) * dimension + (8,))
    g[:] = right
    pts = cell_centres(lo, hi, cells)
    pi = np.pi
    if dimension == 2:
        x, y = np.meshgrid(pts, pts, indexing="ij")
        mid = (hi + lo) / 2
        if c == "sedov" or "blast" in c:
            g[np.where(((x - mid) ** 2 + (y - mid) ** 2) <= (shock - mid) ** 2)] = left
        elif c.startswith("gauss"):
            r = np.sqrt((x - mid) ** 2 + (y - mid) ** 2)
            g[..., 0] = par["y_offset"] + par["ampl"] * np.exp(-((r - mid) ** 2) / par["fwhm"])
        elif c in ("khi", "kelvin-helmholtz") or ("kelvin" in c or "helmholtz" in c):
            g[np.where(y <= shock)] = left
            g[..., 2] = par["perturb_ampl"] * np.sin(par["freq"] * pi * x / (hi - lo))
        elif c in ("ivc", "vortex", "isentropic vortex"):
            xc, yc = (np.min(x) + np.max(x)) / 2, (np.min(y) + np.max(y)) / 2
            r = np.sqrt((x - xc) ** 2 + (y - yc) ** 2)
            T = 1 - (((gamma - 1) * par["vortex_str"] ** 2) / (2 * gamma * (par["freq"] * pi) ** 2)) * np.exp(1 - r ** 2)
            g[..., 0] = T ** (1 / (gamma - 1))
            g[..., 1] = (par["vortex_str"] / (par["freq"] * pi)) * np.exp((1 - r ** 2) / 2)
            g[..., 2] = (par["vortex_str"] / (par["freq"] * pi)) * np.exp((1 - r ** 2) / 2)
            g[..., 4] = T ** (gamma / (gamma - 1))
        elif "ll" in c or "lax-liu" in c:
            g[np.where(x <= shock)] = left
            g[np.where((x <= shock) & (y >= shock))] = par["bottom_left"]
            g[np.where((x > shock) & (y >= shock))] = par["bottom_right"]
        elif c in ("orszag-tang", "orszag", "tang", "ot"):
            g[np.where(y <= shock)] = left
            g[..., 1] = -np.sin(2 * pi * y)
            g[..., 2] = np.sin(2 * pi * x)
            g[..., 5] = -par["ampl"] * np.sin(2 * pi * y)
            g[..., 6] = par["ampl"] * np.sin(4 * pi * x)
        elif "rotor" in c:
            # constructor.py:69-73 assigns the rotation through a fancy-indexed copy, so only the disc state lands
            g[np.where(((x - mid) ** 2 + (y - mid) ** 2) <= (shock - mid) ** 2)] = left
        else:
            g[np.where(x < shock)] = left
    else:
        x = pts
        if c == "sedov" or c.startswith("sq"):
            g[np.where(np.abs(x) <= shock)] = left
        else:
            g[np.where(x <= shock)] = left
        if "shu" in c or "osher" in c:
            g[np.where(x > shock), 0] = par["y_offset"] + par["ampl"] * np.sin(par["freq"] * pi * x[x > shock])
        elif c.startswith("sin"):
            g[..., 0] = par["y_offset"] + par["ampl"] * np.sin(par["freq"] * pi * x)
        elif c.startswith("gauss"):
            g[..., 0] = par["y_offset"] + par["ampl"] * np.exp(-((x - par["peak_pos"]) ** 2) / par["fwhm"])
        elif c.startswith("lin"):
            pert = _A([1, -1, 1, 1, 1.5, 0, 0, 0]) * par["ampl"]
            if "mhd" in c:
                pert = _A([0, 0, -.3333333333333333, .9428090415820634, 0, -.3333333333333333, .9428090415820634, 0]) * par["ampl"]
            g += pert * np.sin(par["freq"] * pi * x)[..., None]
    return g


def initial_state(config, cells, dimension, gamma=1.4, high_order=True, boundary=None):
    """constructor.initialise(sim_variables, convert=True): conservative cell averages, C-order (N[,N],8) float64."""
    prob = problem(config, cells, gamma)
    bc = boundary or prob["boundary"]
    w = primitive_points(config, cells, dimension, gamma, prob)
    return _cons_from_point_prim(w, gamma, bc, high_order, dimension)


def theoretical_primitives(config, cells, dimension, gamma=1.4, boundary=None):
    """constructor.initialise(sim_variables) with its default convert=False: cell averages of the *primitive* initial
    data (fv.high_order_convert('cntr', ...), constructor.py:109) — what analytic.calculate_solution_error compares the
    primitive snapshot with (analytic.py:31)."""
    prob = problem(config, cells, gamma)
    bc = boundary or prob["boundary"]
    w = primitive_points(config, cells, dimension, gamma, prob)
    out = np.copy(w)
    for ax in range(dimension):
        out += 1 / 24 * _d2(w, bc, ax)
    return out


def initial_slab(config, cells_x, cells_y, x_offset, cells_x_total, gamma=1.4, high_order=True):
    """Rows [x_offset, x_offset+cells_x) of an (cells_x_total x cells_y) periodic domain tiled from the square
    ``cells_y x cells_y`` problem (period cells_y along x).  Used by the slab-decomposed weak-scaling runs, which the
    reference cannot build (constructor.py:22 is square-only)."""
    base = initial_state(config, cells_y, 2, gamma, high_order)
    rows = (np.arange(x_offset, x_offset + cells_x) % cells_y)
    return np.ascontiguousarray(base[rows])


def separable_profiles(config, cells, gamma=1.4):
    """(variable, along, values) for the 2D problems whose perturbation is a function of x only or of y only
    (constructor.py:42-44 Kelvin-Helmholtz, :62-67 Orszag-Tang): the reference's own numpy expression, evaluated on the
    1-D array of cell centres instead of the meshgrid (elementwise, so every value has the reference's bits); the
    device expands the tables (``astrea_init_profiles``).  along = 0: function of x (first index), 1: of y."""
    c = config.lower()
    prob = problem(c, cells, gamma)
    lo, hi, par = prob["start_pos"], prob["end_pos"], prob["misc"]
    pts = cell_centres(lo, hi, cells)
    pi = np.pi
    if "kelvin" in c or "helmholtz" in c or c == "khi":
        return [(2, 0, par["perturb_ampl"] * np.sin(par["freq"] * pi * pts / (hi - lo)))]
    if c in ("orszag-tang", "orszag", "tang", "ot"):
        return [(1, 1, -np.sin(2 * pi * pts)), (2, 0, np.sin(2 * pi * pts)),
                (5, 1, -par["ampl"] * np.sin(2 * pi * pts)), (6, 0, par["ampl"] * np.sin(4 * pi * pts))]
    return []


def piecewise_spec(config, cells, gamma=1.4):
    """The ``astrea_init_spec`` of a 2D problem whose pointwise primitive state is piecewise constant up to separable
    profiles (``separable_profiles``), or None (radial exponential profiles — Gauss, isentropic vortex — stay on the
    host: they are not separable and libm's exp is not reproducible bit for bit on the device).
    Regions are listed in the order constructor.py:33-75 assigns them; later ones paint over earlier ones."""
    from . import _native as N
    c = config.lower()
    prob = problem(c, cells, gamma)
    lo, hi, shock, par = prob["start_pos"], prob["end_pos"], prob["shock_pos"], prob["misc"]
    left, right = prob["initial_left"], prob["initial_right"]
    mid = (hi + lo) / 2
    if c == "sedov" or "blast" in c or "rotor" in c:
        regions = [(N.REGION_DISC_LE, mid, (shock - mid) ** 2, left)]
    elif c.startswith("gauss") or c in ("ivc", "vortex", "isentropic vortex"):
        return None
    elif "kelvin" in c or "helmholtz" in c or c == "khi" or c in ("orszag-tang", "orszag", "tang", "ot"):
        regions = [(N.REGION_Y_LE, shock, 0.0, left)]
    elif "ll" in c or "lax-liu" in c:
        regions = [(N.REGION_X_LE, shock, 0.0, left), (N.REGION_X_LE_Y_GE, shock, 0.0, par["bottom_left"]),
                   (N.REGION_X_GT_Y_GE, shock, 0.0, par["bottom_right"])]
    else:
        regions = [(N.REGION_X_LT, shock, 0.0, left)]
    half = abs(hi - lo) / cells / 2
    start, stop = lo - half, hi + half
    spec = N.InitSpec()
    spec.cells, spec.start, spec.step = cells, start, (stop - start) / (cells + 1)       # np.linspace(start, stop, cells + 2)
    spec.background[:] = list(right)
    spec.nregions = len(regions)
    for k, (kind, a, b, state) in enumerate(regions):
        spec.regions[k].kind, spec.regions[k].a, spec.regions[k].b = kind, a, b
        spec.regions[k].state[:] = list(state)
    return spec
