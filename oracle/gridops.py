"""Point-wise physics and stencil helpers of the oracle (test infrastructure; see oracle/__init__.py).

Everything here acts on arrays whose last axis holds the 8 state components
``[rho, vx|mx, vy|my, vz|mz, P|E, Bx, By, Bz]`` (static/tests.py:15).  Arithmetic keeps the
reference's operation order so that results are bit-identical with numpy on the same machine.
"""
import numpy as np

ONE_24TH = 1 / 24


def shifted(a, k, bc, axis=0):
    """``out[i] = a[bc(i + k)]`` along ``axis``: what ``np.pad(a, s, mode=bc)`` + slicing gives (fv.py:57-61)."""
    n = a.shape[axis]
    idx = np.arange(n) + k
    idx = np.mod(idx, n) if bc == "wrap" else np.clip(idx, 0, n - 1)
    return np.take(a, idx, axis=axis)


def extended(a, lo, hi, bc, axis=0):
    """``a`` with ``lo`` ghost entries in front and ``hi`` behind, ghost ``g`` holding ``a[bc(g)]``."""
    n = a.shape[axis]
    idx = np.arange(-lo, n + hi)
    idx = np.mod(idx, n) if bc == "wrap" else np.clip(idx, 0, n - 1)
    return np.take(a, idx, axis=axis)


def safe_div(num, den):
    """fv.py:19-20 — quotient, 0 where the divisor is exactly 0."""
    # LAPACK may hand back a complex-typed spectrum whose imaginary parts are all zero (Lax-Wendroff path)
    num = num.real if np.iscomplexobj(num) else num
    den = den.real if np.iscomplexobj(den) else den
    num, den = np.broadcast_arrays(np.asarray(num, dtype=float), np.asarray(den, dtype=float))
    out = np.zeros(num.shape)
    np.divide(num, den, out=out, where=den != 0)
    return out


def length(vec):
    """fv.py:37-38 — Euclidean norm over the last axis: sqrt((x0^2 + x1^2) + x2^2)."""
    return np.sqrt((vec[..., 0] * vec[..., 0] + vec[..., 1] * vec[..., 1]) + vec[..., 2] * vec[..., 2])


def second_difference(a, bc, axis=0):
    """fv.py:42-44 on a pad-1 array: ``(a[i+1] - a[i]) - (a[i] - a[i-1])`` (undivided)."""
    up, dn = shifted(a, 1, bc, axis), shifted(a, -1, bc, axis)
    return (up - a) - (a - dn)


def pressure_from_cons(q, gamma):
    """fv.py:52-53."""
    vel = safe_div(q[..., 1:4], q[..., 0][..., None])
    return (gamma - 1) * (q[..., 4] - .5 * (q[..., 0] * length(vel) ** 2 + length(q[..., 5:8]) ** 2))


def energy_from_prim(w, gamma):
    """fv.py:50-51."""
    return w[..., 4] / (gamma - 1) + .5 * (w[..., 0] * length(w[..., 1:4]) ** 2 + length(w[..., 5:8]) ** 2)


def prim_of_cons(q, gamma):
    """fv.py:97-101 (pointwise q -> w)."""
    w = np.copy(q)
    w[..., 4] = pressure_from_cons(q, gamma)
    w[..., 1:4] = safe_div(q[..., 1:4], q[..., 0][..., None])
    return w


def cons_of_prim(w, gamma):
    """fv.py:89-93 (pointwise w -> q)."""
    q = np.copy(w)
    q[..., 4] = energy_from_prim(w, gamma)
    q[..., 1:4] = w[..., 1:4] * w[..., 0][..., None]
    return q


def _stencil_axes(ndim_space, kind):
    """fv.py:108-111 — 'cell': every spatial axis of the passed array; 'face': axes >= 1 only."""
    return range(1, ndim_space) if kind == "face" else range(ndim_space)


def prim_avg_of_cons_avg(qbar, cfg, kind="cell"):
    """convert_conservative selector (generic.py:250-255): fv.py:126-143 when 4th-order, else fv.py:97-101."""
    if not cfg.high_order:
        return prim_of_cons(qbar, cfg.gamma)
    q_acc, w_acc = np.copy(qbar), np.zeros_like(qbar)
    for ax in _stencil_axes(cfg.dimension, kind):
        q_acc -= ONE_24TH * second_difference(qbar, cfg.boundary, ax)
        w_acc += ONE_24TH * second_difference(prim_of_cons(qbar, cfg.gamma), cfg.boundary, ax)
    return prim_of_cons(q_acc, cfg.gamma) + w_acc


def cons_avg_of_prim_avg(wbar, cfg, kind="cell"):
    """convert_primitive selector (generic.py:250-255): fv.py:105-122 when 4th-order, else fv.py:89-93."""
    if not cfg.high_order:
        return cons_of_prim(wbar, cfg.gamma)
    w_acc, q_acc = np.copy(wbar), np.zeros_like(wbar)
    for ax in _stencil_axes(cfg.dimension, kind):
        w_acc -= ONE_24TH * second_difference(wbar, cfg.boundary, ax)
        q_acc += ONE_24TH * second_difference(cons_of_prim(wbar, cfg.gamma), cfg.boundary, ax)
    return cons_of_prim(w_acc, cfg.gamma) + q_acc


def centred_of_avg(x, cfg, kind="cell"):
    """fv.py:67-85 with num_scheme 'avg': x - sum_ax d2x/24."""
    out = np.copy(x)
    for ax in _stencil_axes(cfg.dimension, kind):
        out -= ONE_24TH * second_difference(x, cfg.boundary, ax)
    return out


def avg_of_centred(x, cfg, kind="cell"):
    """fv.py:67-85 with num_scheme 'cntr': x + sum_ax d2x/24."""
    out = np.copy(x)
    for ax in _stencil_axes(cfg.dimension, kind):
        out += ONE_24TH * second_difference(x, cfg.boundary, ax)
    return out


def physical_flux(w, gamma, axis):
    """constructor.py:113-125 — ideal-MHD flux along ``axis`` from primitive variables."""
    n, t1, t2 = axis % 3, (axis + 1) % 3, (axis + 2) % 3
    rho, P = w[..., 0], w[..., 4]
    v, B = w[..., 1:4], w[..., 5:8]
    vn, Bn = v[..., axis], B[..., axis]
    f = np.zeros_like(w)
    f[..., 0] = rho * vn
    f[..., n + 1] = rho * vn ** 2 + P + .5 * length(B) ** 2 - Bn ** 2
    f[..., t1 + 1] = rho * vn * v[..., t1] - Bn * B[..., t1]
    f[..., t2 + 1] = rho * vn * v[..., t2] - Bn * B[..., t2]
    vdotB = (v[..., 0] * B[..., 0] + v[..., 1] * B[..., 1]) + v[..., 2] * B[..., 2]
    f[..., 4] = vn * (.5 * rho * length(v) ** 2 + (gamma * P) / (gamma - 1) + length(B) ** 2) - Bn * vdotB
    f[..., t1 + 5] = B[..., t1] * vn - Bn * v[..., t1]
    f[..., t2 + 5] = B[..., t2] * vn - Bn * v[..., t2]
    return f


def roe_state(first, second):
    """constructor.py:167-176 with (left_interface, right_interface) = (first, second).

    Every caller passes (w_plus, w_minus).  Velocity and pressure are weighted
    first*sqrt(rho_first) + second*sqrt(rho_second); the magnetic field is weighted the other
    way round.
    """
    s2, s1 = np.sqrt(second[..., 0]), np.sqrt(first[..., 0])
    out = np.zeros_like(first)
    out[..., 0] = s2 * s1
    out[..., 1:4] = safe_div(first[..., 1:4] * s1[..., None] + second[..., 1:4] * s2[..., None], (s2 + s1)[..., None])
    out[..., 4] = safe_div(s1 * first[..., 4] + s2 * second[..., 4], s2 + s1)
    out[..., 5:8] = safe_div(first[..., 5:8] * s2[..., None] + second[..., 5:8] * s1[..., None], (s2 + s1)[..., None])
    return out


def primitive_jacobian(w, gamma, axis):
    """constructor.py:129-163 — 8x8 Jacobian dF/dw in primitive variables at every point."""
    n, t1, t2 = axis % 3, (axis + 1) % 3, (axis + 2) % 3
    rho, P = w[..., 0], w[..., 4]
    v, B = w[..., 1:4], w[..., 5:8]
    A = np.zeros(w.shape + (8,))
    for d in range(8):
        A[..., d, d] = v[..., n]
    A[..., 0, n + 1] = rho
    A[..., n + 1, 4] = 1 / rho
    A[..., 4, n + 1] = gamma * P
    A[..., n + 5, n + 5] = 0
    bn, b1, b2 = safe_div(B[..., n], rho), safe_div(B[..., t1], rho), safe_div(B[..., t2], rho)
    A[..., n + 1, n + 5] = -bn
    A[..., n + 1, t1 + 5] = b1
    A[..., n + 1, t2 + 5] = b2
    A[..., t1 + 1, t1 + 5] = -bn
    A[..., t2 + 1, t2 + 5] = -bn
    A[..., t1 + 1, n + 5] = -b1
    A[..., t2 + 1, n + 5] = -b2
    A[..., t1 + 5, n + 1] = B[..., t1]
    A[..., t2 + 5, n + 1] = B[..., t2]
    A[..., 4, n + 5] = (gamma - 1) * ((v[..., 0] * B[..., 0] + v[..., 1] * B[..., 1]) + v[..., 2] * B[..., 2])
    A[..., t1 + 5, t1 + 1] = -B[..., n]
    A[..., t2 + 5, t2 + 1] = -B[..., n]
    A[..., t1 + 5, n + 5] = -v[..., t1]
    A[..., t2 + 5, n + 5] = -v[..., t2]
    return A


def closed_spectral_radius(w, gamma, axis):
    """max|eig(primitive_jacobian)| in closed form (what the device computes instead of calling LAPACK).

    The spectrum of constructor.py:129-163 is {0, v, v +- sqrt(x)} for x in {c_a^2, c_f^2, c_s^2}.
    For physical states that is ``|v_n| + c_f`` (agrees with LAPACK to 5.6e-15).  Unphysical
    reconstructed states (negative pressure) give x < 0, i.e. the complex pair v +- i sqrt(-x) whose modulus
    sqrt(v^2 - x) is what np.abs(np.linalg.eigvals(..)) returns at fv.py:157-162.
    """
    with np.errstate(all="ignore"):
        rho, P, B = w[..., 0], w[..., 4], w[..., 5:8]
        a2 = gamma * P / rho
        b2 = ((B[..., 0] * B[..., 0] + B[..., 1] * B[..., 1]) + B[..., 2] * B[..., 2]) / rho
        bn2 = B[..., axis] * B[..., axis] / rho
        s = a2 + b2
        disc = s * s - 4 * (a2 * bn2)
        vn = np.abs(w[..., axis + 1])

        def modulus(x):
            return np.where(x >= 0, vn + np.sqrt(np.abs(x)), np.sqrt(vn * vn + (-x)))

        root = np.sqrt(np.abs(disc))
        cf2, cs2 = .5 * (s + root), .5 * (s - root)
        physical = vn + np.sqrt(np.abs(cf2))
        real_roots = np.maximum(np.maximum(modulus(cf2), modulus(cs2)), modulus(bn2))
        p, q = .5 * s, .5 * np.sqrt(np.abs(disc))
        m = np.sqrt(p * p + q * q)
        al, be = np.sqrt(np.abs(.5 * (m + p))), np.sqrt(np.abs(.5 * (m - p)))
        complex_roots = np.maximum(np.sqrt((vn + al) * (vn + al) + be * be), modulus(bn2))
        return np.where(disc >= 0, np.where((cs2 >= 0) & (bn2 >= 0), physical, real_roots),
                        np.where(disc < 0, complex_roots, disc))


def spectral_radius(w, cfg, axis):
    """fv.py:157-162 — per point max|lambda| of the primitive Jacobian (+ the raw spectrum for Lax-Wendroff)."""
    if cfg.eigen == "closed":
        return None, closed_spectral_radius(w, cfg.gamma, axis)
    spectrum = np.linalg.eigvals(primitive_jacobian(w, cfg.gamma, axis))
    return spectrum, np.max(np.abs(spectrum), axis=-1)
