"""ctypes binding of the C ABI declared in include/astrea_b200.h.

The product library is ``astrea_b200/lib/libastrea_b200.so`` (nvcc, sm_100a; ``python -m astrea_b200.build``).
There is no CPU fallback: if the library is missing or was not built for the device, loading raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ASTREA_B200_LIB: path of an alternative *device* build (kernel-tuning experiments); still refused unless sm_100a
DEVICE_LIB = os.environ.get("ASTREA_B200_LIB") or os.path.join(_HERE, "lib", "libastrea_b200.so")

# enums of include/astrea_b200.h
PCM, PLM, PPM, WENO3, WENO5, WENO7 = range(6)
PPM_MC, PPM_COLELLA, PPM_PH = range(3)
MINMOD, VANLEER, OSPRE, VANALBADA, KOREN, SUPERBEE = range(6)
LLF, LW, HLLC, HLLD = range(4)
EULER, RK4, SSPRK22, SSPRK33, SSPRK43, SSPRK53, SSPRK54, SSPRK104 = range(8)
EDGE, WRAP = 0, 1
E_ARG, E_CUDA, E_NONFINITE, E_STATE = -1, -2, -3, -4


MAX_REGIONS = 8
REGION_X_LT, REGION_X_LE, REGION_Y_LE, REGION_X_LE_Y_GE, REGION_X_GT_Y_GE, REGION_DISC_LE = range(6)


class Region(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("a", C.c_double), ("b", C.c_double), ("state", C.c_double * 8)]


class InitSpec(C.Structure):
    _fields_ = [("cells", C.c_int64), ("start", C.c_double), ("step", C.c_double), ("background", C.c_double * 8),
                ("nregions", C.c_int32), ("reserved", C.c_int32), ("regions", Region * MAX_REGIONS)]


MAX_PROFILES = 4


class InitProfile(C.Structure):
    _fields_ = [("variable", C.c_int32), ("along", C.c_int32), ("values", C.POINTER(C.c_double))]


class Cfg(C.Structure):
    """struct astrea_cfg."""
    _fields_ = [
        ("dimension", C.c_int32), ("boundary", C.c_int32), ("nx", C.c_int64), ("ny", C.c_int64),
        ("gamma", C.c_double), ("dx", C.c_double), ("cfl", C.c_double),
        ("scheme", C.c_int32), ("ppm_author", C.c_int32), ("limiter", C.c_int32), ("solver", C.c_int32),
        ("low_mach", C.c_int32), ("integrator", C.c_int32), ("magnetic_2d", C.c_int32), ("device", C.c_int32),
        ("nx_global", C.c_int64), ("x_offset", C.c_int64),
        ("threads_2d", C.c_int32), ("segment_2d", C.c_int32), ("tile_1d", C.c_int32), ("flags", C.c_int32),
    ]


class AstreaError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"astrea_b200 error {code}: {text}")
        self.code = code


class NonFiniteError(AstreaError, np.linalg.LinAlgError):
    """Where the reference's np.linalg.eigvals raises LinAlgError (fv.py:158)."""


_PD = C.POINTER(C.c_double)
REDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int)
_SIGNATURES = {
    "astrea_create": (C.c_void_p, [C.POINTER(Cfg)]),
    "astrea_destroy": (None, [C.c_void_p]),
    "astrea_last_error": (C.c_char_p, [C.c_void_p]),
    "astrea_upload": (C.c_int, [C.c_void_p, C.c_void_p]),
    "astrea_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "astrea_diagnostics": (C.c_int, [C.c_void_p, _PD, _PD, C.c_int]),
    "astrea_evolve_space": (C.c_int, [C.c_void_p, C.c_int, _PD]),
    "astrea_evolve_time": (C.c_int, [C.c_void_p, C.c_double]),
    "astrea_step": (C.c_int, [C.c_void_p, C.c_double, C.c_double, _PD]),
    "astrea_set_time": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "astrea_step_async": (C.c_int, [C.c_void_p]),
    "astrea_run_steps": (C.c_int, [C.c_void_p, C.c_int64]),
    "astrea_get_time": (C.c_int, [C.c_void_p, _PD, C.POINTER(C.c_int64), _PD]),
    "astrea_dt_history": (C.c_int, [C.c_void_p, _PD, C.c_int]),
    "astrea_dt_async": (C.c_int, [C.c_void_p]),
    "astrea_get_parity": (C.c_int, [C.c_void_p]),
    "astrea_set_parity": (C.c_int, [C.c_void_p, C.c_int]),
    "astrea_download_face_field": (C.c_int, [C.c_void_p, C.c_void_p]),
    "astrea_set_flag_reducer": (C.c_int, [C.c_void_p, REDUCE_FN, C.c_void_p]),
    "astrea_set_key_reducer": (C.c_int, [C.c_void_p, REDUCE_FN, C.c_void_p]),
    "astrea_program_length": (C.c_int, [C.c_void_p]),
    "astrea_instr_is_operator": (C.c_int, [C.c_void_p, C.c_int]),
    "astrea_instr_needs_halo": (C.c_int, [C.c_void_p, C.c_int]),
    "astrea_instr_is_update": (C.c_int, [C.c_void_p, C.c_int]),
    "astrea_run_update_part": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "astrea_set_dt": (C.c_int, [C.c_void_p, C.c_double]),
    "astrea_run_instr": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "astrea_finish_step": (C.c_int, [C.c_void_p]),
    "astrea_halo_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "astrea_halo_ptrs": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(_PD), C.POINTER(_PD), C.POINTER(_PD), C.POINTER(_PD)]),
    "astrea_halo_prepare": (C.c_int, [C.c_void_p, C.c_int]),
    "astrea_eigmax_device": (C.c_int, [C.c_void_p, C.POINTER(_PD)]),
    "astrea_read_eigmax": (C.c_int, [C.c_void_p, _PD]),
    "astrea_sync": (C.c_int, [C.c_void_p]),
    "astrea_stream_handle": (C.c_uint64, [C.c_void_p]),
    "astrea_fp64_probe": (C.c_int, [C.c_void_p, _PD]),
    "astrea_init_piecewise": (C.c_int, [C.c_void_p, C.POINTER(InitSpec)]),
    "astrea_init_profiles": (C.c_int, [C.c_void_p, C.POINTER(InitSpec), C.c_int, C.POINTER(InitProfile)]),
    "astrea_arith_check": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, C.POINTER(C.c_uint64)]),
    "astrea_launch_count": (C.c_int64, [C.c_void_p]),
    "astrea_save_state": (C.c_int, [C.c_void_p]),
    "astrea_restore_state": (C.c_int, [C.c_void_p]),
    "astrea_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "astrea_profile_read": (C.c_int, [C.c_void_p, _PD, C.POINTER(C.c_int64)]),
    "astrea_snapshot_begin": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int]),
    "astrea_snapshot_wait": (C.c_int, [C.c_void_p, C.c_int64]),
    "astrea_solution_error": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, _PD, C.c_int]),
    "astrea_ppm_flattener": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, _PD, C.c_void_p]),
    "astrea_ppm_viscosity": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, _PD, C.c_void_p]),
    "astrea_host_alloc": (C.c_void_p, [C.c_int, C.c_uint64]),
    "astrea_host_free": (None, [C.c_void_p]),
    "astrea_is_device_build": (C.c_int, []),
}
EXPORTS = tuple(_SIGNATURES)


def bind(path):
    """dlopen ``path`` and attach the prototypes of include/astrea_b200.h."""
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


_device_lib = None


def device_library():
    """The sm_100a library.  Raises if it has not been built: the package has no other execution path."""
    global _device_lib
    if _device_lib is None:
        if not os.path.exists(DEVICE_LIB):
            raise ImportError(f"{DEVICE_LIB} is missing: build it with `python -m astrea_b200.build` "
                              "(astrea_b200 has no CPU fallback)")
        lib = bind(DEVICE_LIB)
        if lib.astrea_is_device_build() != 1:
            raise ImportError(f"{DEVICE_LIB} is not an sm_100a device build")
        _device_lib = lib
    return _device_lib


class PinnedPool:
    """numpy arrays in page-locked host memory (``astrea_host_alloc``), recycled when the caller drops them.

    The drop-in's ``evolve_time`` returns its result in such an array; the reference loop rebinds ``grid`` to it and
    passes it to the next ``evolve_space`` (astrea.py:67,81), so after the first step both transfers are direct DMA.
    A block goes back to the free list when the last view of its array is garbage collected."""

    def __init__(self, lib, device=0, keep=3):
        self.lib, self.device, self.keep = lib, device, keep
        self.free = {}          # bytes -> [address, ...]
        self.closed = False

    class _Block:
        def __init__(self, pool, address, nbytes):
            self.pool, self.address, self.nbytes = pool, address, nbytes

        def __del__(self):
            try:
                self.pool._give_back(self.address, self.nbytes)
            except Exception:
                pass

    def _give_back(self, address, nbytes):
        spare = self.free.setdefault(nbytes, [])
        if self.closed or len(spare) >= self.keep:
            self.lib.astrea_host_free(address)
        else:
            spare.append(address)

    def empty(self, shape):
        """An uninitialised C-contiguous float64 array of ``shape`` in pinned memory."""
        count = int(np.prod(shape))
        nbytes = max(count, 1) * 8
        spare = self.free.get(nbytes)
        address = spare.pop() if spare else self.lib.astrea_host_alloc(self.device, nbytes)
        if not address:
            return np.empty(shape, dtype=np.float64)       # out of pinned memory: a pageable array still works
        buf = (C.c_double * max(count, 1)).from_address(address)
        buf._astrea_block = PinnedPool._Block(self, address, nbytes)      # lives as long as any view of the array
        return np.frombuffer(buf, dtype=np.float64, count=count).reshape(shape)

    def close(self):
        self.closed = True
        for spare in self.free.values():
            for address in spare:
                self.lib.astrea_host_free(address)
        self.free.clear()


class Context:
    """One ``astrea_ctx``: a grid (or slab) resident on one GPU plus the step program of its integrator."""

    def __init__(self, cfg, lib=None):
        self.lib = lib if lib is not None else device_library()
        self.cfg = cfg
        self._h = self.lib.astrea_create(C.byref(cfg))
        if not self._h:
            raise AstreaError(E_ARG, self.lib.astrea_last_error(None).decode())
        self.shape = (cfg.nx, 8) if cfg.dimension == 1 else (cfg.nx, cfg.ny, 8)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.astrea_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, code):
        if code != 0:
            text = self.lib.astrea_last_error(self._h).decode()
            raise (NonFiniteError if code == E_NONFINITE else AstreaError)(code, text)

    # -- data movement
    def upload(self, grid):
        g = np.ascontiguousarray(grid, dtype=np.float64)
        if g.shape != tuple(self.shape):
            raise ValueError(f"grid shape {g.shape} != {tuple(self.shape)}")
        self._check(self.lib.astrea_upload(self._h, g.ctypes.data))

    def upload_ptr(self, host_ptr):
        self._check(self.lib.astrea_upload(self._h, host_ptr))

    def download(self, primitive=False, out=None):
        """``primitive``: False = conservative averages, True = the astrea.py:47 snapshot, 2 = the same on a slab whose
        ghost rows the caller has exchanged."""
        if out is None:
            out = np.empty(self.shape, dtype=np.float64)
        elif (not isinstance(out, np.ndarray) or out.shape != tuple(self.shape) or out.dtype != np.float64
              or not out.flags.c_contiguous or not out.flags.writeable):
            raise ValueError(f"out must be a writeable C-contiguous float64 array of shape {tuple(self.shape)}")
        self._check(self.lib.astrea_download(self._h, out.ctypes.data, int(primitive)))
        return out

    def snapshot_begin(self, out, external_rows=False):
        """Start the astrea.py:47 snapshot (primitive, transposed by ortho_axis) into ``out`` — shape (ny, nx, 8) in 2D —
        and return a ticket; ``snapshot_wait(ticket)`` blocks until ``out`` is complete."""
        want = (self.shape[1], self.shape[0], 8) if len(self.shape) == 3 else tuple(self.shape)
        if (not isinstance(out, np.ndarray) or out.shape != want or out.dtype != np.float64 or not out.flags.c_contiguous
                or not out.flags.writeable):
            raise ValueError(f"out must be a writeable C-contiguous float64 array of shape {want}")
        ticket = self.lib.astrea_snapshot_begin(self._h, out.ctypes.data, 1 if external_rows else 0)
        if ticket < 0:
            self._check(int(ticket))
        return int(ticket)

    def snapshot_wait(self, ticket):
        self._check(self.lib.astrea_snapshot_wait(self._h, int(ticket)))

    def diagnostics(self, external_rows=False):
        """(totals[8], total_variation[8]) of the current grid, reduced on the device (functions/analytic.py:48-77)."""
        tot, tv = (C.c_double * 8)(), (C.c_double * 8)()
        self._check(self.lib.astrea_diagnostics(self._h, tot, tv, 1 if external_rows else 0))
        return np.array(tot), np.array(tv)

    def download_ptr(self, host_ptr, primitive=False):
        self._check(self.lib.astrea_download(self._h, host_ptr, 1 if primitive else 0))

    # -- the two calls of the reference seam
    def evolve_space(self, parity):
        eig = (C.c_double * 2)()
        self._check(self.lib.astrea_evolve_space(self._h, int(parity), eig))
        return [eig[a] for a in range(self.cfg.dimension)]

    def evolve_time(self, dt):
        self._check(self.lib.astrea_evolve_time(self._h, float(dt)))

    def step(self, t=0.0, t_stop=0.0):
        dt = C.c_double()
        self._check(self.lib.astrea_step(self._h, float(t), float(t_stop), C.byref(dt)))
        return dt.value

    # -- the same without host round trips (device-side dt and clock)
    def set_time(self, t=0.0, t_stop=0.0):
        self._check(self.lib.astrea_set_time(self._h, float(t), float(t_stop)))

    def step_async(self):
        self._check(self.lib.astrea_step_async(self._h))

    def run_steps(self, nsteps):
        """``nsteps`` x step_async in one call."""
        self._check(self.lib.astrea_run_steps(self._h, int(nsteps)))

    def dt_async(self):
        self._check(self.lib.astrea_dt_async(self._h))

    def get_time(self):
        """(t, steps, last dt) after synchronising; raises NonFiniteError if a step since the last check went bad."""
        t, n, dt = C.c_double(), C.c_int64(), C.c_double()
        self._check(self.lib.astrea_get_time(self._h, C.byref(t), C.byref(n), C.byref(dt)))
        return t.value, n.value, dt.value

    def dt_history(self, n):
        out = (C.c_double * n)()
        self._check(self.lib.astrea_dt_history(self._h, out, n))
        return list(out)

    @property
    def parity(self):
        return self.lib.astrea_get_parity(self._h)

    @parity.setter
    def parity(self, p):
        self._check(self.lib.astrea_set_parity(self._h, int(p)))

    def set_flag_reducer(self, fn):
        """``fn(device_ptr, count)``: in-place cross-rank maximum of ``count`` int32 on the context's stream (PPM authors
        'c' / 'ph' on a decomposed grid)."""
        def trampoline(_user, ptr, count):
            try:
                fn(ptr, count)
                return 0
            except Exception:          # an exception must not cross the C frame
                return 1
        self._reducer = REDUCE_FN(trampoline)       # keep the callback object alive as long as the context
        self._check(self.lib.astrea_set_flag_reducer(self._h, self._reducer, None))

    def set_key_reducer(self, fn):
        """``fn(device_ptr, count)``: in-place cross-rank minimum of ``count`` uint64 on the context's stream (the
        Lax-Wendroff column search on a decomposed grid)."""
        def trampoline(_user, ptr, count):
            try:
                fn(ptr, count)
                return 0
            except Exception:
                return 1
        self._key_reducer = REDUCE_FN(trampoline)
        self._check(self.lib.astrea_set_key_reducer(self._h, self._key_reducer, None))

    # -- step program (multi-GPU hosts drive it instruction by instruction)
    def program(self):
        n = self.lib.astrea_program_length(self._h)
        return [bool(self.lib.astrea_instr_is_operator(self._h, i)) for i in range(n)]

    def halo_readers(self):
        n = self.lib.astrea_program_length(self._h)
        return [self.lib.astrea_instr_needs_halo(self._h, i) == 1 for i in range(n)]

    def updates(self):
        n = self.lib.astrea_program_length(self._h)
        return [self.lib.astrea_instr_is_update(self._h, i) == 1 for i in range(n)]

    def run_update_part(self, i, part):
        self._check(self.lib.astrea_run_update_part(self._h, i, int(part)))

    def set_dt(self, dt):
        self._check(self.lib.astrea_set_dt(self._h, float(dt)))

    def run_instr(self, i, external_rows=False):
        self._check(self.lib.astrea_run_instr(self._h, i, 1 if external_rows else 0))

    def finish_step(self):
        self._check(self.lib.astrea_finish_step(self._h))

    def halo_info(self):
        rows, n = C.c_int64(), C.c_int64()
        self._check(self.lib.astrea_halo_info(self._h, C.byref(rows), C.byref(n)))
        return rows.value, n.value

    def halo_ptrs(self, i):
        p = [_PD() for _ in range(4)]
        self._check(self.lib.astrea_halo_ptrs(self._h, i, *[C.byref(x) for x in p]))
        return [C.cast(x, C.c_void_p).value for x in p]   # send_lo, send_hi, recv_lo, recv_hi

    def halo_prepare(self, i):
        self._check(self.lib.astrea_halo_prepare(self._h, i))

    def eigmax_device(self):
        p = _PD()
        self._check(self.lib.astrea_eigmax_device(self._h, C.byref(p)))
        return C.cast(p, C.c_void_p).value

    def read_eigmax(self):
        eig = (C.c_double * 2)()
        self._check(self.lib.astrea_read_eigmax(self._h, eig))
        return [eig[a] for a in range(self.cfg.dimension)]

    def fp64_probe(self):
        """Sustained fp64 FMA rate of this GPU in TFLOP/s (measurement aid of bench.py)."""
        t = C.c_double()
        self._check(self.lib.astrea_fp64_probe(self._h, C.byref(t)))
        return t.value

    def init_piecewise(self, spec, profiles=()):
        """Initial conditions evaluated on the device (``initial.piecewise_spec``) instead of an upload.  ``profiles``:
        (variable, along, values) triples of separable 1-D profiles (``initial.separable_profiles``)."""
        if not profiles:
            self._check(self.lib.astrea_init_piecewise(self._h, C.byref(spec)))
            return
        arr = (InitProfile * len(profiles))()
        keep = []
        for k, (variable, along, values) in enumerate(profiles):
            v = np.ascontiguousarray(values, dtype=np.float64)
            if v.shape != (spec.cells,):
                raise ValueError("a profile holds one value per cell centre")
            keep.append(v)
            arr[k].variable, arr[k].along = int(variable), int(along)
            arr[k].values = v.ctypes.data_as(C.POINTER(C.c_double))
        self._check(self.lib.astrea_init_profiles(self._h, C.byref(spec), len(profiles), arr))

    def arith_check(self, samples=1 << 24, seed=1):
        """(accepted, wrong, declined) of the branch-free division / square root against the IEEE routines."""
        counts = (C.c_uint64 * 3)()
        self._check(self.lib.astrea_arith_check(self._h, samples, seed, counts))
        return int(counts[0]), int(counts[1]), int(counts[2])

    def solution_error(self, w_theo, norm, external_rows=False):
        """Per-channel reduction of |w_num - w_theo| (functions/analytic.py:24-44) before normalisation: 10 doubles."""
        w = np.ascontiguousarray(w_theo, dtype=np.float64)
        if w.shape != tuple(self.shape):
            raise ValueError(f"w_theo shape {w.shape} != {tuple(self.shape)}")
        out = (C.c_double * 10)()
        self._check(self.lib.astrea_solution_error(self._h, w.ctypes.data, float(norm), out, 1 if external_rows else 0))
        return np.array(out)

    def ppm_flattener(self, ws, axis, slope_determinants=None):
        """ppm.apply_flattener (ppm.py:111-134) of primitive averages ``ws`` (sweep frame): chi, shape of the grid."""
        w = np.ascontiguousarray(ws, dtype=np.float64)
        if w.shape != tuple(self.shape):
            raise ValueError(f"wS shape {w.shape} != {tuple(self.shape)}")
        knobs = None if slope_determinants is None else (C.c_double * 3)(*slope_determinants)
        out = np.empty(self.shape[:-1], dtype=np.float64)
        self._check(self.lib.astrea_ppm_flattener(self._h, w.ctypes.data, int(axis), knobs, out.ctypes.data))
        return out

    def ppm_viscosity(self, ws, axis, viscosity_determinants=None):
        """ppm.apply_artificial_viscosity (ppm.py:138-170) of primitive averages ``ws``: mu, shape of ``ws`` (1D)."""
        w = np.ascontiguousarray(ws, dtype=np.float64)
        if w.shape != tuple(self.shape):
            raise ValueError(f"wS shape {w.shape} != {tuple(self.shape)}")
        knobs = None if viscosity_determinants is None else (C.c_double * 2)(*viscosity_determinants)
        out = np.empty(self.shape, dtype=np.float64)
        self._check(self.lib.astrea_ppm_viscosity(self._h, w.ctypes.data, int(axis), knobs, out.ctypes.data))
        return out

    def sync(self):
        self._check(self.lib.astrea_sync(self._h))

    def save_state(self):
        self._check(self.lib.astrea_save_state(self._h))

    def restore_state(self):
        self._check(self.lib.astrea_restore_state(self._h))

    def profile(self, enable=True):
        self._check(self.lib.astrea_profile(self._h, 1 if enable else 0))

    def profile_read(self):
        """{class: (milliseconds, launches)} since the last read."""
        ms, n = (C.c_double * 6)(), (C.c_int64 * 6)()
        self._check(self.lib.astrea_profile_read(self._h, ms, n))
        return {name: (ms[k], n[k]) for k, name in enumerate(("flux", "transpose", "update", "halo", "prim", "recon"))}

    @property
    def stream_handle(self):
        return self.lib.astrea_stream_handle(self._h)

    @property
    def launch_count(self):
        return self.lib.astrea_launch_count(self._h)
