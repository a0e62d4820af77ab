"""Device-resident time loop: the body of ``core_run`` (astrea.py:31-85) with the grid kept in HBM.

Single GPU: every step is ``astrea_step`` (operator, dt = cfl*min(dx/eigmax), Runge-Kutta stages, parity flip).

Several GPUs (one process per GPU, ``torch.distributed``): the grid is cut into slabs along x.
The host walks the step program of the context; before each spatial-operator evaluation it exchanges GHOST
ghost rows of the register that operator reads with the two neighbouring ranks (NCCL send/recv on the context's
stream; the ring closes for periodic boundaries, a physical 'edge' boundary is filled locally), and after the
first operator it all-reduces (MAX) the two per-axis wave speeds so that every rank takes the same dt
(astrea.py:70-71).  No other collective is on the path.
"""
import ctypes

import os
import numpy as np

from . import _native as N
from .initial import initial_slab, initial_state, piecewise_spec, problem, separable_profiles
from .selectors import MAGNETIC_2D, make_cfg, scheme_enum, stages_of


class _CudaBlock:
    """A raw device address as something torch can wrap (``__cuda_array_interface__``)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3,
                                         "strides": None}


def _tensor_at(ptr, count, on_device, device_index=0, int32=False, int64=False):
    import torch
    if on_device:
        block = _CudaBlock(ptr, count)
        if int32 or int64:
            block.__cuda_array_interface__["typestr"] = "<i4" if int32 else "<i8"
        return torch.as_tensor(block, device=torch.device("cuda", device_index))
    ctype, dtype = (ctypes.c_int32, np.int32) if int32 else ((ctypes.c_int64, np.int64) if int64 else (ctypes.c_double, np.float64))
    buf = (ctype * count).from_address(ptr)
    return torch.from_numpy(np.frombuffer(buf, dtype=dtype))


class SlabExchange:
    """Halo rows and the wave-speed reduction of one rank, over ``torch.distributed`` (nccl on GPUs, gloo in tests)."""

    def __init__(self, ctx, rank, world, periodic, on_device, device_index=0):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.ctx, self.rank, self.world, self.periodic = ctx, rank, world, periodic
        self.on_device, self.device_index = on_device, device_index
        self.rows, self.count = ctx.halo_info()
        self.lo = (rank - 1) % world if (periodic or rank > 0) else None
        self.hi = (rank + 1) % world if (periodic or rank < world - 1) else None
        self._views = {}
        self.stream = torch.cuda.ExternalStream(ctx.stream_handle, device=device_index) if on_device else None
        # high priority: the exchange kernels must get SMs while the interior update still fills the device
        self.comm = torch.cuda.Stream(device=device_index, priority=-1) if on_device else None
        self._pending = None
        self._eig = _tensor_at(ctx.eigmax_device(), 3, on_device, device_index)

    def _blocks(self, instr):
        if instr not in self._views:
            self._views[instr] = [_tensor_at(p, self.count, self.on_device, self.device_index) for p in self.ctx.halo_ptrs(instr)]
        return self._views[instr]

    def _ops(self, instr):
        dist = self.dist
        send_lo, send_hi, recv_lo, recv_hi = self._blocks(instr)

        def op(kind, block, peer, tag):
            return dist.P2POp(kind, block, peer) if self.on_device else dist.P2POp(kind, block, peer, tag=tag)

        # order matters when both neighbours are the same peer (world == 2): the first message travels "upwards"
        ops = []
        if self.hi is not None:
            ops.append(op(dist.isend, send_hi, self.hi, 0))
        if self.lo is not None:
            ops.append(op(dist.isend, send_lo, self.lo, 1))
        if self.lo is not None:
            ops.append(op(dist.irecv, recv_lo, self.lo, 0))
        if self.hi is not None:
            ops.append(op(dist.irecv, recv_hi, self.hi, 1))
        return ops

    def halo(self, instr):
        """Exchange the ghost rows of the register operator ``instr`` reads, in order on the context's stream."""
        self.ctx.halo_prepare(instr)
        ops = self._ops(instr)
        if not ops:
            return
        if self.on_device:
            with self.torch.cuda.stream(self.stream):
                for w in self.dist.batch_isend_irecv(ops):
                    w.wait()
        else:
            self.ctx.sync()
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()

    def start(self, instr):
        """Begin the same exchange on the communication stream, behind what the context's stream has done so far
        (the edge rows of the register update); the caller keeps enqueueing the interior rows meanwhile."""
        self.ctx.halo_prepare(instr)
        ops = self._ops(instr)
        self._pending = None
        if not ops:
            return
        if not self.on_device:
            self.ctx.sync()
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()
            return
        torch = self.torch
        ready, done = torch.cuda.Event(), torch.cuda.Event()
        ready.record(self.stream)
        self.comm.wait_event(ready)
        with torch.cuda.stream(self.comm):
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()
            done.record(self.comm)
        self._pending = done

    def finish(self):
        """Make the context's stream wait for the exchange begun by ``start``."""
        if self.on_device and self._pending is not None:
            self.stream.wait_event(self._pending)
        self._pending = None

    def reduce_eigmax(self):
        """all-reduce(MAX), in place on the device, of the three doubles behind ``astrea_eigmax_device``: the two
        per-axis wave speeds of operator 0 and the "non-finite seen" flag — so every rank takes the same dt and
        raises together where the reference raises.  Stream-ordered, no host synchronisation."""
        dist = self.dist
        if self.on_device:
            with self.torch.cuda.stream(self.stream):
                dist.all_reduce(self._eig, op=dist.ReduceOp.MAX)
        else:
            self.ctx.sync()
            dist.all_reduce(self._eig, op=dist.ReduceOp.MAX)

    def reduce_flags(self, ptr, count):
        """In-place maximum (logical OR) over the ranks of ``count`` int32 switches on the device, ordered on the context's
        stream: the grid-wide ``any()`` of the PPM authors 'c' / 'ph' (limiters.py:58,164)."""
        dist = self.dist
        flags = _tensor_at(ptr, count, self.on_device, self.device_index, int32=True)
        if self.on_device:
            with self.torch.cuda.stream(self.stream):
                dist.all_reduce(flags, op=dist.ReduceOp.MAX)
        else:
            self.ctx.sync()
            dist.all_reduce(flags, op=dist.ReduceOp.MAX)

    def reduce_keys(self, ptr, count):
        """In-place minimum over the ranks of ``count`` unsigned 64-bit keys on the device, ordered on the context's
        stream: the first non-zero entry of the Lax-Wendroff spectrum columns over the whole grid (solvers.py:79-88).
        torch has no uint64 reduction: flipping the top bit maps the unsigned order onto the signed one."""
        dist = self.dist
        keys = _tensor_at(ptr, count, self.on_device, self.device_index, int64=True)
        top = -(1 << 63)
        if self.on_device:
            with self.torch.cuda.stream(self.stream):
                keys.bitwise_xor_(top)
                dist.all_reduce(keys, op=dist.ReduceOp.MIN)
                keys.bitwise_xor_(top)
        else:
            self.ctx.sync()
            keys.bitwise_xor_(top)
            dist.all_reduce(keys, op=dist.ReduceOp.MIN)
            keys.bitwise_xor_(top)

    def global_eigmax(self):
        self.reduce_eigmax()
        return self.ctx.read_eigmax()      # synchronises; raises NonFiniteError on every rank if any rank saw one


class Simulation:
    """One run of the reference's time loop on the device(s).

    ``config`` / ``cells`` / ``subgrid`` / ``solver`` / ``timestep`` / ``cfl`` / ``gamma`` have the meaning of the
    reference's parameters.yml keys (static/.default.yml).  With ``world > 1`` this rank holds ``cells_x`` rows
    starting at ``rank * cells_x`` of a global ``(world * cells_x) x cells`` grid.

    ``overlap=True`` exchanges the ghost rows behind the register update that produces them (edge rows first, NCCL
    on a second high-priority stream).  It is bit-identical and tested, but measured slower than the in-order
    exchange at 2048^2 and 4096^2 per GPU (5.15 vs 4.91 ms and 16.91 vs 16.71 ms per step on 2 B200): the exchange
    costs ~60 us per stage in order, less than the extra launches and stream hand-offs of hiding it.  Off by default.
    """

    def __init__(self, config, cells, dimension, subgrid, solver, timestep, cfl=0.5, gamma=1.4, device=0, boundary=None,
                 rank=0, world=1, cells_x=None, grid=None, overlap=False, device_init=True, _lib=None, **geometry):
        self.config, self.cells, self.dimension = config.lower(), int(cells), int(dimension)
        prob = problem(self.config, self.cells, gamma)
        self.boundary = boundary or prob["boundary"]
        self.dx, self.t_end = prob["dx"], prob["t_end"]
        self.start_pos, self.end_pos = prob["start_pos"], prob["end_pos"]
        self.cfl, self.gamma = cfl, gamma
        self.rank, self.world = rank, world
        self.high_order = scheme_enum(subgrid) >= N.PPM
        self.magnetic_2d = self.config in MAGNETIC_2D
        nx = self.cells if cells_x is None else int(cells_x)
        if world > 1 and dimension != 2:
            raise ValueError("only 2D grids are decomposed")
        self.nx_local, self.nx_global, self.x_offset = nx, nx * world, rank * nx
        self.cfg = make_cfg(dimension=dimension, nx=nx, ny=self.cells, boundary=self.boundary, gamma=gamma, dx=self.dx, cfl=cfl,
                            subgrid=subgrid, solver=solver, timestep=timestep, magnetic_2d=self.magnetic_2d, device=device,
                            nx_global=self.nx_global, x_offset=self.x_offset, **geometry)
        self.ctx = N.Context(self.cfg, lib=_lib)
        self.stages = stages_of(self.cfg.integrator)
        self.on_device = self.ctx.lib.astrea_is_device_build() == 1
        self.exchange = SlabExchange(self.ctx, rank, world, self.boundary == "wrap", self.on_device, device) if world > 1 else None
        if self.exchange is not None and self.cfg.scheme == N.PPM and self.cfg.ppm_author != N.PPM_MC:
            self.ctx.set_flag_reducer(self.exchange.reduce_flags)
        if self.exchange is not None and self.cfg.solver == N.LW:
            self.ctx.set_key_reducer(self.exchange.reduce_keys)
        self.t, self.steps_done = 0.0, 0
        self._program = self.ctx.program()
        self._updates = self.ctx.updates()
        self._readers = self.ctx.halo_readers()    # instructions that read ghost rows (operators, refine_grid)
        self._halo_ready = False          # the ghost rows of the grid were already exchanged behind the last update
        self._snap_pool = None
        self.overlap = overlap
        # piecewise-constant problems are initialised on the device (astrea_init_piecewise): nothing crosses PCIe
        spec = piecewise_spec(self.config, self.cells, gamma) if (grid is None and dimension == 2 and device_init) else None
        if spec is not None:
            self.ctx.init_piecewise(spec, separable_profiles(self.config, self.cells, gamma))
        else:
            self.ctx.upload(self.initial_grid() if grid is None else grid)

    def initial_grid(self):
        """constructor.initialise(sim_variables, convert=True) for this rank's rows."""
        if self.dimension == 2 and (self.nx_local != self.cells or self.world > 1):
            return initial_slab(self.config, self.nx_local, self.cells, self.x_offset, self.nx_global, self.gamma, self.high_order)
        return initial_state(self.config, self.cells, self.dimension, self.gamma, self.high_order, boundary=self.boundary)

    def _walk(self, after_first_operator):
        """Run the step program on a slab.  The ghost rows an operator needs are exchanged while the register update
        that produces its input is still running: the update does its edge rows first (``astrea_run_update_part``),
        the exchange starts on a second stream, the interior rows follow on the context's stream."""
        ctx, ex, prog = self.ctx, self.exchange, self._program
        last = len(prog) - 1
        ready = {0} if self._halo_ready else set()       # operators whose ghost rows are already in place
        self._halo_ready = False
        for i, is_operator in enumerate(prog):
            if is_operator:
                if i not in ready:
                    ex.halo(i)
                ctx.run_instr(i, external_rows=True)
                if i == 0:
                    after_first_operator()
            elif self.overlap and self._updates[i] and (i == last or prog[i + 1]):
                ctx.run_update_part(i, 0)
                ex.start(0 if i == last else i + 1)       # the last update produces the grid the next step starts from
                ctx.run_update_part(i, 1)
                ex.finish()
                if i == last:
                    self._halo_ready = True
                else:
                    ready.add(i + 1)
            elif self._readers[i]:
                # refine_grid of constrained transport reads the ghost rows of the register it refines
                ex.halo(i)
                ctx.run_instr(i, external_rows=True)
            else:
                ctx.run_instr(i)
        ctx.finish_step()

    def step(self, t_stop=None, dt=None):
        """One pass of astrea.py:67-85.  Returns dt.  ``dt``: take this time step instead of cfl*min(dx/eigmax) (parity
        runs that replay the reference's own dt sequence); the wave speeds are still reduced and checked."""
        stop = self.t - 1.0 if t_stop is None else t_stop
        forced = dt
        if self.exchange is None and forced is not None:
            self.ctx.evolve_space(self.ctx.parity)
            self.ctx.evolve_time(forced)
            self.ctx.parity = self.ctx.parity ^ 1
            dt = forced
        elif self.exchange is None:
            dt = self.ctx.step(self.t, stop)
        else:
            box = {}

            def choose_dt():
                eig = self.exchange.global_eigmax()
                dt = self.cfl * min(self.dx / e for e in eig)
                if stop > self.t and self.t + dt >= stop:
                    dt = stop - self.t
                if forced is not None:
                    dt = forced
                self.ctx.set_dt(dt)
                box["dt"] = dt

            self._walk(choose_dt)
            dt = box["dt"]
        self.t += dt
        self.steps_done += 1
        return dt

    def step_async(self):
        """One pass of astrea.py:67-85 enqueued without any host synchronisation: dt is computed on the device
        (``astrea_dt_async``) and t advances there.  Call ``set_time`` first; read the clock with ``time()``."""
        if self.exchange is None:
            self.ctx.step_async()
        else:
            def device_dt():
                self.exchange.reduce_eigmax()
                self.ctx.dt_async()

            self._walk(device_dt)
        self.steps_done += 1

    def run_steps(self, nsteps):
        """``nsteps`` x ``step_async`` (one call into the library on a single GPU)."""
        if self.exchange is None:
            self.ctx.run_steps(nsteps)
            self.steps_done += nsteps
        else:
            for _ in range(nsteps):
                self.step_async()

    def upload(self, grid):
        self._halo_ready = False
        self.ctx.upload(grid)

    def save_state(self):
        self.ctx.save_state()

    def restore_state(self):
        self._halo_ready = False
        self.ctx.restore_state()

    def set_time(self, t=0.0, t_stop=None):
        self.t = t
        self.ctx.set_time(t, t - 1.0 if t_stop is None else t_stop)

    def time(self):
        """(t, steps, last dt) from the device clock; synchronises and raises where the reference would have."""
        if self.exchange is not None:
            self.exchange.reduce_eigmax()        # make the non-finite flag global before reading it
        self.t, steps, dt = self.ctx.get_time()
        return self.t, steps, dt

    def check_finite(self):
        """Raise NonFiniteError (a LinAlgError) on every rank if any Runge-Kutta stage so far saw a non-finite wave
        speed — the reference raises inside evolve_time at that stage (fv.py:158)."""
        if self.exchange is not None:
            self.exchange.global_eigmax()
        else:
            self.ctx.read_eigmax()

    def run(self, nsteps):
        return [self.step() for _ in range(nsteps)]

    def state(self, primitive=False):
        """This rank's rows as the reference's ndarray (conservative averages, or the astrea.py:47 primitive snapshot)."""
        if primitive and self.exchange is not None:
            self.exchange.halo(0)                  # the 4th-order conversion reads one ghost row
            self._halo_ready = False
            return self.ctx.download(primitive=2)
        return self.ctx.download(primitive=primitive)

    def snapshot(self):
        """What astrea.py:47-50 stores per step: the primitive grid transposed by ``ortho_axis`` ((y, x, 8) in 2D).
        Conversion and transpose run on the device; this call waits for the copy (``snapshot_async`` does not)."""
        out, ticket = self.snapshot_async()
        self.ctx.snapshot_wait(ticket)
        return out

    def snapshot_async(self, out=None):
        """Start a snapshot and return ``(array, ticket)`` at once: the device converts and transposes, a second stream
        copies into page-locked host memory, and the steps enqueued next run meanwhile.  The array is complete after
        ``ctx.snapshot_wait(ticket)`` — the moment the reference would hand it to h5py (astrea.py:48-50)."""
        if out is None:
            if self._snap_pool is None:
                self._snap_pool = N.PinnedPool(self.ctx.lib, self.cfg.device)
            shape = (self.cells, self.nx_local, 8) if self.dimension == 2 else (self.cells, 8)
            out = self._snap_pool.empty(shape)
        external = False
        if self.exchange is not None:
            self.exchange.halo(0)                  # the 4th-order conversion reads one ghost row
            self._halo_ready = False
            external = True
        return out, self.ctx.snapshot_begin(out, external_rows=external)

    def diagnostics(self):
        """Conservation totals (times the box volume, functions/analytic.py:66-77) and total variation (:48-62) of the
        current grid, reduced on the device; summed over the ranks of a decomposed run (the total variation then
        misses the differences across slab seams)."""
        if self.exchange is not None:
            self.exchange.halo(0)
            self._halo_ready = False
            tot, tv = self.ctx.diagnostics(external_rows=True)
            torch = self.exchange.torch
            where = torch.device("cuda", self.exchange.device_index) if self.on_device else torch.device("cpu")
            t = torch.tensor(np.concatenate([tot, tv]), dtype=torch.float64, device=where)
            self.exchange.dist.all_reduce(t)
            tot, tv = t[:8].cpu().numpy(), t[8:].cpu().numpy()
        else:
            tot, tv = self.ctx.diagnostics()
        box = abs(self.end_pos - self.start_pos) ** self.dimension
        return tot * box, tv

    def solution_error(self, norm=1):
        """analytic.calculate_solution_error(grid, sim_variables, norm) (functions/analytic.py:24-44) of the current grid
        against the initial state, reduced on the device: 10 values (8 primitives, E_tot / rho, E_int)."""
        from .initial import theoretical_primitives
        theo = theoretical_primitives(self.config, self.cells, self.dimension, self.gamma, self.boundary)
        if self.exchange is not None:
            theo = theo[self.x_offset:self.x_offset + self.nx_local] if self.nx_global == self.cells else None
            if theo is None:
                raise ValueError("solution_error: the theoretical state exists for the reference's square grids only")
            self.exchange.halo(0)
            self._halo_ready = False
            part = self.ctx.solution_error(theo, norm, external_rows=True)
            torch = self.exchange.torch
            where = torch.device("cuda", self.exchange.device_index) if self.on_device else torch.device("cpu")
            t = torch.tensor(part, dtype=torch.float64, device=where)
            self.exchange.dist.all_reduce(t, op=self.exchange.dist.ReduceOp.MAX if norm > 10 else self.exchange.dist.ReduceOp.SUM)
            part = t.cpu().numpy()
        else:
            part = self.ctx.solution_error(theo, norm)
        if norm > 10:
            return part
        factor = 1 / (self.cells ** self.dimension)
        return factor * part if norm <= 0 else (factor * part) ** (1 / norm)

    def sync(self):
        self.ctx.sync()

    def close(self):
        self.ctx.close()
        if self._snap_pool is not None:
            self._snap_pool.close()
