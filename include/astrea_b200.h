/* astrea_b200 — C ABI of the B200-native per-timestep finite-volume update of mervyzr/astrea.
 *
 * The reference has no plugin or FFI interface: the seam is the Python call pair
 *     fluxes = evolvers.evolve_space(grid, sim_variables)            astrea.py:67   (num_methods/evolvers.py:12-34)
 *     grid   = evolvers.evolve_time(grid, fluxes, dt, sim_variables) astrea.py:81   (num_methods/evolvers.py:38-206)
 * plus the read of fluxes[axes]['eigmax'] at astrea.py:70-71 and the primitive snapshot at astrea.py:47.
 * Every entry point below names the reference line(s) it stands in for.  All functions take plain pointers
 * and sizes; host buffers stay owned by the caller, device state is owned by the context.  One host thread
 * per context; all work of a context is ordered on one CUDA stream.
 *
 * Return value: 0 on success, a negative ASTREA_E_* code on failure (astrea_last_error() gives the text).
 */
#ifndef ASTREA_B200_H
#define ASTREA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct astrea_ctx astrea_ctx;

enum { ASTREA_PCM = 0, ASTREA_PLM = 1, ASTREA_PPM = 2, ASTREA_WENO3 = 3, ASTREA_WENO5 = 4, ASTREA_WENO7 = 5 };   /* sim_variables.subgrid, evolvers.py:14-21 */
enum { ASTREA_PPM_MC = 0, ASTREA_PPM_COLELLA = 1, ASTREA_PPM_PH = 2 };   /* ppm.run(author=...): evolvers.py:17 passes 'mc'; 'c' / 'ph' = limiters.py:53-78,144-201 */
enum { ASTREA_MINMOD = 0, ASTREA_VANLEER = 1, ASTREA_OSPRE = 2, ASTREA_VANALBADA = 3, ASTREA_KOREN = 4, ASTREA_SUPERBEE = 5 }; /* limiters.py:10-49 */
enum { ASTREA_LLF = 0, ASTREA_LW = 1, ASTREA_HLLC = 2, ASTREA_HLLD = 3 };                                         /* sim_variables.solver, solvers.py:13-31;
                                                    ASTREA_LW: grids without v_z / B, one GPU (solvers.py:84 sorts LAPACK's eigenvalue slots grid-wide) */
enum { ASTREA_EULER = 0, ASTREA_RK4 = 1, ASTREA_SSPRK22 = 2, ASTREA_SSPRK33 = 3, ASTREA_SSPRK43 = 4,
       ASTREA_SSPRK53 = 5, ASTREA_SSPRK54 = 6, ASTREA_SSPRK104 = 7 };                                          /* sim_variables.timestep, evolvers.py:79-206 */
enum { ASTREA_EDGE = 0, ASTREA_WRAP = 1 };                                                                       /* sim_variables.boundary (np.pad mode), fv.py:57-61 */

enum {
    ASTREA_OK = 0,
    ASTREA_E_ARG = -1,        /* bad argument / unsupported selector combination */
    ASTREA_E_CUDA = -2,       /* CUDA runtime error */
    ASTREA_E_NONFINITE = -3,  /* non-finite wave speed: where the reference raises LinAlgError out of fv.py:158 (astrea.py:211-217) */
    ASTREA_E_STATE = -4       /* call order violated (e.g. evolve_time without evolve_space) */
};

/* The subset of the reference's sim_variables namedtuple that the hot path reads
 * (functions/generic.py:159-286, static/tests.py:317-328). */
typedef struct astrea_cfg {
    int32_t dimension;     /* 1 | 2 */
    int32_t boundary;      /* ASTREA_EDGE | ASTREA_WRAP */
    int64_t nx;            /* local cells along x (1D: the only axis) */
    int64_t ny;            /* local cells along y (1D: 1) */
    double gamma;
    double dx;             /* cell width, dx == dy (tests.py:326-327) */
    double cfl;
    int32_t scheme;        /* ASTREA_PCM .. ASTREA_WENO7 */
    int32_t ppm_author;    /* ASTREA_PPM_MC (what the reference's evolve_space uses) | ASTREA_PPM_COLELLA | ASTREA_PPM_PH */
    int32_t limiter;       /* slope limiter of PLM */
    int32_t solver;        /* ASTREA_LLF .. ASTREA_HLLD */
    int32_t low_mach;      /* solvers.py:92 low_mach switch of HLLC */
    int32_t integrator;    /* ASTREA_EULER .. ASTREA_SSPRK104 */
    int32_t magnetic_2d;   /* constrained-transport update on (generic.py:243) */
    int32_t device;        /* CUDA device ordinal */
    /* slab decomposition along x; single process: nx_global = nx, x_offset = 0 */
    int64_t nx_global;
    int64_t x_offset;
    int32_t threads_2d;    /* 0 = default (128); threads per block of the 2D flux stage (a multiple of 32, <= 128) */
    int32_t segment_2d;    /* 0 = default (64; 128 from 4096 cells per sweep); cells a thread of the 2D reconstruction stage marches along the sweep */
    int32_t tile_1d;       /* 0 = default; cells per block of the 1D sweep kernel */
    int32_t flags;         /* bit 0: keep the general 8-variable kernels even when the grid has no v_z / B (testing);
                              bit 1: never replay astrea_step_async as a CUDA graph (small grids do by default);
                              bit 2: reconstruction march with register prefetch instead of bulk asynchronous copies (A/B runs);
                              bit 3 / bit 4: flux stage with warp-wide / block-wide rows of transverse points whatever the
                              grid width (default: block-wide from 4096 columns);
                              bit 6: evaluate the interface wave speeds in every operator of a step (A/B runs).  Default:
                              the operators after the first, whose speeds the reference computes but uses only to raise on
                              NaN / Inf (fv.py:158), evaluate them just for states that could give a non-finite one */
} astrea_cfg;

/* sim_variables -> device context.  Stands in for the namedtuple built at astrea.py:132-133. */
astrea_ctx* astrea_create(const astrea_cfg* cfg);
void astrea_destroy(astrea_ctx* ctx);
const char* astrea_last_error(const astrea_ctx* ctx);   /* ctx may be NULL: error of the last failed astrea_create */

/* grid (conservative cell averages, C-order (nx[,ny],8) float64 host array; constructor.py:11-109 output,
 * the `grid` argument of astrea.py:67) -> device. */
int astrea_upload(astrea_ctx* ctx, const double* grid_aos);
/* device -> host, same layout.  as_primitive != 0 applies sim_variables.convert_conservative first, i.e. what
 * astrea.py:47 snapshots (without its transpose); as_primitive == 2: the caller (a slab host) has already exchanged
 * the ghost rows of instruction 0, which the 4th-order conversion reads. */
int astrea_download(astrea_ctx* ctx, double* grid_aos, int as_primitive);

/* The per-step snapshot of astrea.py:47-50 off the critical path: sim_variables.convert_conservative(grid) transposed by
 * ortho_axis — host layout (ny, nx, 8) in 2D, (nx, 8) in 1D, the array the reference hands to h5py.  astrea_snapshot_begin
 * enqueues the primitive conversion and the transposing pack on the context's stream, then the device-to-host copy on a
 * second stream, and returns a ticket (>= 0) without waiting: the next steps may be enqueued right away and run while the
 * snapshot travels.  ``host_dst`` must stay valid until astrea_snapshot_wait(ticket) has returned; give it page-locked
 * memory (astrea_host_alloc) for the copy to overlap.  One device staging buffer of the grid's size is allocated on first
 * use; the next snapshot's conversion waits on the device for the previous copy.  external_rows as in astrea_download
 * (as_primitive == 2).  Negative return: ASTREA_E_*. */
int64_t astrea_snapshot_begin(astrea_ctx* ctx, double* host_dst, int external_rows);
int astrea_snapshot_wait(astrea_ctx* ctx, int64_t ticket);

/* evolvers.evolve_space(grid, sim_variables) for the uploaded grid (astrea.py:67).  step_parity = number of
 * permutation reversals so far mod 2 (astrea.py:85).  eigmax[a] receives fluxes[axes_a]['eigmax'] for
 * sweep axis a = 0..dimension-1 (astrea.py:70).  Returns ASTREA_E_NONFINITE where fv.py:158 would raise. */
int astrea_evolve_space(astrea_ctx* ctx, int step_parity, double* eigmax);
/* evolvers.evolve_time(grid, fluxes, dt, sim_variables) (astrea.py:81): all remaining Runge-Kutta stages.  The new
 * grid replaces the context's state. */
int astrea_evolve_time(astrea_ctx* ctx, double dt);
/* One pass of the loop body astrea.py:67-85: evolve_space, dt = cfl*min(dx/eigmax) (:70-71), clip so that
 * t + dt does not pass t_stop (:74-75; pass t_stop <= t to disable), evolve_time, flip the parity. */
int astrea_step(astrea_ctx* ctx, double t, double t_stop, double* dt_out);
/* The same loop body without a host round trip: the time step is computed on the device from the wave speeds
 * (astrea.py:70-71), clipped against t_stop (:74-75), and t / the step count advance on the device, so steps can be
 * enqueued back to back.  astrea_set_time sets t and t_stop (t_stop <= t: no clipping) and resets the step count;
 * astrea_get_time synchronises, returns t, the number of steps and the last dt, and reports ASTREA_E_NONFINITE if
 * any operator since the last check saw a non-finite wave speed (the reference would have raised at that step);
 * astrea_dt_history returns the dt of the last n steps (n <= 1024), oldest first. */
int astrea_set_time(astrea_ctx* ctx, double t, double t_stop);
int astrea_step_async(astrea_ctx* ctx);
/* nsteps passes of the loop body (astrea.py:67-85) enqueued at once: nsteps calls of astrea_step_async (small grids
 * replay a CUDA graph per step).  The t_stop clip of astrea_set_time applies as usual. */
int astrea_run_steps(astrea_ctx* ctx, int64_t nsteps);
int astrea_get_time(astrea_ctx* ctx, double* t, int64_t* steps, double* last_dt);
int astrea_dt_history(astrea_ctx* ctx, double* out, int n);
/* multi-GPU hosts: after instruction 0 and the all-reduce(MAX) of the three doubles at astrea_eigmax_device
 * (two wave speeds + the non-finite flag), enqueue the device-side dt computation */
int astrea_dt_async(astrea_ctx* ctx);
int astrea_get_parity(const astrea_ctx* ctx);
int astrea_set_parity(astrea_ctx* ctx, int step_parity);

/* Reductions of the current grid on the device, the quantities functions/analytic.py computes from the HDF5 snapshots:
 * totals[8] = sum over the (local) cells of every conservative variable (calculate_conservation, :66-77, before its
 * box-width factor); total_variation[8] = sum of |np.diff along every axis in turn| of the primitive snapshot
 * (calculate_TV, :48-62).  Deterministic (fixed reduction order).  external_rows as in astrea_run_instr. */
int astrea_diagnostics(astrea_ctx* ctx, double* totals, double* total_variation, int external_rows);

/* functions/analytic.py:24-44 calculate_solution_error, reduced on the device: error[10] = over the (local) cells, per
 * channel (the 8 primitive variables, E_tot / rho, E_int: analytic.py:33-37), the maximum (norm > 10), the sum (norm <= 0)
 * or the sum of the norm-th powers of |w_num - w_theo|, where w_num is the primitive snapshot of the current grid
 * (astrea.py:47) and ``w_theo_aos`` the caller's theoretical state (constructor.initialise(sim_variables), host array of
 * the grid's shape).  The caller applies the normalising factor 1 / cells^dimension and the 1 / norm root
 * (analytic.py:39-44), after summing over ranks on a decomposed grid.  Deterministic reduction order. */
int astrea_solution_error(astrea_ctx* ctx, const double* w_theo_aos, double norm, double* error, int external_rows);

/* schemes/ppm.py:111-170 at function level: the two pieces ppm.run(dissipate=True) adds to the McCorquodale-Colella
 * reconstruction.  The reference cannot take a time step with dissipate=True (ppm.py:67 raises a broadcast error), but
 * both functions run on their own and are reproduced bit for bit.  ``ws_aos``: primitive cell averages in the sweep
 * frame, host array of the context's grid shape whose axis 0 is the sweep direction (the reference passes
 * convert_conservative(grid.transpose(axes))); ``axis``: the permutation key, i.e. the velocity component
 * w[..., axis + 1].  Boundary mode = the context's.
 *   astrea_ppm_flattener  = ppm.apply_flattener(wS, axis, boundary, slope_determinants) (ppm.py:111-134): chi, host array
 *                           (nx[,ny]) — the reference returns it repeated over the 8 variables.  slope_determinants =
 *                           {delta, z0, z1}, NULL for the reference's defaults {.33, .75, .85}.
 *   astrea_ppm_viscosity  = ppm.apply_artificial_viscosity(wS, axis, sim_variables, viscosity_determinants)
 *                           (ppm.py:138-170) read cell by cell: mu (nx, 8).  1D only: the reference's 2D branch raises
 *                           at ppm.py:154-156, and so does this call (ASTREA_E_ARG).  viscosity_determinants = {alpha, beta},
 *                           NULL for {.3, .3}. */
int astrea_ppm_flattener(astrea_ctx* ctx, const double* ws_aos, int axis, const double* slope_determinants, double* chi);
int astrea_ppm_viscosity(astrea_ctx* ctx, const double* ws_aos, int axis, const double* viscosity_determinants, double* mu_aos);

/* magnetic_2d only: evolve_time overwrites the in-plane B of the caller's grid with face averages before the
 * stages (evolvers.py:73-76).  Copies those two components (host array (nx,ny,2): Bx, By). */
int astrea_download_face_field(astrea_ctx* ctx, double* bxy_aos);

/* ---- Step program: the spatial-operator evaluations and Runge-Kutta register updates of one time step, in the
 * order evolvers.py:70-206 performs them.  Instruction 0 is always the operator on the current grid (what
 * evolve_space does); astrea_evolve_time runs instructions 1..n-1.  A multi-GPU host (one process per GPU) drives
 * the instructions itself so that it can exchange the ghost rows of the register an operator is about to read:
 *     for i in range(astrea_program_length(ctx)):
 *         if astrea_instr_is_operator(ctx, i): <NCCL send/recv on astrea_halo_ptrs(ctx, i, ...)>
 *         astrea_run_instr(ctx, i, external_rows)
 * Each halo block is ghost_rows x 8 variables x col_pitch doubles, contiguous (the [row][var][col] layout). */
/* The PPM authors 'c' / 'ph' (ASTREA_PPM_COLELLA / ASTREA_PPM_PH) switch their limiters on ``mask.any()`` over the WHOLE
 * grid (limiters.py:58,164).  On a decomposed grid the library evaluates the masks of its slab into four
 * int32 switches on the device and then calls ``fn(user, device_ptr, 4)``, which must replace them, in place and ordered
 * on the context's stream, by their maximum (= logical OR) over all ranks — e.g. one ncclAllReduce(ncclMax) — and
 * return 0.  Called twice per sweep of such an operator, from inside astrea_run_instr.  Not needed on a whole grid. */
typedef int (*astrea_reduce_fn)(void* user, void* device_int32, int count);
int astrea_set_flag_reducer(astrea_ctx* ctx, astrea_reduce_fn fn, void* user);
/* Lax-Wendroff (solvers.py:79-88) takes column 1 of ``np.unique(characteristics, axis=-1)``: a lexicographic sort of the
 * spectrum columns over the WHOLE padded array.  For states without v_z / B the pick is fixed by the first
 * non-zero entry of each of three columns; a slab finds its own (uint64 keys: 2 x position in the whole array + sign
 * bit, all ones = none) and then calls ``fn(user, device_ptr, 4)``, which must replace the ``count`` UNSIGNED 64-bit
 * values, in place and ordered on the context's stream, by their minimum over all ranks — e.g. one
 * ncclAllReduce(ncclUint64, ncclMin) — and return 0.  Called once per sweep of such an operator. */
int astrea_set_key_reducer(astrea_ctx* ctx, astrea_reduce_fn fn, void* user);
int astrea_program_length(const astrea_ctx* ctx);
int astrea_instr_is_operator(const astrea_ctx* ctx, int instr);
/* 1 if instruction `instr` reads ghost rows of a register (a spatial operator, or the inverse reconstruction of
 * constrained transport): a slab host exchanges them first (astrea_halo_prepare / astrea_halo_ptrs) and passes
 * external_rows = 1 to astrea_run_instr */
int astrea_instr_needs_halo(const astrea_ctx* ctx, int instr);
/* 1 if instruction `instr` is a Runge-Kutta register update (not an operator, not a constrained-transport special) */
int astrea_instr_is_update(const astrea_ctx* ctx, int instr);
/* A register update in two parts, so that a slab host can overlap the halo exchange of the register it produces
 * with most of the update: part 0 = the first and last 32 rows (what the neighbours' ghost rows are made of),
 * part 1 = the rows in between (completes the instruction).  astrea_run_instr does both at once. */
int astrea_run_update_part(astrea_ctx* ctx, int instr, int part);
int astrea_set_dt(astrea_ctx* ctx, double dt);                 /* dt of astrea.py:70-78, read by the register updates */
/* external_rows != 0: the caller has filled the x ghost rows (neighbour ranks); only ghost columns are filled here */
int astrea_run_instr(astrea_ctx* ctx, int instr, int external_rows);
int astrea_finish_step(astrea_ctx* ctx);                       /* adopt the last register as the grid, flip the parity (astrea.py:81,85) */
int astrea_halo_info(const astrea_ctx* ctx, int64_t* ghost_rows, int64_t* doubles_per_block);
int astrea_halo_ptrs(astrea_ctx* ctx, int instr, double** send_lo, double** send_hi, double** recv_lo, double** recv_hi);
/* fill the ghost columns of the interior rows of the register instruction `instr` reads, so that the rows handed to
 * the neighbours carry their corner cells; call before the exchange */
int astrea_halo_prepare(astrea_ctx* ctx, int instr);
/* device address of three doubles: eigmax[2] as written by the last operator 0 and the non-finite flag (0.0 / 1.0),
 * for one all-reduce(MAX) across ranks */
int astrea_eigmax_device(astrea_ctx* ctx, double** eigmax_dev);
int astrea_read_eigmax(astrea_ctx* ctx, double* eigmax);        /* sync + copy to host, ASTREA_E_NONFINITE as above */
int astrea_sync(astrea_ctx* ctx);
/* cudaStream_t of the context as an integer (for torch.cuda.ExternalStream); 0 in the host-simulated build */
uint64_t astrea_stream_handle(const astrea_ctx* ctx);

/* Device-side copy of the current grid (and step parity) and its restoration: lets a host loop return to a known
 * state without a host round trip (the reference re-runs constructor.initialise, astrea.py:35). */
int astrea_save_state(astrea_ctx* ctx);
int astrea_restore_state(astrea_ctx* ctx);

/* Per-launch timing with CUDA events on the context's stream, summed per kernel class:
 * 0 = flux stage (2D) / fused sweep (1D), 1 = transpose, 2 = rate assembly + Runge-Kutta update,
 * 3 = halo fill / pack / unpack, 4 = primitive stage, 5 = reconstruction stage.
 * astrea_profile_read synchronises, returns the sums since the last read (arrays of 6) and clears them. */
int astrea_profile(astrea_ctx* ctx, int enable);
int astrea_profile_read(astrea_ctx* ctx, double* ms_by_class, int64_t* launches_by_class);

/* Measurement aid: the fp64 FMA rate (TFLOP/s, FMA = 2) this GPU sustains on independent DFMA chains, so that the
 * roofline report can say how far the fp64-bound stages are from the fp64 peak.  0 in the host-simulated build. */
int astrea_fp64_probe(astrea_ctx* ctx, double* tflops);

/* Initial conditions on the device for the problems of static/tests.py whose pointwise
 * primitive state is piecewise constant — the Lax-Liu family (constructor.py:57-60), the discs of Sedov / blast /
 * rotor (:33-34, :69-73) and the default half-plane (:75): replaces constructor.initialise(sim_variables,
 * convert=True) (constructor.py:11-109) + astrea_upload, so that a rank's slab never crosses PCIe (4.3 GB at 8192^2).
 * The grid starts as ``background`` (tests.py ``initial_right``) and the regions are painted over it in order, like the
 * reference's successive ``grid[np.where(...)] = state`` assignments; the conversion to conservative cell averages
 * (generic.py:250-255, fv.py:67-85, 126-143) follows the scheme of the context.  Coordinates: cell i sits at
 * (i + 1) * step + start, the values np.linspace(lo - half, hi + half, cells + 2) produces (constructor.py:13-16).
 * On a slab, row r of the context is row (r + x_offset) mod cells of the square problem (periodic tiling along x).
 * Bit-identical to the upload of the reference's array (tests/test_hostsim_parity.py, tests/test_gpu_parity.py). */
#define ASTREA_MAX_REGIONS 8
enum astrea_region_kind {
    ASTREA_REGION_X_LT = 0,          /* x <  a                      constructor.py:75            */
    ASTREA_REGION_X_LE = 1,          /* x <= a                      :58                          */
    ASTREA_REGION_Y_LE = 2,          /* y <= a                      :43, :63                     */
    ASTREA_REGION_X_LE_Y_GE = 3,     /* x <= a and y >= a           :59                          */
    ASTREA_REGION_X_GT_Y_GE = 4,     /* x >  a and y >= a           :60                          */
    ASTREA_REGION_DISC_LE = 5        /* (x-a)^2 + (y-a)^2 <= b      :33-34, :70                  */
};
typedef struct astrea_region {
    int32_t kind;
    int32_t reserved;
    double a, b;
    double state[8];                 /* primitive [rho, vx, vy, vz, P, Bx, By, Bz] */
} astrea_region;
typedef struct astrea_init_spec {
    int64_t cells;                   /* cells per side of the square problem (= ny of the context) */
    double start, step;              /* first point and spacing of np.linspace(lo - half, hi + half, cells + 2) */
    double background[8];
    int32_t nregions;
    int32_t reserved;
    astrea_region regions[ASTREA_MAX_REGIONS];
} astrea_init_spec;
int astrea_init_piecewise(astrea_ctx* ctx, const astrea_init_spec* spec);
/* The same with up to ASTREA_MAX_PROFILES separable profiles painted over the regions, for the problems whose
 * perturbation depends on x or on y only — Kelvin-Helmholtz (constructor.py:42-44: v_y = ampl*sin(freq*pi*x/(hi-lo)))
 * and Orszag-Tang (:62-67: v_x, B_x from sin(2*pi*y); v_y, B_y from sin(2*pi*x), sin(4*pi*x)), BASELINE configs 3 and 4.
 * ``values`` is the profile on the ``cells`` cell centres, evaluated by the caller with the reference's own numpy
 * expression on the 1-D coordinate array (so the transcendental function is bit for bit the reference's); primitive
 * variable ``variable`` of point (i, j) becomes values[along == 0 ? i : j].  Uploads cells doubles per profile instead
 * of the 8 * cells^2 of the finished grid (1 GB at 4096^2). */
#define ASTREA_MAX_PROFILES 4
typedef struct astrea_init_profile {
    int32_t variable;                /* 0..7: rho, vx, vy, vz, P, Bx, By, Bz */
    int32_t along;                   /* 0: function of x (row index), 1: function of y (column index) */
    const double* values;            /* host array, spec->cells entries */
} astrea_init_profile;
int astrea_init_profiles(astrea_ctx* ctx, const astrea_init_spec* spec, int nprofiles, const astrea_init_profile* profiles);

/* Self-check of the device arithmetic (no reference counterpart; the reference's divisions and square roots are
 * numpy's IEEE ones, fv.py:19-20,37-38).  The kernels evaluate every division and square root with a branch-free
 * fused-multiply-add sequence ("Fast", csrc/common.cuh) that is IEEE-exact for ordinary operands and hand the
 * warp / block to the compiler's IEEE routines otherwise.  This call runs Fast against those routines on about
 * ``samples`` generated operand pairs (6 operations each: both policies of the division, two square roots): counts[0] = operations Fast accepted, counts[1] = accepted
 * operations whose result differs in any bit from IEEE (must be 0), counts[2] = operations Fast declined. */
int astrea_arith_check(astrea_ctx* ctx, int64_t samples, uint64_t seed, uint64_t* counts);

/* Page-locked host memory for the arrays that cross the seam every step.  evolvers.evolve_time returns a NEW array
 * (astrea.py:81 rebinds `grid` to it) and the caller hands that array to the next evolve_space (astrea.py:67): when the
 * drop-in allocates its return value here, every transfer after the first upload is a direct DMA instead of a staged
 * pageable copy.  astrea_host_alloc returns NULL on failure; astrea_host_free(NULL) is a no-op.  (Host-simulated
 * build: malloc / free.) */
void* astrea_host_alloc(int device, uint64_t bytes);
void astrea_host_free(void* ptr);

/* Number of kernels this library launched on the context's stream since creation (bench.py "gpu_launches"). */
int64_t astrea_launch_count(const astrea_ctx* ctx);
/* 1 when built by nvcc for sm_100a, 0 for the host-simulated test build. */
int astrea_is_device_build(void);

#ifdef __cplusplus
}
#endif
#endif
