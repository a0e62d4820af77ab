// C ABI of astrea_b200 (include/astrea_b200.h): context, register file, step program, launches.
//
// HBM layout: every state register is one ghost-padded plane [row][var][col] of fp64 with
// GHOST cells on every side (2D: row = x, col = y; 1D: one row).  A spatial-operator evaluation is
//     halo fill -> PrimBothStage (primitive averages of both sweep frames)
//               -> x sweep: ReconStage -> FluxStage -> F_x    (x frame)
//               -> y sweep: ReconStage -> FluxStage -> F_y^T  (y frame, transposed planes)
//               [-> constrained transport: corner E_z]
// and a Runge-Kutta register update assembles L = -(dF_x + dF_y)/dx on the fly (UpdateKernel).  The order of
// operator evaluations and register updates of a time step is a small "step program" built once per context
// from the integrator (evolvers.py:79-206).
#include "../../include/astrea_b200.h"
#include "aux_kernels.cuh"
#include "ct_kernels.cuh"
#include "dispatch.cuh"

#include <algorithm>
#include <cmath>
#include <string>
#include <thread>
#include <vector>

namespace astrea {

}  // namespace astrea

using namespace astrea;

namespace {

constexpr int DT_HISTORY = 1024;
// reach of the transverse PPM of constrained transport beyond a slab: cells 0 .. n+1 with a stencil of -3 .. +4
constexpr int CT_LO = 3, CT_HI = 6;
constexpr int64_t UPDATE_EDGE = 32;              // rows of a register update done first when a slab host overlaps its halo exchange
constexpr int64_t GRAPH_MAX_CELLS = 1 << 18;     // astrea_step_async replays a CUDA graph up to 512^2 cells

struct Reg {
    double* mem = nullptr;
    Plane plane{};
};

// one term of a register update: a state register or a rate buffer, with its literal coefficient
struct Term { int is_rate; int index; double coef; };
enum { SP_NONE = 0, SP_FACE_FIELD = 1, SP_REFINE = 2 };
struct Instr {
    int is_operator;            // 1: rate[rate_out] = L(reg[src]);   0: reg[out] = combination of terms (or a special)
    int src, rate_out;
    int out, bracket;
    double scale;
    std::vector<Term> terms;
    int defer_rate = 0;         // operator: do not assemble L, the next register update does it on the fly
    int fused_rate = -1;        // update: index of the rate buffer assembled on the fly from the flux planes (-1: none)
    int store_rate = 0;         // update with fused_rate: also store that rate (a later formula re-uses it)
    int refine = 0;             // magnetic_2d: refine_grid (mag_field.inverse_reconstruct) wraps this update
    int special = SP_NONE;      // SP_FACE_FIELD: evolvers.py:73-76;  SP_REFINE: refine_grid applied to reg[out]
};

thread_local std::string g_create_error;

}  // namespace

struct astrea_ctx {
    astrea_cfg cfg{};
    Stream st{};
    std::string err;
    int parity = 0;
    int64_t launches = 0;
    int64_t nrow = 0, ncol = 0;       // interior rows / columns of a register plane
    int ghost_r = 0;                  // ghost rows (0 in 1D)
    size_t plane_doubles = 0;
    std::vector<Reg> regs, rates;
    Reg qT, d0, d1t;                  // transposed input of the y sweep; interface fluxes of the x / y sweep (1D: flux difference)
    Reg ws, ws2, wp, wm;              // 2D scratch: primitive averages (x frame, y frame), interface states of the sweep in flight
    Reg wfx, wfy, ct0;                // magnetic_2d: face states of the two sweeps (each in its frame), one more scratch
    double* emf = nullptr;            // magnetic_2d: corner electric field [nrow (+1)][ncol]
    int64_t emf_rows = 0;
    // hydro specialisation (physics.cuh): the uploaded grid has no v_z / B, so only [rho, m_x, m_y, E] are processed
    bool hydro = false, saved_hydro = false, saved_field_free = false;
    int* mhd_flag = nullptr;          // device: set by the upload when a v_z / B component is non-zero
    int* ppm_flags = nullptr;         // device [4]: grid-wide switches of the PPM authors 'c' / 'ph' (recon.cuh)
    unsigned long long* lw_keys = nullptr;   // device [4]: Lax-Wendroff column search (FluxStage / Sweep1D)
    bool field_free = false;          // the uploaded grid has no v_z / B
    VarList vars() const { return hydro ? hydro_vars() : all_vars(); }
    // the ghost rows beyond the low / high end of this slab hold genuine neighbour data (exchanged), i.e. the end is
    // not a physical 'edge' boundary and the grid is decomposed
    bool slab() const { return cfg.dimension == 2 && cfg.nx != cfg.nx_global; }
    bool ext_lo() const { return slab() && !(cfg.boundary == BC_EDGE && cfg.x_offset == 0); }
    bool ext_hi() const { return slab() && !(cfg.boundary == BC_EDGE && cfg.x_offset + cfg.nx == cfg.nx_global); }
    unsigned long long* eig_bits = nullptr;   // [2] bit patterns of the per-axis max wave speed (operator 0)
    unsigned long long* eig_scratch = nullptr; // [2] same for the later stages (checked for finiteness only)
    unsigned long long* flag = nullptr;   // eig_bits[2]: 1.0 once a non-finite wave speed was seen (sticky until read)
    double* clock = nullptr;          // device: t, t_stop, steps, last dt (ClockKernel)
    double* dt_history = nullptr;     // device: dt of the last DT_HISTORY steps
    double* dt_dev = nullptr;
    std::vector<Instr> prog;
    int grid_reg = 0;                 // register holding the current grid
    int final_reg = 0;                // register the last instruction writes
    int next_instr = 0;               // 0: nothing run for this step yet
    int tile1d = 0, threads1d = 0;
    bool stream_owned = false;
    // snapshot path (astrea.py:47-50): device staging buffer, copy stream, events of the tickets in flight
    static constexpr int SNAP_RING = 4;
    double* snap_dev = nullptr;
    int64_t snap_tickets = 0;
#ifdef ASTREA_DEVICE_BUILD
    cudaStream_t copy_st = nullptr;
    cudaEvent_t snap_ready = nullptr, snap_done[SNAP_RING] = {nullptr, nullptr, nullptr, nullptr};
#endif
    // cross-rank OR of the grid-wide switches of the PPM authors 'c' / 'ph' on a decomposed grid (astrea_set_flag_reducer)
    astrea_reduce_fn reduce_fn = nullptr;
    void* reduce_user = nullptr;
    // cross-rank minimum of the Lax-Wendroff search keys on a decomposed grid (astrea_set_key_reducer)
    astrea_reduce_fn key_reduce_fn = nullptr;
    void* key_reduce_user = nullptr;
    // transfers between the device and a pageable host array (the first grid of a run, np.empty destinations): LANES
    // host threads, each with its own stream and two page-locked bounce buffers (staged_copy)
    static constexpr int LANES = 12;
#ifdef ASTREA_DEVICE_BUILD
    void* bounce[LANES][2] = {};
    cudaStream_t lane_st[LANES] = {};
    cudaEvent_t bounce_done[LANES][2] = {};
    cudaEvent_t lane_begin = nullptr;
#endif
    Reg saved;                        // astrea_save_state copy of the grid
    int saved_parity = 0;
    // optional per-launch timing (astrea_profile): event pairs per kernel class
    int profiling = 0;
#ifdef ASTREA_DEVICE_BUILD
    struct Span { int cls; cudaEvent_t a, b; };
    std::vector<Span> spans;
    // astrea_step_async on small grids replays a captured CUDA graph of the whole step: [step parity][hydro]
    cudaGraphExec_t step_graph[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    int eager_steps[2][2] = {{0, 0}, {0, 0}};
    int64_t graph_launches[2][2] = {{0, 0}, {0, 0}};      // kernels inside each captured step (for astrea_launch_count)
#endif
};

namespace {

// instructions that read the ghost rows of a register: the spatial operator and refine_grid (inverse_reconstruct)
bool needs_ghost_rows(const Instr& ins) { return ins.is_operator || ins.special == SP_REFINE; }
int halo_register(const Instr& ins) { return ins.is_operator ? ins.src : ins.out; }

int fail(astrea_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

enum { CLS_SWEEP = 0, CLS_TRANSPOSE = 1, CLS_UPDATE = 2, CLS_HALO = 3, CLS_PRIM = 4, CLS_RECON = 5, CLS_COUNT = 6 };

// Brackets one launch with CUDA events on the context's stream when profiling is on.
struct Timed {
    astrea_ctx* c;
    Timed(astrea_ctx* ctx, int cls) : c(ctx) {
#ifdef ASTREA_DEVICE_BUILD
        if (c->profiling) {
            astrea_ctx::Span sp{cls, nullptr, nullptr};
            cudaEventCreate(&sp.a);
            cudaEventCreate(&sp.b);
            cudaEventRecord(sp.a, c->st.s);
            c->spans.push_back(sp);
        }
#else
        (void)cls;
#endif
    }
    ~Timed() {
#ifdef ASTREA_DEVICE_BUILD
        if (c->profiling) cudaEventRecord(c->spans.back().b, c->st.s);
#endif
        c->launches++;
    }
};

#ifdef ASTREA_DEVICE_BUILD
std::string cuda_text(int e) { return std::string(cudaGetErrorString((cudaError_t)e)); }
#else
std::string cuda_text(int e) { return "hostsim error " + std::to_string(e); }
#endif

// Every entry point that touches CUDA selects the context's device for its duration and restores the caller's
// afterwards: two contexts on different GPUs may be driven from one host thread, and a context may be called from a
// thread whose current device is another one.
struct DeviceGuard {
#ifdef ASTREA_DEVICE_BUILD
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
#else
    explicit DeviceGuard(int) {}
#endif
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define ASTREA_ON_DEVICE(c) DeviceGuard _device_guard((c)->cfg.device)

#define ASTREA_TRY(expr)                                                                   \
    do {                                                                                   \
        const int _e = (expr);                                                             \
        if (_e != 0) return fail(c, _e < 0 ? ASTREA_E_ARG : ASTREA_E_CUDA,                 \
                                 std::string(#expr) + ": " + (_e < 0 ? "unsupported selector combination" : cuda_text(_e))); \
    } while (0)

// doubles between the variables of one plane row: the columns and their ghosts, rounded up to an even count so that
// every row segment starting at an even column is 16-byte aligned (what the bulk copies of the TMA engine need)
int64_t pitch_of(int64_t ncol) { return (ncol + 2 * GHOST + 1) & ~(int64_t)1; }

Plane make_plane(double* mem, int64_t ncol, int ghost_r) {
    Plane p;
    p.col_pitch = pitch_of(ncol);
    p.row_pitch = NVAR * p.col_pitch;
    p.base = mem + (int64_t)ghost_r * p.row_pitch + GHOST;
    return p;
}

bool alloc_reg(astrea_ctx* c, Reg& r, int64_t ncol) {
    r.mem = (double*)dev_alloc(c->plane_doubles * sizeof(double));
    if (!r.mem) return false;
    r.plane = make_plane(r.mem, ncol, c->ghost_r);
    return dev_zero(r.mem, c->plane_doubles * sizeof(double), c->st) == 0;
}

// ---------------------------------------------------------------------------------------- step programs
// Registers: 0 = u (the grid).  Rates: 0 = L of the most recent operator unless a formula re-uses older ones.
void add_op(std::vector<Instr>& p, int src, int rate_out) { p.push_back(Instr{1, src, rate_out, 0, 0, 1.0, {}}); }
void add_comb(std::vector<Instr>& p, int out, double scale, std::vector<Term> t, int bracket = 0, int refine = 1) {
    Instr ins{0, 0, 0, out, bracket, scale, std::move(t)};
    ins.refine = refine;
    p.push_back(std::move(ins));
}
void add_copy(std::vector<Instr>& p, int out, int src) { add_comb(p, out, 1.0, {Term{0, src, 1.0}}, 0, 0); }
Term R(int reg, double coef) { return Term{0, reg, coef}; }
Term L(int rate, double coef) { return Term{1, rate, coef}; }

// evolvers.py:79-206, literal coefficients and evaluation order.  Returns (#registers, #rates).
void build_program(int integrator, bool mhd, std::vector<Instr>& p, int& nregs, int& nrates, int& final_reg) {
    p.clear();
    add_op(p, 0, 0);   // evolve_space on the grid (astrea.py:67)
    switch (integrator) {
        case INT_SSPRK104: {   // evolvers.py:84-103; u = reg0, k = reg1, k5 = reg2, _k = reg3, (increment = reg4)
            nregs = mhd ? 5 : 4; nrates = 1;
            // k += refine_grid(1/6*dt*L): the refinement acts on the increment, so with magnetic_2d the increment
            // is formed in its own register first
            auto increment = [&](int k) {
                if (mhd) {
                    add_comb(p, 4, 1.0, {L(0, 1.0 / 6)}, 0, 1);
                    add_comb(p, k, 1.0, {R(k, 1.0), R(4, 1.0)}, 0, 0);
                } else {
                    add_comb(p, k, 1.0, {R(k, 1.0), L(0, 1.0 / 6)});
                }
            };
            add_copy(p, 1, 0);                                                 // k = copy(grid)
            for (int s = 0; s < 5; ++s) {
                increment(1);
                add_op(p, 1, 0);
            }
            add_comb(p, 2, 1.0, {R(0, 3.0 / 5), R(1, 6.0 / 15), L(0, 1.0 / 15)});
            add_op(p, 2, 0);
            add_copy(p, 3, 2);
            for (int s = 0; s < 4; ++s) {
                increment(3);
                add_op(p, 3, 0);
            }
            add_comb(p, 0, 1.0, {R(0, -11.0 / 35), R(2, 5.0 / 7), R(3, 3.0 / 5), L(0, 1.0 / 10)});
            final_reg = 0;
            break;
        }
        case INT_SSPRK54: {    // evolvers.py:105-124; k1..k4 = reg1..4; rate0 = latest, rate1 = L(k3)
            nregs = 5; nrates = 2;
            add_comb(p, 1, 1.0, {R(0, 1.0), L(0, .39175222657189)});
            add_op(p, 1, 0);
            add_comb(p, 2, 1.0, {R(0, .444370493651235), R(1, .555629506348765), L(0, .368410593050371)});
            add_op(p, 2, 0);
            add_comb(p, 3, 1.0, {R(0, .620101851488403), R(2, .379898148511597), L(0, .251891774271694)});
            add_op(p, 3, 1);
            add_comb(p, 4, 1.0, {R(0, .178079954393132), R(3, .821920045606868), L(1, .544974750228521)});
            add_op(p, 4, 0);
            add_comb(p, 0, 1.0, {R(2, .517231671970585), R(3, .096059710526147), L(1, .06369246866629),
                                 R(4, .386708617503269), L(0, .226007483236906)});
            final_reg = 0;
            break;
        }
        case INT_SSPRK53: {    // evolvers.py:127-146; rate0 = L0, rate1 = L1, rate2 = latest
            nregs = 5; nrates = 3;
            add_comb(p, 1, 1.0, {R(0, 1.0), L(0, .3772689151171)});
            add_op(p, 1, 1);
            add_comb(p, 2, 1.0, {R(1, 1.0), L(1, .3772689151171)});
            add_op(p, 2, 2);
            add_comb(p, 3, 1.0, {R(0, .56656131914033), R(2, .43343868085967), L(2, .16352294089771)});
            add_op(p, 3, 2);
            add_comb(p, 4, 1.0, {R(0, .09299483444413), R(1, .0000209036962), R(3, .90698426185967), L(0, .00071997378654),
                                 L(2, .34217696850008)});
            add_op(p, 4, 2);
            add_comb(p, 0, 1.0, {R(0, .0073613226092), R(1, .20127980325145), R(2, .00182955389682), R(4, .78952932024253),
                                 L(0, .0027771981946), L(1, .00001567934613), L(2, .29786487010104)}, 1);
            final_reg = 0;
            break;
        }
        case INT_SSPRK43: {    // evolvers.py:148-163
            nregs = 2; nrates = 1;
            add_comb(p, 1, 1.0, {R(0, 1.0), L(0, .5)});
            add_op(p, 1, 0);
            add_comb(p, 1, 1.0, {R(1, 1.0), L(0, .5)});
            add_op(p, 1, 0);
            add_comb(p, 1, 1.0 / 6, {R(0, 4.0), R(1, 2.0), L(0, 1.0)});
            add_op(p, 1, 0);
            add_comb(p, 0, 1.0, {R(1, 1.0), L(0, .5)});
            final_reg = 0;
            break;
        }
        case INT_SSPRK33: {    // evolvers.py:165-176
            nregs = 2; nrates = 1;
            add_comb(p, 1, 1.0, {R(0, 1.0), L(0, 1.0)});
            add_op(p, 1, 0);
            add_comb(p, 1, .25, {R(0, 3.0), R(1, 1.0), L(0, 1.0)});
            add_op(p, 1, 0);
            add_comb(p, 0, 1.0 / 3, {R(0, 1.0), R(1, 2.0), L(0, 2.0)});
            final_reg = 0;
            break;
        }
        case INT_SSPRK22: {    // evolvers.py:178-185
            nregs = 2; nrates = 1;
            add_comb(p, 1, 1.0, {R(0, 1.0), L(0, 1.0)});
            add_op(p, 1, 0);
            add_comb(p, 0, .5, {R(0, 1.0), R(1, 1.0), L(0, 1.0)});
            final_reg = 0;
            break;
        }
        case INT_RK4: {        // evolvers.py:187-202; rates 0..3 = L0..L3
            nregs = 2; nrates = 4;
            add_comb(p, 1, 1.0, {R(0, 1.0), L(0, .5)});
            add_op(p, 1, 1);
            add_comb(p, 1, 1.0, {R(0, 1.0), L(1, .5)});
            add_op(p, 1, 2);
            add_comb(p, 1, 1.0, {R(0, 1.0), L(2, 1.0)});
            add_op(p, 1, 3);
            add_comb(p, 0, 1.0 / 6, {R(0, 1.0), L(0, 1.0), L(1, 2.0), L(2, 2.0), L(3, 1.0)}, 1);
            final_reg = 0;
            break;
        }
        default: {             // forward Euler, evolvers.py:204-206
            nregs = 1; nrates = 1;
            add_comb(p, 0, 1.0, {R(0, 1.0), L(0, 1.0)});
            final_reg = 0;
            break;
        }
    }
    // Fuse the rate assembly into the register update that follows an operator when that update is the first reader
    // of the operator's rate buffer; the rate is stored only if a later formula reads it again.
    for (size_t i = 0; i + 1 < p.size(); ++i) {
        if (!p[i].is_operator || p[i + 1].is_operator) continue;
        const int r = p[i].rate_out;
        bool used_next = false;
        for (const Term& t : p[i + 1].terms) used_next = used_next || (t.is_rate && t.index == r);
        if (!used_next) continue;
        bool used_later = false;
        for (size_t k = i + 2; k < p.size(); ++k) {
            if (p[k].is_operator) { if (p[k].rate_out == r) break; else continue; }
            for (const Term& t : p[k].terms) used_later = used_later || (t.is_rate && t.index == r);
        }
        p[i].defer_rate = 1;
        p[i + 1].fused_rate = r;
        p[i + 1].store_rate = used_later ? 1 : 0;
    }
    if (!mhd) return;
    // magnetic_2d: the B slots of the grid become face averages before the stages (evolvers.py:73-76) and every
    // register update is followed by refine_grid (evolvers.py:63-67)
    std::vector<Instr> q;
    for (size_t i = 0; i < p.size(); ++i) {
        q.push_back(p[i]);
        if (i == 0) {
            Instr f{0, 0, 0, 0, 0, 1.0, {}};
            f.special = SP_FACE_FIELD;
            q.push_back(f);
        } else if (!p[i].is_operator && p[i].refine) {
            Instr r{0, 0, 0, p[i].out, 0, 1.0, {}};
            r.special = SP_REFINE;
            q.push_back(r);
        }
    }
    p.swap(q);
}

// ---------------------------------------------------------------------------------------- launches
int fill_halo(astrea_ctx* c, Plane pl, int external_rows) {
    HaloParams h{pl, c->nrow, c->ncol, c->cfg.boundary, 0, 1, 1, c->vars(), 0, 0, 0};
    { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<HaloKernel>(h, 1, (int)c->nrow, 64, 0, c->st)); }
    if (c->ghost_r > 0) {
        h.phase = 1;
        // interior slab edges are provided by the neighbour ranks; a physical 'edge' boundary is always local
        const bool lo_phys = c->cfg.x_offset == 0, hi_phys = c->cfg.x_offset + c->cfg.nx == c->cfg.nx_global;
        const bool single = c->cfg.nx == c->cfg.nx_global;
        if (external_rows && !single) {
            h.fill_lo = (c->cfg.boundary == BC_EDGE && lo_phys) ? 1 : 0;
            h.fill_hi = (c->cfg.boundary == BC_EDGE && hi_phys) ? 1 : 0;
        }
        if (h.fill_lo || h.fill_hi) {
            const int gx = (int)((c->ncol + 2 * GHOST + 255) / 256);
            { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<HaloKernel>(h, gx, 2 * GHOST * NVAR, 256, 0, c->st)); }
        }
    }
    return 0;
}

int transpose_plane(astrea_ctx* c, Plane src, Plane dst, int64_t src_rows, int64_t src_cols) {
    TransposeParams t{src, dst, -(int64_t)GHOST, src_rows + GHOST, -(int64_t)GHOST, src_cols + GHOST, all_vars()};
    const int gx = (int)((src_cols + 2 * GHOST + 31) / 32), gy = (int)((src_rows + 2 * GHOST + 31) / 32);
    Timed timed(c, CLS_TRANSPOSE);
    ASTREA_TRY(launch<TransposeKernel>(t, gx, gy, 256, TransposeKernel::smem_bytes(), c->st));
    return 0;
}

// "pad the derived array" for a plane: ghost columns (phase 0), then ghost rows (phase 1).  A side whose ghost cells
// were computed from genuine neighbour data (slab interior) is left alone.
int fill_plane_halo(astrea_ctx* c, Plane pl, int64_t rows, int64_t cols, bool keep_row_lo, bool keep_row_hi, bool keep_col_lo, bool keep_col_hi) {
    HaloParams h{pl, rows, cols, c->cfg.boundary, 0, keep_row_lo ? 0 : 1, keep_row_hi ? 0 : 1, all_vars(), 0, keep_col_lo ? 1 : 0, keep_col_hi ? 1 : 0};
    // ghost rows that are kept hold genuine values in their interior columns only: their ghost columns are padded too
    h.row0 = keep_row_lo ? -(int64_t)GHOST : 0;
    const int64_t row_end = keep_row_hi ? rows + GHOST : rows;
    if (!(keep_col_lo && keep_col_hi)) { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<HaloKernel>(h, 1, (int)(row_end - h.row0), 64, 0, c->st)); }
    h.phase = 1;
    const int gx = (int)((cols + 2 * GHOST + 255) / 256);
    if (h.fill_lo || h.fill_hi) { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<HaloKernel>(h, gx, 2 * GHOST * NVAR, 256, 0, c->st)); }
    return 0;
}

// mag_field.compute_corner (mag_field.py:125-187) from the face states of the two sweeps -> c->emf
int corner_field(astrea_ctx* c) {
    const astrea_cfg& g = c->cfg;
    const int64_t nx = c->nrow, ny = c->ncol;
    const bool edge = g.boundary == BC_EDGE;
    // The reconstruction of each sweep wrote its face states in the OTHER frame (ReconStageParams::wf_t): wfx holds the
    // face states of the x sweep as a y-frame plane [y][v][x], wfy those of the y sweep as an x-frame plane [x][v][y] —
    // the frames their transverse reconstruction marches in (mag_field.py:15 ``wF.transpose(ortho_axis)``).
    // "pad the derived array": ghost cells of the face-state arrays are copies, not reconstructions
    const bool xl = c->ext_lo(), xh = c->ext_hi();
    const Plane fx = make_plane(c->wfx.mem, nx, GHOST), fy = make_plane(c->wfy.mem, ny, GHOST);
    if (int e = fill_plane_halo(c, fx, ny, nx, false, false, xl, xh)) return e;      // y frame: columns are x
    if (int e = fill_plane_halo(c, fy, nx, ny, xl, xh, false, false)) return e;      // x frame: rows are x
    auto transverse_ppm = [&](Plane face_t, Plane d, Plane u, int transposed, int64_t ns, int64_t ns_glob, int64_t s_off, int64_t nt, int64_t i_hi) -> int {
        ReconStageParams rp{};
        rp.w = face_t; rp.wp = d; rp.wm = u; rp.wf = Plane{nullptr, 0, 0};
        rp.ns = ns; rp.ns_glob = ns_glob; rp.s_off = s_off; rp.c_lo = 0; rp.c_hi = nt;
        rp.i_lo = 0; rp.i_hi = i_hi;
        rp.bc = g.boundary; rp.limiter = LIM_MINMOD; rp.seg = 64; rp.cell_aligned = 1; rp.out_t = transposed;
        // compute_corner reads rho, v_x, v_y, P and B of the corner states (mag_field.py:128-185): v_z is never used
        rp.nvar = NVAR - 1;
        const int used[NVAR] = {0, 1, 2, 4, 5, 6, 7, 0};
        for (int k = 0; k < NVAR; ++k) rp.vars[k] = used[k];
        rp.bulk = (g.flags & 4) ? 0 : 1;
        rp.ppm_author = PPM_MC; rp.pass = 0; rp.force_any3 = 0; rp.ppm_flags = c->ppm_flags; rp.nt = nt;   // mag_field.py:11: author='mc'
        const int nthreads = 128;
        const int gx = (int)((nt + nthreads - 1) / nthreads);
        const int nseg = (int)((rp.i_hi - rp.i_lo + 1 + rp.seg - 1) / rp.seg);
        Timed timed(c, CLS_RECON);
        ASTREA_TRY(launch_recon(SCH_PPM, rp, gx, nseg * rp.nvar, nthreads, c->st));
        return 0;
    };
    // bundle 1: face states of the x sweep, reconstructed along y (marching in the y frame), corner states written
    // transposed, i.e. straight into x-frame planes.  Cells 0..ny along y (pad(wD)[1:] needs cell ny when periodic);
    // one more column of x when the row behind the slab is genuine.
    const Plane d1 = make_plane(c->qT.mem, ny, GHOST), u1 = make_plane(c->ws.mem, ny, GHOST);
    if (int e = transverse_ppm(fx, d1, u1, 1, ny, ny, 0, xh ? nx + 1 : nx, edge ? ny - 1 : ny)) return e;
    // bundle 0: face states of the y sweep, reconstructed along x (x frame); cells 0..nx (+1 behind a slab: the corner
    // row nx needs wD of cell nx + 1)
    const Plane d0 = make_plane(c->wm.mem, ny, GHOST), u0 = make_plane(c->ct0.mem, ny, GHOST);
    if (int e = transverse_ppm(fy, d0, u0, 0, nx, g.nx_global, g.x_offset, ny, xh ? nx + 1 : (edge ? nx - 1 : nx))) return e;
    c->emf_rows = xh ? nx + 1 : nx;
    CornerEmfParams ep{d0, u0, d1, u1, c->emf, nx, ny, g.nx_global, g.x_offset, g.gamma, g.boundary, c->parity};
    Timed timed(c, CLS_UPDATE);
    ASTREA_TRY(launch<CornerEmfKernel>(ep, (int)((ny + 127) / 128), (int)c->emf_rows, 128, 0, c->st));
    return 0;
}

int run_special(astrea_ctx* c, const Instr& ins, int external_rows) {
    const astrea_cfg& g = c->cfg;
    if (ins.special == SP_FACE_FIELD) {
        FaceFieldParams fp{c->regs[c->grid_reg].plane, make_plane(c->wfx.mem, c->nrow, GHOST), make_plane(c->wfy.mem, c->ncol, GHOST), c->nrow, c->ncol};
        Timed timed(c, CLS_UPDATE);
        ASTREA_TRY(launch<FaceFieldKernel>(fp, (int)((c->ncol + 31) / 32), (int)((c->nrow + 31) / 32), 256, FaceFieldKernel::smem_bytes(), c->st));
        return 0;
    }
    // refine_grid = mag_field.inverse_reconstruct on register `out`
    Plane reg = c->regs[ins.out].plane;
    if (int e = fill_halo(c, reg, external_rows)) return e;
    RefineFieldParams rp{reg, make_plane(c->ws.mem, c->ncol, GHOST), c->nrow, c->ncol, g.nx_global, g.x_offset, g.boundary, 0};
    const int gx = (int)((c->ncol + RefineFieldKernel::TX - 1) / RefineFieldKernel::TX), gy = (int)((c->nrow + RefineFieldKernel::TY - 1) / RefineFieldKernel::TY);
    { Timed timed(c, CLS_UPDATE); ASTREA_TRY(launch<RefineFieldKernel>(rp, gx, gy, 256, RefineFieldKernel::smem_bytes(), c->st)); }
    rp.copy_back = 1;       // the refinement reads neighbours of what it replaces: out of place, then copied back
    { Timed timed(c, CLS_UPDATE); ASTREA_TRY(launch<RefineFieldKernel>(rp, gx, gy, 256, 0, c->st)); }
    return 0;
}

RateParams rate_params(astrea_ctx* c) {
    const astrea_cfg& g = c->cfg;
    RateParams r{};
    r.f0 = c->d0.plane; r.f1t = c->d1t.plane; r.d0 = c->d0.plane;
    r.nrow = c->nrow; r.ncol = c->ncol; r.dimension = g.dimension; r.emf = g.magnetic_2d ? c->emf : nullptr; r.emf_rows = c->emf_rows;
    r.nx_glob = g.nx_global; r.x_off = g.x_offset; r.dx = g.dx; r.bc = g.boundary;
    {   // dx = 2^k with 1 / dx representable: the update multiplies instead of dividing (aux_kernels.cuh)
        int e = 0;
        const double m = std::frexp(g.dx, &e);
        r.inv_dx_exact = (m == 0.5 && e > -1000 && e < 1000) ? 1.0 / g.dx : 0.0;
    }
    r.vars = c->vars();
    r.row_lo = 0; r.row_hi = c->nrow;
    return r;
}

int run_operator(astrea_ctx* c, const Instr& ins, int external_rows, bool first) {
    const astrea_cfg& g = c->cfg;
    Plane q = c->regs[ins.src].plane;
    if (int e = fill_halo(c, q, external_rows)) return e;
    unsigned long long* eig = first ? c->eig_bits : c->eig_scratch;
    ASTREA_TRY(dev_zero(eig, 2 * sizeof(unsigned long long), c->st));
    const bool lw = g.solver == SOL_LW;
    if (lw && !(g.dimension == 1 ? c->field_free : c->hydro))
        return fail(c, ASTREA_E_ARG, "Lax-Wendroff on the device needs a grid without v_z / B (its column pick follows LAPACK's "
                                     "eigenvalue slot order, which is only reproducible for that spectrum; SURVEY Q11)");
    if (g.dimension == 1) {
        Sweep1DParams p{};
        p.q = q; p.d = c->d0.plane; p.n = g.ny; p.gamma = g.gamma; p.dx = g.dx;
        p.bc = g.boundary; p.limiter = g.limiter; p.low_mach = g.low_mach; p.tile = c->tile1d;
        p.eigmax_bits = eig; p.flag = c->flag;
        p.ppm_author = g.ppm_author; p.pass = 0; p.ppm_flags = c->ppm_flags;
        p.lw_pass = 0; p.lw_keys = c->lw_keys;
        if (lw) ASTREA_TRY(dev_ones(c->lw_keys, 4 * sizeof(unsigned long long), c->st));
        if (g.scheme == SCH_PPM && g.ppm_author != PPM_MC) {
            // the grid-wide switches of the interface / extrapolant limiters first (two flag passes)
            ASTREA_TRY(dev_zero(c->ppm_flags, 4 * sizeof(int), c->st));
            for (int pass = 1; pass <= 2; ++pass) {
                p.pass = pass;
                Timed timed(c, CLS_RECON);
                ASTREA_TRY(launch_sweep1d(g.scheme, g.solver, p, c->threads1d, c->st));
            }
            p.pass = 0;
        }
        if (lw) {      // search pass of the Lax-Wendroff column pick
            p.lw_pass = 1;
            { Timed timed(c, CLS_SWEEP); ASTREA_TRY(launch_sweep1d(g.scheme, g.solver, p, c->threads1d, c->st)); }
            p.lw_pass = 0;
        }
        { Timed timed(c, CLS_SWEEP); ASTREA_TRY(launch_sweep1d(g.scheme, g.solver, p, c->threads1d, c->st)); }
    } else {
        // sweep order and the solver's private axis counter (solvers.py:34-36,63; astrea.py:85)
        const int order[2] = {c->parity ? 1 : 0, c->parity ? 0 : 1};
        const bool ho = scheme_high_order(g.scheme), pcm = g.scheme == SCH_PCM;
        const int kind = pcm ? 0 : (ho ? 2 : 1);
        const int lo = recon_lo(g.scheme), hi = recon_hi(g.scheme);
        const int ht = ho ? 2 : 1;                 // transverse reach of the flux stage
        if (pcm) {
            // PCM's flux stage reads q itself (pcm.py:33-34): the y sweep needs the transposed copy of the register
            TransposeParams t{q, c->qT.plane, -(int64_t)GHOST, c->nrow + GHOST, -(int64_t)GHOST, c->ncol + GHOST, c->vars()};
            const int gx = (int)((c->ncol + 2 * GHOST + 31) / 32), gy = (int)((c->nrow + 2 * GHOST + 31) / 32);
            Timed timed(c, CLS_TRANSPOSE);
            ASTREA_TRY(launch<TransposeKernel>(t, gx, gy, 256, TransposeKernel::smem_bytes(), c->st));
        } else {
            // primitive averages of both sweep frames from one read of the register (PrimBothStage)
            PrimBothParams pb{};
            pb.q = q; pb.wx = make_plane(c->ws.mem, c->ncol, GHOST); pb.wy = make_plane(c->ws2.mem, c->nrow, GHOST);
            pb.gamma = g.gamma; pb.high_order = ho ? 1 : 0;
            pb.r_lo = -(int64_t)(lo + 1); pb.r_hi = c->nrow + hi + 2; pb.c_lo = -(int64_t)(lo + 1); pb.c_hi = c->ncol + hi + 2;
            if (g.magnetic_2d && c->slab()) {
                // constrained transport reconstructs the y-sweep face states along x: they are recomputed in the
                // ghost rows of a slab (CT_LO / CT_HI cells deep) instead of being exchanged
                pb.r_lo = std::min<int64_t>(pb.r_lo, -(int64_t)CT_LO); pb.r_hi = std::max<int64_t>(pb.r_hi, c->nrow + CT_HI);
            }
            pb.r_min = -(int64_t)GHOST; pb.r_max = c->nrow + GHOST - 1; pb.c_min = -(int64_t)GHOST; pb.c_max = c->ncol + GHOST - 1;
            const int gx = (int)((pb.c_hi - pb.c_lo + PrimBothStage<false>::TX - 1) / PrimBothStage<false>::TX);
            const int gy = (int)((pb.r_hi - pb.r_lo + PrimBothStage<false>::TY - 1) / PrimBothStage<false>::TY);
            Timed timed(c, CLS_PRIM);
            if (c->hydro) ASTREA_TRY(launch<PrimBothStage<true>>(pb, gx, gy, 256, PrimBothStage<true>::smem_bytes(), c->st));
            else ASTREA_TRY(launch<PrimBothStage<false>>(pb, gx, gy, 256, PrimBothStage<false>::smem_bytes(), c->st));
        }
        for (int k = 0; k < 2; ++k) {
            const int ax = order[k], sax = k;
            int64_t ns, nt, ns_glob, s_off, nt_glob, t_off;
            Plane qf, ff;
            if (ax == 0) {
                qf = q; ff = c->d0.plane;
                ns = c->nrow; nt = c->ncol; ns_glob = g.nx_global; s_off = g.x_offset; nt_glob = c->ncol; t_off = 0;
            } else {
                qf = c->qT.plane; ff = c->d1t.plane;
                ns = c->ncol; nt = c->nrow; ns_glob = c->ncol; s_off = 0; nt_glob = g.nx_global; t_off = g.x_offset;
            }
            // scratch planes in the shape of this frame
            const Plane ws = make_plane((ax == 1 && !pcm) ? c->ws2.mem : c->ws.mem, nt, GHOST);
            const Plane wp = make_plane(c->wp.mem, nt, GHOST), wm = make_plane(c->wm.mem, nt, GHOST);
            const bool edge = g.boundary == BC_EDGE;
            const bool lo_phys = s_off == 0, hi_phys = s_off + ns == ns_glob;
            // cells to reconstruct: one beyond each end (two at the upper end: LLF looks at interface j+1), except
            // across a physical 'edge' boundary where the interface states are copies ("pad the derived array")
            const int64_t i_lo = (edge && lo_phys) ? 0 : -1, i_hi = (edge && hi_phys) ? ns - 1 : ns + 1;
            if (pcm) {
                PrimStageParams pp{};
                pp.q = qf; pp.w = ws; pp.gamma = g.gamma; pp.high_order = ho ? 1 : 0;
                pp.r_lo = -(int64_t)(lo + 1); pp.r_hi = ns + hi + 2; pp.c_lo = -(int64_t)ht; pp.c_hi = nt + ht;
                if (g.magnetic_2d && ax == 0) {
                    // pcm.py:35: the face state is wS itself.  The x-frame primitives are kept as the face states of the y
                    // sweep (corner_field), which constrained transport reconstructs along x: also in the ghost rows of a slab
                    if (c->ext_lo()) pp.r_lo = std::min<int64_t>(pp.r_lo, -(int64_t)CT_LO);
                    if (c->ext_hi()) pp.r_hi = std::max<int64_t>(pp.r_hi, ns + CT_HI);
                }
                pp.r_min = -(int64_t)GHOST; pp.r_max = ns + GHOST - 1; pp.c_min = -(int64_t)GHOST; pp.c_max = nt + GHOST - 1;
                const int gx = (int)((pp.c_hi - pp.c_lo + PrimStage<false>::TX - 1) / PrimStage<false>::TX);
                const int gy = (int)((pp.r_hi - pp.r_lo + PrimStage<false>::TY - 1) / PrimStage<false>::TY);
                Timed timed(c, CLS_PRIM);
                if (c->hydro) ASTREA_TRY(launch<PrimStage<true>>(pp, gx, gy, 256, PrimStage<true>::smem_bytes(pp.high_order), c->st));
                else ASTREA_TRY(launch<PrimStage<false>>(pp, gx, gy, 256, PrimStage<false>::smem_bytes(pp.high_order), c->st));
            }
            if (pcm && g.magnetic_2d)      // pcm.py:35: the face state is the cell average
                // the face states of a sweep are kept in the other frame (corner_field): the pointwise primitives of the
                // x frame are the y sweep's face states as an x-frame plane, and the other way round
                ASTREA_TRY(copy_d2d(ax == 0 ? c->wfy.mem : c->wfx.mem, c->ws.mem, c->plane_doubles * sizeof(double), c->st));
            if (!pcm) {
                ReconStageParams rp{};
                rp.w = ws; rp.wp = wp; rp.wm = wm; rp.wf = Plane{nullptr, 0, 0};
                if (g.magnetic_2d) {       // data[axes]['wF'] (plm.py:57, ppm.py:101, weno.py:184), written in the other frame
                    rp.wf = (ax == 0) ? make_plane(c->wfx.mem, c->nrow, GHOST) : make_plane(c->wfy.mem, c->ncol, GHOST);
                    rp.wf_t = 1;
                }
                rp.ns = ns; rp.ns_glob = ns_glob; rp.s_off = s_off; rp.c_lo = -(int64_t)ht; rp.c_hi = nt + ht;
                if (g.magnetic_2d && ax == 1) {      // face states of the y sweep in the ghost rows of a slab (columns here)
                    if (c->ext_lo()) rp.c_lo = -(int64_t)CT_LO;
                    if (c->ext_hi()) rp.c_hi = nt + CT_HI;
                }
                rp.i_lo = i_lo; rp.i_hi = i_hi; rp.bc = g.boundary; rp.limiter = g.limiter;
                // cells a thread marches: the window is primed once per segment, so longer segments on long sweeps
                // (measured at 8192^2: 32 / 64 / 128 / 256 cells -> 11.54 / 11.19 / 11.05 / 11.10 ms of reconstruction per step)
                rp.seg = g.segment_2d > 0 ? g.segment_2d : (ns >= 4096 ? 128 : 64);
                // bulk-copy march (ReconStage BULK): row segments must start at an even column; PLM's range starts at -1 and
                // may take one more ghost column (the primitive stage fills columns from -(lo + 1) = -2)
                rp.bulk = (g.flags & 4) ? 0 : 1;
                if (rp.bulk && (rp.c_lo & 1)) { if (rp.c_lo == -1) rp.c_lo = -2; else rp.bulk = 0; }
                const VarList vl = c->vars();
                rp.nvar = vl.n;
                for (int k = 0; k < NVAR; ++k) rp.vars[k] = vl.v[k];
                const int nthreads = 128;
                const int gx = (int)((rp.c_hi - rp.c_lo + nthreads - 1) / nthreads);
                const int nseg = (int)((i_hi - i_lo + 1 + rp.seg - 1) / rp.seg);
                rp.ppm_author = g.ppm_author; rp.pass = 0; rp.ppm_flags = c->ppm_flags; rp.nt = nt;
                rp.force_any3 = c->hydro ? 1 : 0;       // an identically zero variable makes `cell_extrema.any()` true (0 * 0 <= 0)
                if (g.scheme == SCH_PPM && g.ppm_author != PPM_MC) {
                    if (c->slab() && c->reduce_fn == nullptr)
                        return fail(c, ASTREA_E_STATE, "PPM authors 'c' / 'ph' switch on grid-wide any() tests (limiters.py:58,164): a decomposed grid "
                                                       "needs astrea_set_flag_reducer");
                    ASTREA_TRY(dev_zero(c->ppm_flags, 4 * sizeof(int), c->st));
                    for (int pass = 1; pass <= 2; ++pass) {
                        rp.pass = pass;
                        { Timed timed(c, CLS_RECON); ASTREA_TRY(launch_recon(g.scheme, rp, gx, nseg * rp.nvar, nthreads, c->st)); }
                        // the any() is over the whole grid: OR the switches of all slabs before the next pass reads them
                        if (c->slab() && c->reduce_fn(c->reduce_user, c->ppm_flags, 4) != 0)
                            return fail(c, ASTREA_E_STATE, "the flag reducer reported a failure");
                    }
                    rp.pass = 0;
                }
                Timed timed(c, CLS_RECON);
                ASTREA_TRY(launch_recon(g.scheme, rp, gx, nseg * rp.nvar, nthreads, c->st));
            }
            {
                FluxStageParams fp{};
                fp.wp = wp; fp.wm = wm; fp.ws = ws; fp.q = qf; fp.f = ff;
                fp.ns = ns; fp.nt = nt; fp.ns_glob = ns_glob; fp.s_off = s_off; fp.nt_glob = nt_glob; fp.t_off = t_off;
                fp.gamma = g.gamma; fp.bc = g.boundary; fp.low_mach = g.low_mach;
                fp.eigmax_bits = eig + ax; fp.flag = c->flag;
                const int nthreads = (g.threads_2d >= 32 && g.threads_2d <= 128) ? g.threads_2d / 32 * 32 : 128;
                // block-wide rows of transverse points pay off on wide grids only (stages2d.cuh, FluxStage BT);
                // flags bit 3 / bit 4 force warp rows / block rows (A/B measurements, geometry tests)
                fp.block_tile = (g.flags & 8) ? 0 : ((g.flags & 16) ? 1 : (nt >= 4096 ? 1 : 0));
                int own = 0, nwarp = 0;
                flux_stage_geometry(kind, g.solver, (!lw && c->hydro) ? 1 : 0, fp.block_tile, nthreads, own, nwarp);
                const int gx = (int)((nt + own - 1) / own), gy = (int)((ns + 1 + nwarp - 1) / nwarp);
                fp.lw_pass = 0; fp.lw_keys = c->lw_keys;
                fp.need_speed = (first || (g.flags & 64)) ? 1 : 0;      // flags bit 6: evaluate them in every operator (A/B runs)
                if (lw) {      // search pass of the Lax-Wendroff column pick, per sweep
                    if (c->slab() && c->key_reduce_fn == nullptr)
                        return fail(c, ASTREA_E_STATE, "Lax-Wendroff picks its spectrum column over the whole grid (solvers.py:79-88): a decomposed "
                                                       "grid needs astrea_set_key_reducer");
                    ASTREA_TRY(dev_ones(c->lw_keys, 4 * sizeof(unsigned long long), c->st));
                    fp.lw_pass = 1;
                    { Timed timed(c, CLS_SWEEP); ASTREA_TRY(launch_flux(kind, g.solver, ax, sax, 1, fp, gx, gy, nthreads, c->st)); }
                    fp.lw_pass = 0;
                    // first non-zero entry of each column over the whole grid: the smallest key of all slabs
                    if (c->slab() && c->key_reduce_fn(c->key_reduce_user, c->lw_keys, 4) != 0)
                        return fail(c, ASTREA_E_STATE, "the key reducer reported a failure");
                }
                Timed timed(c, CLS_SWEEP);
                ASTREA_TRY(launch_flux(kind, g.solver, ax, sax, c->hydro ? 1 : 0, fp, gx, gy, nthreads, c->st));
            }
        }
    }
    if (g.magnetic_2d) {
        if (int e = corner_field(c)) return e;
    }
    if (ins.defer_rate) return 0;      // the register update that follows assembles L on the fly (UpdateKernel)
    RateParams r = rate_params(c);
    r.out = c->rates[ins.rate_out].plane;
    {
        const int gx = (int)((c->ncol + 31) / 32), gy = (int)((c->nrow + 31) / 32);
        { Timed timed(c, CLS_UPDATE); ASTREA_TRY(launch<RateKernel>(r, gx, gy, 256, RateKernel::smem_bytes(), c->st)); }
    }
    return 0;
}

// rows [row_lo, row_hi) of one register update
int run_combine(astrea_ctx* c, const Instr& ins, int64_t row_lo, int64_t row_hi) {
    if (row_hi <= row_lo) return 0;
    CombineParams p{};
    p.out = c->regs[ins.out].plane;
    p.nterms = (int)ins.terms.size();
    for (int k = 0; k < p.nterms; ++k) {
        const Term& t = ins.terms[k];
        p.term[k] = t.is_rate ? c->rates[t.index].plane : c->regs[t.index].plane;
        p.coef[k] = t.coef;
        p.is_rate[k] = (t.is_rate && t.index == ins.fused_rate) ? 2 : t.is_rate;
    }
    p.bracket_rates = ins.bracket;
    p.scale = ins.scale;
    p.dt = c->dt_dev;
    p.nrow = c->nrow; p.ncol = c->ncol;
    p.vars = c->vars();
    p.row_lo = row_lo;
    if (ins.fused_rate >= 0) {
        UpdateParams u{rate_params(c), p, ins.store_rate ? c->rates[ins.fused_rate].plane : Plane{nullptr, 0, 0}};
        u.rate.row_lo = row_lo; u.rate.row_hi = row_hi;
        const bool one_d = c->cfg.dimension == 1;
        const int gx = one_d ? (int)((c->ncol + 255) / 256) : (int)((c->ncol + 31) / 32);
        const int gy = one_d ? 1 : (int)((row_hi - row_lo + 31) / 32);
        Timed timed(c, CLS_UPDATE);
        const size_t sm = UpdateKernel<1, false>::smem_bytes();
        int e = -1;
        if (!p.bracket_rates) {
            switch (p.nterms) {
                case 1: e = launch<UpdateKernel<1, false>>(u, gx, gy, 256, sm, c->st); break;
                case 2: e = launch<UpdateKernel<2, false>>(u, gx, gy, 256, sm, c->st); break;
                case 3: e = launch<UpdateKernel<3, false>>(u, gx, gy, 256, sm, c->st); break;
                case 4: e = launch<UpdateKernel<4, false>>(u, gx, gy, 256, sm, c->st); break;
                case 5: e = launch<UpdateKernel<5, false>>(u, gx, gy, 256, sm, c->st); break;
                default: break;
            }
        } else {
            switch (p.nterms) {
                case 5: e = launch<UpdateKernel<5, true>>(u, gx, gy, 256, sm, c->st); break;
                case 7: e = launch<UpdateKernel<7, true>>(u, gx, gy, 256, sm, c->st); break;
                default: break;
            }
        }
        ASTREA_TRY(e);
        return 0;
    }
    const int gx = (int)((c->ncol + 255) / 256);
    { Timed timed(c, CLS_UPDATE); ASTREA_TRY(launch<CombineKernel>(p, gx, (int)(row_hi - row_lo), 256, 0, c->st)); }
    return 0;
}

int check_cfg(const astrea_cfg* g, std::string& why) {
    if (!g) { why = "cfg is NULL"; return -1; }
    if (g->dimension != 1 && g->dimension != 2) { why = "dimension must be 1 or 2"; return -1; }
    if (g->boundary != ASTREA_EDGE && g->boundary != ASTREA_WRAP) { why = "boundary must be ASTREA_EDGE or ASTREA_WRAP"; return -1; }
    if (g->scheme < ASTREA_PCM || g->scheme > ASTREA_WENO7) { why = "unknown scheme"; return -1; }
    if (g->ppm_author < ASTREA_PPM_MC || g->ppm_author > ASTREA_PPM_PH) { why = "unknown PPM author"; return -1; }
    if (g->limiter < ASTREA_MINMOD || g->limiter > ASTREA_SUPERBEE) { why = "unknown slope limiter"; return -1; }
    if (g->solver < ASTREA_LLF || g->solver > ASTREA_HLLD) { why = "unknown solver"; return -1; }
    if (g->solver == ASTREA_LW && g->magnetic_2d) {
        why = "Lax-Wendroff (solvers.py:79-88) picks a spectrum column by a grid-wide lexicographic sort (SURVEY Q11): "
              "available for states without v_z / B";
        return -1;
    }
    if (g->integrator < ASTREA_EULER || g->integrator > ASTREA_SSPRK104) { why = "unknown integrator"; return -1; }
    if (g->magnetic_2d) {
        if (g->dimension != 2) { why = "magnetic_2d needs dimension == 2"; return -1; }
    }
    if (g->dimension == 1) {
        if (g->nx < 1 || g->ny != 1) { why = "1D: nx >= 1 cells, ny == 1"; return -1; }
        if (g->nx_global != g->nx || g->x_offset != 0) { why = "1D grids are not decomposed"; return -1; }
    } else {
        if (g->nx < 1 || g->ny < 1) { why = "2D: nx, ny >= 1"; return -1; }
        if (g->nx_global < g->nx || g->x_offset < 0 || g->x_offset + g->nx > g->nx_global) { why = "slab outside the global grid"; return -1; }
        if (g->nx != g->nx_global && g->nx < GHOST) { why = "a slab needs at least GHOST rows"; return -1; }
    }
    if (!(g->gamma > 1.0) || !(g->dx > 0.0) || !(g->cfl > 0.0)) { why = "gamma > 1, dx > 0, cfl > 0 required"; return -1; }
    return 0;
}

}  // namespace

extern "C" {

astrea_ctx* astrea_create(const astrea_cfg* cfg) {
    std::string why;
    if (check_cfg(cfg, why) != 0) { g_create_error = why; return nullptr; }
    astrea_ctx* c = new astrea_ctx();
    c->cfg = *cfg;
    ASTREA_ON_DEVICE(c);            // the caller's current device is restored on return
#ifdef ASTREA_DEVICE_BUILD
    {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e == cudaSuccess && (cfg->device < 0 || cfg->device >= ndev)) e = cudaErrorInvalidDevice;
        if (e == cudaSuccess) e = cudaSetDevice(cfg->device);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->st.s, cudaStreamNonBlocking);
        if (e != cudaSuccess) { g_create_error = std::string("CUDA: ") + cudaGetErrorString(e); delete c; return nullptr; }
        c->stream_owned = true;
    }
#endif
    const astrea_cfg& g = c->cfg;
    if (g.dimension == 1) { c->nrow = 1; c->ncol = g.nx; c->ghost_r = 0; c->cfg.ny = g.nx; c->cfg.nx = 1; c->cfg.nx_global = 1; }
    else { c->nrow = g.nx; c->ncol = g.ny; c->ghost_r = GHOST; }
    // the transposed planes of the y sweep have the same number of elements
    c->plane_doubles = (size_t)(c->nrow + 2 * c->ghost_r) * NVAR * (size_t)pitch_of(c->ncol);
    if (g.dimension == 2) c->plane_doubles = std::max(c->plane_doubles, (size_t)(c->ncol + 2 * GHOST) * NVAR * (size_t)pitch_of(c->nrow));

    int nregs = 1, nrates = 1;
    build_program(g.integrator, g.magnetic_2d != 0, c->prog, nregs, nrates, c->final_reg);
    bool ok = true;
    c->regs.resize(nregs);
    c->rates.resize(nrates);
    for (auto& r : c->regs) ok = ok && alloc_reg(c, r, c->ncol);
    for (auto& r : c->rates) ok = ok && alloc_reg(c, r, c->ncol);
    ok = ok && alloc_reg(c, c->d0, c->ncol);
    if (g.dimension == 2) {
        ok = ok && alloc_reg(c, c->qT, c->nrow) && alloc_reg(c, c->d1t, c->nrow);
        ok = ok && alloc_reg(c, c->ws, c->ncol) && alloc_reg(c, c->ws2, c->nrow) && alloc_reg(c, c->wp, c->ncol) && alloc_reg(c, c->wm, c->ncol);
        if (g.magnetic_2d) {
            ok = ok && alloc_reg(c, c->wfx, c->ncol) && alloc_reg(c, c->wfy, c->nrow) && alloc_reg(c, c->ct0, c->ncol);
            c->emf = (double*)dev_alloc(sizeof(double) * (size_t)(c->nrow + 1) * c->ncol);
            ok = ok && c->emf;
        }
    } else {
        ok = ok && alloc_reg(c, c->qT, c->ncol);   // scratch for primitive downloads
    }
    c->mhd_flag = (int*)dev_alloc(sizeof(int));
    c->ppm_flags = (int*)dev_alloc(4 * sizeof(int));
    c->lw_keys = (unsigned long long*)dev_alloc(4 * sizeof(unsigned long long));
    ok = ok && c->mhd_flag && c->ppm_flags && c->lw_keys;
    if (ok) dev_zero(c->ppm_flags, 4 * sizeof(int), c->st);
    c->eig_bits = (unsigned long long*)dev_alloc(8 * sizeof(unsigned long long));
    c->clock = (double*)dev_alloc((4 + DT_HISTORY) * sizeof(double));
    c->dt_dev = (double*)dev_alloc(sizeof(double));
    ok = ok && c->eig_bits && c->clock && c->dt_dev;
    if (!ok) {
        g_create_error = "device allocation failed";
        astrea_destroy(c);
        return nullptr;
    }
    c->flag = c->eig_bits + 2;
    c->eig_scratch = c->eig_bits + 4;
    c->dt_history = c->clock + 4;
    dev_zero(c->eig_bits, 8 * sizeof(unsigned long long), c->st);
    dev_zero(c->clock, (4 + DT_HISTORY) * sizeof(double), c->st);
    dev_zero(c->dt_dev, sizeof(double), c->st);

    // launch geometry
    if (g.dimension == 1) {
        const int lo = recon_lo(g.scheme) + 2, hi = recon_hi(g.scheme) + 3;
        int tile = g.tile_1d > 0 ? g.tile_1d : 256 - lo - hi;
        tile = std::max(1, std::min(tile, 256 - lo - hi));
        c->tile1d = tile;
        c->threads1d = tile + lo + hi;
    }
    return c;
}

void astrea_destroy(astrea_ctx* c) {
    if (!c) return;
    ASTREA_ON_DEVICE(c);
    stream_sync(c->st);
    for (auto& r : c->regs) dev_free(r.mem);
    for (auto& r : c->rates) dev_free(r.mem);
    dev_free(c->qT.mem); dev_free(c->d0.mem); dev_free(c->d1t.mem);
    dev_free(c->ws.mem); dev_free(c->ws2.mem); dev_free(c->wp.mem); dev_free(c->wm.mem);
    dev_free(c->wfx.mem); dev_free(c->wfy.mem); dev_free(c->ct0.mem); dev_free(c->emf);
    dev_free(c->eig_bits); dev_free(c->clock); dev_free(c->dt_dev); dev_free(c->saved.mem); dev_free(c->mhd_flag); dev_free(c->ppm_flags); dev_free(c->lw_keys);
    dev_free(c->snap_dev);
#ifdef ASTREA_DEVICE_BUILD
    for (int t = 0; t < astrea_ctx::LANES; ++t) {
        if (c->lane_st[t]) { cudaStreamSynchronize(c->lane_st[t]); cudaStreamDestroy(c->lane_st[t]); }
        for (int k = 0; k < 2; ++k) {
            if (c->bounce[t][k]) cudaFreeHost(c->bounce[t][k]);
            if (c->bounce_done[t][k]) cudaEventDestroy(c->bounce_done[t][k]);
        }
    }
    if (c->lane_begin) cudaEventDestroy(c->lane_begin);
#endif
#ifdef ASTREA_DEVICE_BUILD
    if (c->copy_st) { cudaStreamSynchronize(c->copy_st); cudaStreamDestroy(c->copy_st); }
    if (c->snap_ready) cudaEventDestroy(c->snap_ready);
    for (auto& e : c->snap_done) if (e) cudaEventDestroy(e);
    for (auto& row : c->step_graph)
        for (auto& g : row)
            if (g) cudaGraphExecDestroy(g);
    if (c->stream_owned) cudaStreamDestroy(c->st.s);
#endif
    delete c;
}

const char* astrea_last_error(const astrea_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

#ifdef ASTREA_DEVICE_BUILD
// A pageable host array is staged by the driver through one small bounce buffer at a fraction of the PCIe rate
// (measured on the B200 box: 21 GB/s down, against 55 GB/s for page-locked memory).  Large pageable arrays (the
// reference's first grid, astrea.py:35: 4.3 GB at 8192^2; np.empty destinations) are cut into one contiguous part per
// host thread instead; every thread moves its part in 4 MB pieces through two page-locked buffers on a stream of its
// own, so that its memcpy of one piece overlaps the DMA of the previous one and the parts overlap each other.  The
// context's stream waits for the lanes (uploads) or the host does (downloads).  Page-locked arrays (what evolve_time
// returns) are copied directly.
static int staged_copy(astrea_ctx* c, void* dev, void* host, size_t bytes, bool to_device) {
    constexpr size_t PIECE = 4u << 20;     // 2 MB / 1 MB / 512 KB pieces: 100 / 131 / 174 ms for the 4.3 GB of 8192^2 against 87
    auto direct = [&]() { return to_device ? copy_h2d(dev, host, bytes, c->st) : copy_d2h(host, dev, bytes, c->st); };
    cudaPointerAttributes attr;
    const bool pageable = cudaPointerGetAttributes(&attr, host) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered;
    cudaGetLastError();
    if (!pageable || bytes < 32 * PIECE) return direct();
    const unsigned hw = std::thread::hardware_concurrency();
    const int lanes = (int)std::max(1u, std::min((unsigned)astrea_ctx::LANES, hw / 2));
    for (int t = 0; t < lanes; ++t) {
        cudaError_t e = cudaSuccess;
        if (!c->lane_st[t]) e = cudaStreamCreateWithFlags(&c->lane_st[t], cudaStreamNonBlocking);
        for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
            if (!c->bounce[t][k]) e = cudaHostAlloc(&c->bounce[t][k], PIECE, cudaHostAllocDefault);
            if (e == cudaSuccess && !c->bounce_done[t][k]) e = cudaEventCreateWithFlags(&c->bounce_done[t][k], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) { cudaGetLastError(); return direct(); }
    }
    if (!c->lane_begin && cudaEventCreateWithFlags(&c->lane_begin, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return direct(); }
    // the lanes start after what the context's stream has queued so far (the staging plane may still be read)
    if (cudaEventRecord(c->lane_begin, c->st.s) != cudaSuccess) return (int)cudaGetLastError();
    const size_t part = ((bytes + lanes - 1) / lanes + 255) / 256 * 256;
    std::vector<int> status(lanes, 0);
    auto lane = [&](int t) {
        auto ok = [&](cudaError_t e) { if (e != cudaSuccess && !status[t]) status[t] = (int)e; return e == cudaSuccess; };
        if (!ok(cudaSetDevice(c->cfg.device))) return;
        cudaStream_t st = c->lane_st[t];
        if (!ok(cudaStreamWaitEvent(st, c->lane_begin, 0))) return;
        const size_t lo = std::min(bytes, (size_t)t * part), hi = std::min(bytes, (size_t)(t + 1) * part);
        char* h = (char*)host + lo;
        char* d = (char*)dev + lo;
        const size_t n = hi - lo, pieces = (n + PIECE - 1) / PIECE;
        for (size_t k = 0; k < pieces + (to_device ? 0 : 1); ++k) {
            const int b = (int)(k & 1);
            const size_t off = k * PIECE, len = k < pieces ? std::min(PIECE, n - off) : 0;
            if (to_device) {
                if (!ok(cudaEventSynchronize(c->bounce_done[t][b]))) return;     // the buffer's previous piece has left it
                std::memcpy(c->bounce[t][b], h + off, len);
                if (!ok(cudaMemcpyAsync(d + off, c->bounce[t][b], len, cudaMemcpyHostToDevice, st))) return;
                if (!ok(cudaEventRecord(c->bounce_done[t][b], st))) return;
            } else {
                // piece k is requested, then piece k - 1 (in the other buffer) is waited for and copied out
                if (k < pieces) {
                    if (!ok(cudaMemcpyAsync(c->bounce[t][b], d + off, len, cudaMemcpyDeviceToHost, st))) return;
                    if (!ok(cudaEventRecord(c->bounce_done[t][b], st))) return;
                }
                if (k >= 1) {
                    const size_t poff = (k - 1) * PIECE, plen = std::min(PIECE, n - poff);
                    if (!ok(cudaEventSynchronize(c->bounce_done[t][b ^ 1]))) return;
                    std::memcpy(h + poff, c->bounce[t][b ^ 1], plen);
                }
            }
        }
        // uploads: the buffers are reused by the next call only after these events; the context's stream waits for them
        if (to_device)
            for (int b = 0; b < 2; ++b) if (!ok(cudaStreamWaitEvent(c->st.s, c->bounce_done[t][b], 0))) return;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < lanes; ++t) pool.emplace_back(lane, t);
    lane(0);
    for (auto& th : pool) th.join();
    for (int t = 0; t < lanes; ++t) if (status[t]) return status[t];
    return 0;
}
#else
static int staged_copy(astrea_ctx* c, void* dev, void* host, size_t bytes, bool to_device) {
    return to_device ? copy_h2d(dev, host, bytes, c->st) : copy_d2h(host, dev, bytes, c->st);
}
#endif

int astrea_upload(astrea_ctx* c, const double* grid_aos) {
    if (!c || !grid_aos) return fail(c, ASTREA_E_ARG, "astrea_upload: NULL argument");
    ASTREA_ON_DEVICE(c);
    const size_t bytes = (size_t)c->nrow * c->ncol * NVAR * sizeof(double);
    double* staging = c->d0.mem;     // d0 is scratch between operator evaluations
    ASTREA_TRY(staged_copy(c, staging, const_cast<double*>(grid_aos), bytes, true));
    ASTREA_TRY(dev_zero(c->mhd_flag, sizeof(int), c->st));
    ASTREA_TRY(dev_zero(c->flag, sizeof(unsigned long long), c->st));      // a new grid starts with a clean non-finite flag
    PackParams p{c->regs[c->grid_reg].plane, staging, c->nrow, c->ncol, 1, c->mhd_flag};
    { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<PackKernel>(p, (int)((c->ncol + 255) / 256), (int)c->nrow, 256, 0, c->st)); }
    c->next_instr = 0;
    int has_field = 1;
    ASTREA_TRY(copy_d2h(&has_field, c->mhd_flag, sizeof(int), c->st));
    if (stream_sync(c->st) != 0) return fail(c, ASTREA_E_CUDA, "astrea_upload: stream sync failed");
    // 2D hydro with LLF / HLLC: v_z and B stay identically zero, the kernels skip them (bit-identical results)
    c->field_free = !has_field;
    c->hydro = !has_field && c->cfg.dimension == 2 && !c->cfg.magnetic_2d && c->cfg.solver != SOL_HLLD && !(c->cfg.flags & 1);
    return 0;
}

int astrea_init_piecewise(astrea_ctx* c, const astrea_init_spec* spec) { return astrea_init_profiles(c, spec, 0, nullptr); }

int astrea_init_profiles(astrea_ctx* c, const astrea_init_spec* spec, int nprofiles, const astrea_init_profile* profiles) {
    if (!c || !spec) return fail(c, ASTREA_E_ARG, "astrea_init_piecewise: NULL argument");
    ASTREA_ON_DEVICE(c);
    if (nprofiles < 0 || nprofiles > ASTREA_MAX_PROFILES || (nprofiles > 0 && !profiles))
        return fail(c, ASTREA_E_ARG, "astrea_init_profiles: at most 4 profiles");
    if (c->cfg.dimension != 2) return fail(c, ASTREA_E_ARG, "astrea_init_piecewise: 2D grids only");
    if (spec->cells < 1 || spec->cells != c->ncol) return fail(c, ASTREA_E_ARG, "astrea_init_piecewise: spec.cells must equal ny");
    if (spec->nregions < 0 || spec->nregions > ASTREA_MAX_REGIONS) return fail(c, ASTREA_E_ARG, "astrea_init_piecewise: too many regions");
    InitParams p{};
    p.out = c->regs[c->grid_reg].plane;
    p.nrow = c->nrow; p.ncol = c->ncol; p.x_off = c->cfg.x_offset; p.n = spec->cells;
    p.start = spec->start; p.step = spec->step; p.gamma = c->cfg.gamma;
    p.bc = c->cfg.boundary; p.high_order = scheme_high_order(c->cfg.scheme) ? 1 : 0; p.nregions = spec->nregions;
    for (int v = 0; v < NVAR; ++v) p.state[0][v] = spec->background[v];
    for (int k = 0; k < spec->nregions; ++k) {
        if (spec->regions[k].kind < ASTREA_REGION_X_LT || spec->regions[k].kind > ASTREA_REGION_DISC_LE)
            return fail(c, ASTREA_E_ARG, "astrea_init_piecewise: unknown region kind");
        p.kind[k] = spec->regions[k].kind; p.a[k] = spec->regions[k].a; p.b[k] = spec->regions[k].b;
        for (int v = 0; v < NVAR; ++v) p.state[k + 1][v] = spec->regions[k].state[v];
    }
    p.mhd_flag = c->mhd_flag;
    // the 1-D tables travel through the start of a scratch plane (cells doubles each; the planes are far larger)
    p.nprofiles = nprofiles;
    for (int k = 0; k < nprofiles; ++k) {
        if (profiles[k].variable < 0 || profiles[k].variable >= NVAR || (profiles[k].along != 0 && profiles[k].along != 1) || !profiles[k].values)
            return fail(c, ASTREA_E_ARG, "astrea_init_profiles: bad profile");
        double* dst = c->d0.mem + (size_t)k * spec->cells;
        if ((size_t)(k + 1) * spec->cells > c->plane_doubles) return fail(c, ASTREA_E_ARG, "astrea_init_profiles: grid too small for the tables");
        ASTREA_TRY(copy_h2d(dst, profiles[k].values, (size_t)spec->cells * sizeof(double), c->st));
        p.prof_var[k] = profiles[k].variable; p.prof_along[k] = profiles[k].along; p.prof_tab[k] = dst;
    }
    ASTREA_TRY(dev_zero(c->mhd_flag, sizeof(int), c->st));
    ASTREA_TRY(dev_zero(c->flag, sizeof(unsigned long long), c->st));
    { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<InitKernel>(p, (int)((c->ncol + 127) / 128), (int)c->nrow, 128, 0, c->st)); }
    c->next_instr = 0;
    int has_field = 1;
    ASTREA_TRY(copy_d2h(&has_field, c->mhd_flag, sizeof(int), c->st));
    if (stream_sync(c->st) != 0) return fail(c, ASTREA_E_CUDA, "astrea_init_piecewise: stream sync failed");
    c->field_free = !has_field;        // as astrea_upload: 2D hydro states take the four-variable kernels
    c->hydro = !has_field && c->cfg.dimension == 2 && !c->cfg.magnetic_2d && c->cfg.solver != SOL_HLLD && !(c->cfg.flags & 1);
    return 0;
}

int astrea_download(astrea_ctx* c, double* grid_aos, int as_primitive) {
    if (!c || !grid_aos) return fail(c, ASTREA_E_ARG, "astrea_download: NULL argument");
    ASTREA_ON_DEVICE(c);
    if (c->next_instr != 0) return fail(c, ASTREA_E_STATE, "astrea_download: a step is in flight (between evolve_space and evolve_time the scratch planes are live)");
    Plane src = c->regs[c->grid_reg].plane;
    if (as_primitive) {
        if (as_primitive != 2 && c->slab() && scheme_high_order(c->cfg.scheme))
            return fail(c, ASTREA_E_STATE, "astrea_download: the 4th-order primitive snapshot of a slab reads ghost rows: exchange those of instruction 0 first and pass as_primitive = 2");
        if (int e = fill_halo(c, src, as_primitive == 2 ? 1 : 0)) return e;
        Plane w = make_plane(c->qT.mem, c->ncol, c->ghost_r);
        PrimParams pp{src, w, c->nrow, c->ncol, c->cfg.dimension, scheme_high_order(c->cfg.scheme) ? 1 : 0, c->cfg.gamma};
        { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<PrimKernel>(pp, (int)((c->ncol + 255) / 256), (int)c->nrow, 256, 0, c->st)); }
        src = w;
    }
    const size_t bytes = (size_t)c->nrow * c->ncol * NVAR * sizeof(double);
    double* staging = c->d0.mem;
    PackParams p{src, staging, c->nrow, c->ncol, 0, nullptr};
    { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<PackKernel>(p, (int)((c->ncol + 255) / 256), (int)c->nrow, 256, 0, c->st)); }
    ASTREA_TRY(staged_copy(c, staging, grid_aos, bytes, false));
    return stream_sync(c->st) == 0 ? 0 : fail(c, ASTREA_E_CUDA, "astrea_download: stream sync failed");
}

int64_t astrea_snapshot_begin(astrea_ctx* c, double* host_dst, int external_rows) {
    if (!c || !host_dst) return fail(c, ASTREA_E_ARG, "astrea_snapshot_begin: NULL argument");
    ASTREA_ON_DEVICE(c);
    if (c->next_instr != 0) return fail(c, ASTREA_E_STATE, "astrea_snapshot_begin: a step is in flight");
    if (!external_rows && c->slab() && scheme_high_order(c->cfg.scheme))
        return fail(c, ASTREA_E_STATE, "astrea_snapshot_begin: the 4th-order primitive snapshot of a slab reads ghost rows: exchange them first and pass external_rows = 1");
    const size_t bytes = (size_t)c->nrow * c->ncol * NVAR * sizeof(double);
    if (!c->snap_dev) {
        c->snap_dev = (double*)dev_alloc(bytes);
        if (!c->snap_dev) return fail(c, ASTREA_E_CUDA, "astrea_snapshot_begin: device allocation failed");
#ifdef ASTREA_DEVICE_BUILD
        cudaError_t e = cudaStreamCreateWithFlags(&c->copy_st, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->snap_ready, cudaEventDisableTiming);
        for (auto& ev : c->snap_done) if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (e != cudaSuccess) return fail(c, ASTREA_E_CUDA, std::string("astrea_snapshot_begin: ") + cudaGetErrorString(e));
#endif
    }
    const int64_t ticket = c->snap_tickets;
#ifdef ASTREA_DEVICE_BUILD
    // the staging buffer is free once the previous snapshot has left the device; the ticket's event slot once its
    // previous user (SNAP_RING tickets ago) has been waited for by the host or has simply completed
    if (ticket > 0) ASTREA_TRY((int)cudaStreamWaitEvent(c->st.s, c->snap_done[(ticket - 1) % astrea_ctx::SNAP_RING], 0));
#endif
    Plane src = c->regs[c->grid_reg].plane;
    if (int e = fill_halo(c, src, external_rows)) return e;
    Plane w = make_plane(c->qT.mem, c->ncol, c->ghost_r);
    PrimParams pp{src, w, c->nrow, c->ncol, c->cfg.dimension, scheme_high_order(c->cfg.scheme) ? 1 : 0, c->cfg.gamma};
    { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<PrimKernel>(pp, (int)((c->ncol + 255) / 256), (int)c->nrow, 256, 0, c->st)); }
    if (c->cfg.dimension == 2) {
        PackTransposedParams tp{w, c->snap_dev, c->nrow, c->ncol};
        Timed timed(c, CLS_TRANSPOSE);
        ASTREA_TRY(launch<PackTransposedKernel>(tp, (int)((c->ncol + 31) / 32), (int)((c->nrow + 31) / 32), 256, PackTransposedKernel::smem_bytes(), c->st));
    } else {
        PackParams p{w, c->snap_dev, c->nrow, c->ncol, 0, nullptr};
        Timed timed(c, CLS_HALO);
        ASTREA_TRY(launch<PackKernel>(p, (int)((c->ncol + 255) / 256), (int)c->nrow, 256, 0, c->st));
    }
#ifdef ASTREA_DEVICE_BUILD
    ASTREA_TRY((int)cudaEventRecord(c->snap_ready, c->st.s));
    ASTREA_TRY((int)cudaStreamWaitEvent(c->copy_st, c->snap_ready, 0));
    ASTREA_TRY((int)cudaMemcpyAsync(host_dst, c->snap_dev, bytes, cudaMemcpyDeviceToHost, c->copy_st));
    ASTREA_TRY((int)cudaEventRecord(c->snap_done[ticket % astrea_ctx::SNAP_RING], c->copy_st));
#else
    std::memcpy(host_dst, c->snap_dev, bytes);
#endif
    c->snap_tickets = ticket + 1;
    return ticket;
}

int astrea_snapshot_wait(astrea_ctx* c, int64_t ticket) {
    if (!c) return ASTREA_E_ARG;
    ASTREA_ON_DEVICE(c);
    if (ticket < 0 || ticket >= c->snap_tickets) return fail(c, ASTREA_E_ARG, "astrea_snapshot_wait: no such ticket");
    if (ticket + astrea_ctx::SNAP_RING < c->snap_tickets) return 0;        // its event slot has been reused: long done (stream order)
#ifdef ASTREA_DEVICE_BUILD
    if (cudaEventSynchronize(c->snap_done[ticket % astrea_ctx::SNAP_RING]) != cudaSuccess) return fail(c, ASTREA_E_CUDA, "astrea_snapshot_wait: event sync failed");
#endif
    return 0;
}

int astrea_diagnostics(astrea_ctx* c, double* totals, double* total_variation, int external_rows) {
    if (!c || !totals || !total_variation) return fail(c, ASTREA_E_ARG, "astrea_diagnostics: NULL argument");
    ASTREA_ON_DEVICE(c);
    if (c->next_instr != 0) return fail(c, ASTREA_E_STATE, "astrea_diagnostics: a step is in flight");
    Plane q = c->regs[c->grid_reg].plane;
    if (int e = fill_halo(c, q, external_rows)) return e;
    Plane w = make_plane(c->qT.mem, c->ncol, c->ghost_r);
    PrimParams pp{q, w, c->nrow, c->ncol, c->cfg.dimension, scheme_high_order(c->cfg.scheme) ? 1 : 0, c->cfg.gamma};
    { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<PrimKernel>(pp, (int)((c->ncol + 255) / 256), (int)c->nrow, 256, 0, c->st)); }
    const int gx = (int)((c->ncol + 255) / 256), gy = (int)c->nrow;
    const size_t n = (size_t)gx * gy * 2 * NVAR;
    double* partial = c->d0.mem;          // scratch between steps
    DiagParams dp{q, w, c->nrow, c->ncol, c->cfg.dimension, partial};
    { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<DiagKernel>(dp, gx, gy, 256, DiagKernel::smem_bytes(), c->st)); }
    std::vector<double> host(n);
    ASTREA_TRY(copy_d2h(host.data(), partial, n * sizeof(double), c->st));
    if (stream_sync(c->st) != 0) return fail(c, ASTREA_E_CUDA, "astrea_diagnostics: stream sync failed");
    for (int k = 0; k < NVAR; ++k) { totals[k] = 0.0; total_variation[k] = 0.0; }
    for (size_t b = 0; b < (size_t)gx * gy; ++b)
        for (int k = 0; k < NVAR; ++k) {
            totals[k] += host[b * 2 * NVAR + k];
            total_variation[k] += host[b * 2 * NVAR + NVAR + k];
        }
    return 0;
}

// schemes/ppm.py:111-170 at function level (aux_kernels.cuh DissipationKernel)
static int run_dissipation(astrea_ctx* c, const double* ws_aos, int axis, int what, const double* knobs, double* out) {
    if (!ws_aos || !out) return fail(c, ASTREA_E_ARG, "ppm dissipation: NULL argument");
    if (c->next_instr != 0) return fail(c, ASTREA_E_STATE, "ppm dissipation: a step is in flight");
    if (axis < 0 || axis >= c->cfg.dimension) return fail(c, ASTREA_E_ARG, "ppm dissipation: axis must be below the dimension");
    if (c->slab()) return fail(c, ASTREA_E_ARG, "ppm dissipation: whole grids only");
    if (what == 1 && c->cfg.dimension != 1)
        return fail(c, ASTREA_E_ARG, "apply_artificial_viscosity: the reference's 2D branch raises (ppm.py:154-156: operands could not be "
                                     "broadcast together); available in 1D");
    const size_t bytes = (size_t)c->nrow * c->ncol * NVAR * sizeof(double);
    // wS -> a scratch plane (ghost cells = np.pad of wS), result -> another one
    Plane w = make_plane(c->qT.mem, c->ncol, c->ghost_r), res = c->rates[0].plane;
    double* staging = c->d0.mem;
    ASTREA_TRY(copy_h2d(staging, ws_aos, bytes, c->st));
    PackParams pk{w, staging, c->nrow, c->ncol, 1, nullptr};
    { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<PackKernel>(pk, (int)((c->ncol + 255) / 256), (int)c->nrow, 256, 0, c->st)); }
    {
        HaloParams h{w, c->nrow, c->ncol, c->cfg.boundary, 0, 1, 1, all_vars(), 0, 0, 0};
        { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<HaloKernel>(h, 1, (int)c->nrow, 64, 0, c->st)); }
        if (c->ghost_r > 0) {
            h.phase = 1;
            Timed timed(c, CLS_HALO);
            ASTREA_TRY(launch<HaloKernel>(h, (int)((c->ncol + 2 * GHOST + 255) / 256), 2 * GHOST * NVAR, 256, 0, c->st));
        }
    }
    DissipationParams dp{};
    dp.w = w; dp.out = res; dp.nrow = c->nrow; dp.ncol = c->ncol; dp.dimension = c->cfg.dimension; dp.axis = axis;
    dp.bc = c->cfg.boundary; dp.what = what; dp.gamma = c->cfg.gamma; dp.dx = c->cfg.dx;
    if (what == 0) { dp.delta = knobs[0]; dp.z0 = knobs[1]; dp.z1 = knobs[2]; } else { dp.alpha = knobs[0]; dp.beta = knobs[1]; }
    { Timed timed(c, CLS_RECON); ASTREA_TRY(launch<DissipationKernel>(dp, (int)((c->ncol + 127) / 128), (int)c->nrow, 128, 0, c->st)); }
    if (what == 0) {
        ASTREA_TRY(copy_d2h_2d(out, (size_t)c->ncol * sizeof(double), res.at(0, 0, 0), (size_t)res.row_pitch * sizeof(double),
                               (size_t)c->ncol * sizeof(double), (size_t)c->nrow, c->st));
    } else {
        PackParams up{res, staging, c->nrow, c->ncol, 0, nullptr};
        { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<PackKernel>(up, (int)((c->ncol + 255) / 256), (int)c->nrow, 256, 0, c->st)); }
        ASTREA_TRY(copy_d2h(out, staging, bytes, c->st));
    }
    return stream_sync(c->st) == 0 ? 0 : fail(c, ASTREA_E_CUDA, "ppm dissipation: stream sync failed");
}

int astrea_ppm_flattener(astrea_ctx* c, const double* ws_aos, int axis, const double* slope_determinants, double* chi) {
    if (!c) return ASTREA_E_ARG;
    ASTREA_ON_DEVICE(c);
    const double standard[3] = {.33, .75, .85};                      // ppm.py:112
    return run_dissipation(c, ws_aos, axis, 0, slope_determinants ? slope_determinants : standard, chi);
}

int astrea_ppm_viscosity(astrea_ctx* c, const double* ws_aos, int axis, const double* viscosity_determinants, double* mu_aos) {
    if (!c) return ASTREA_E_ARG;
    ASTREA_ON_DEVICE(c);
    const double standard[2] = {.3, .3};                             // ppm.py:139
    return run_dissipation(c, ws_aos, axis, 1, viscosity_determinants ? viscosity_determinants : standard, mu_aos);
}

int astrea_solution_error(astrea_ctx* c, const double* w_theo_aos, double norm, double* error, int external_rows) {
    if (!c || !w_theo_aos || !error) return fail(c, ASTREA_E_ARG, "astrea_solution_error: NULL argument");
    ASTREA_ON_DEVICE(c);
    if (c->next_instr != 0) return fail(c, ASTREA_E_STATE, "astrea_solution_error: a step is in flight");
    Plane q = c->regs[c->grid_reg].plane;
    if (int e = fill_halo(c, q, external_rows)) return e;
    // numerical primitives -> qT (as astrea_download(as_primitive)), theoretical primitives -> rates[0]
    Plane w = make_plane(c->qT.mem, c->ncol, c->ghost_r), theo = c->rates[0].plane;
    PrimParams pp{q, w, c->nrow, c->ncol, c->cfg.dimension, scheme_high_order(c->cfg.scheme) ? 1 : 0, c->cfg.gamma};
    { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<PrimKernel>(pp, (int)((c->ncol + 255) / 256), (int)c->nrow, 256, 0, c->st)); }
    const size_t bytes = (size_t)c->nrow * c->ncol * NVAR * sizeof(double);
    double* staging = c->d0.mem;
    ASTREA_TRY(copy_h2d(staging, w_theo_aos, bytes, c->st));
    PackParams pk{theo, staging, c->nrow, c->ncol, 1, nullptr};
    { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<PackKernel>(pk, (int)((c->ncol + 255) / 256), (int)c->nrow, 256, 0, c->st)); }
    const int gx = (int)((c->ncol + 255) / 256), gy = (int)c->nrow;
    const size_t n = (size_t)gx * gy * ERR_CHANNELS;
    double* partial = staging;          // the staging copy has been consumed (stream order)
    ErrorParams ep{w, theo, c->nrow, c->ncol, c->cfg.gamma, norm, partial};
    { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<ErrorKernel>(ep, gx, gy, 256, ErrorKernel::smem_bytes(), c->st)); }
    std::vector<double> host(n);
    ASTREA_TRY(copy_d2h(host.data(), partial, n * sizeof(double), c->st));
    if (stream_sync(c->st) != 0) return fail(c, ASTREA_E_CUDA, "astrea_solution_error: stream sync failed");
    const bool use_max = norm > 10.0;
    for (int k = 0; k < ERR_CHANNELS; ++k) error[k] = 0.0;
    for (size_t b = 0; b < (size_t)gx * gy; ++b)
        for (int k = 0; k < ERR_CHANNELS; ++k) {
            const double x = host[b * ERR_CHANNELS + k];
            error[k] = use_max ? npmax(error[k], x) : error[k] + x;
        }
    return 0;
}

int astrea_fp64_probe(astrea_ctx* c, double* tflops) {
    if (!c || !tflops) return fail(c, ASTREA_E_ARG, "astrea_fp64_probe: NULL argument");
    ASTREA_ON_DEVICE(c);
    *tflops = 0.0;
#ifdef ASTREA_DEVICE_BUILD
    if (c->next_instr != 0) return fail(c, ASTREA_E_STATE, "astrea_fp64_probe: a step is in flight");
    const int blocks = 148 * 8, threads = 256, iters = 1 << 14;
    if ((size_t)blocks * threads > c->plane_doubles) return 0;      // tiny grid: no scratch to write to
    Fp64ProbeParams p{c->d0.mem, iters, 1.0000001, 1e-9};
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    ASTREA_TRY(launch<Fp64ProbeKernel>(p, blocks, 1, threads, 0, c->st));       // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a, c->st.s);
        ASTREA_TRY(launch<Fp64ProbeKernel>(p, blocks, 1, threads, 0, c->st));
        cudaEventRecord(b, c->st.s);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *tflops = 2.0 * 8.0 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
#endif
    return 0;
}

int astrea_arith_check(astrea_ctx* c, int64_t samples, uint64_t seed, uint64_t* counts) {
    if (!c || !counts || samples < 1) return fail(c, ASTREA_E_ARG, "astrea_arith_check: bad argument");
    ASTREA_ON_DEVICE(c);
    if (c->next_instr != 0) return fail(c, ASTREA_E_STATE, "astrea_arith_check: a step is in flight");
    unsigned long long* dev = (unsigned long long*)dev_alloc(3 * sizeof(unsigned long long));
    if (!dev) return fail(c, ASTREA_E_CUDA, "astrea_arith_check: allocation failed");
    const int threads = 256, per_thread = 256;
    const int blocks = (int)((samples + (int64_t)threads * per_thread - 1) / ((int64_t)threads * per_thread));
    int e = dev_zero(dev, 3 * sizeof(unsigned long long), c->st);
    ArithCheckParams p{dev, (unsigned long long)seed, per_thread};
    if (!e) e = launch<ArithCheckKernel>(p, blocks, 1, threads, 0, c->st);
    unsigned long long host[3] = {0, 0, 0};
    if (!e) e = copy_d2h(host, dev, sizeof(host), c->st);
    if (!e) e = stream_sync(c->st);
    dev_free(dev);
    if (e) return fail(c, ASTREA_E_CUDA, "astrea_arith_check: launch failed");
    for (int k = 0; k < 3; ++k) counts[k] = host[k];
    return 0;
}

int astrea_set_flag_reducer(astrea_ctx* c, astrea_reduce_fn fn, void* user) {
    if (!c) return ASTREA_E_ARG;
    c->reduce_fn = fn;
    c->reduce_user = user;
    return 0;
}

int astrea_set_key_reducer(astrea_ctx* c, astrea_reduce_fn fn, void* user) {
    if (!c) return ASTREA_E_ARG;
    c->key_reduce_fn = fn;
    c->key_reduce_user = user;
    return 0;
}

int astrea_program_length(const astrea_ctx* c) { return c ? (int)c->prog.size() : ASTREA_E_ARG; }

int astrea_instr_is_operator(const astrea_ctx* c, int i) {
    if (!c || i < 0 || i >= (int)c->prog.size()) return ASTREA_E_ARG;
    return c->prog[i].is_operator;
}

int astrea_set_dt(astrea_ctx* c, double dt) {
    if (!c) return ASTREA_E_ARG;
    ASTREA_ON_DEVICE(c);
    ASTREA_TRY(copy_h2d(c->dt_dev, &dt, sizeof(double), c->st));
    // the source is a stack variable: finish the copy before returning
    return stream_sync(c->st) == 0 ? 0 : fail(c, ASTREA_E_CUDA, "astrea_set_dt: stream sync failed");
}

int astrea_run_instr(astrea_ctx* c, int i, int external_rows) {
    if (!c || i < 0 || i >= (int)c->prog.size()) return fail(c, ASTREA_E_ARG, "astrea_run_instr: bad instruction index");
    ASTREA_ON_DEVICE(c);
    if (i != c->next_instr) return fail(c, ASTREA_E_STATE, "astrea_run_instr: instructions must run in order (expected " + std::to_string(c->next_instr) + ")");
    const Instr& ins = c->prog[i];
    const int e = ins.is_operator ? run_operator(c, ins, external_rows, i == 0)
                                  : (ins.special != SP_NONE ? run_special(c, ins, external_rows) : run_combine(c, ins, 0, c->nrow));
    if (e) return e;
    c->next_instr = i + 1;
    return 0;
}

int astrea_instr_needs_halo(const astrea_ctx* c, int i) {
    if (!c || i < 0 || i >= (int)c->prog.size()) return ASTREA_E_ARG;
    return needs_ghost_rows(c->prog[i]) ? 1 : 0;
}

int astrea_instr_is_update(const astrea_ctx* c, int i) {
    if (!c || i < 0 || i >= (int)c->prog.size()) return ASTREA_E_ARG;
    return (!c->prog[i].is_operator && c->prog[i].special == SP_NONE) ? 1 : 0;
}

int astrea_run_update_part(astrea_ctx* c, int i, int part) {
    if (!c || i < 0 || i >= (int)c->prog.size() || astrea_instr_is_update(c, i) != 1)
        return fail(c, ASTREA_E_ARG, "astrea_run_update_part: not a register update");
    ASTREA_ON_DEVICE(c);
    if (i != c->next_instr) return fail(c, ASTREA_E_STATE, "astrea_run_update_part: instructions must run in order (expected " + std::to_string(c->next_instr) + ")");
    // the first / last UPDATE_EDGE rows are what the neighbours' ghost rows are made of (GHOST <= UPDATE_EDGE); a slab
    // too short for two disjoint edge blocks of at least GHOST rows is updated whole in part 0
    const Instr& ins = c->prog[i];
    const bool whole = c->nrow < 2 * (int64_t)GHOST;
    const int64_t edge = whole ? c->nrow : std::min<int64_t>(UPDATE_EDGE, c->nrow / 2);
    if (part == 0) {
        if (int e = run_combine(c, ins, 0, edge)) return e;
        return whole ? 0 : run_combine(c, ins, std::max<int64_t>(edge, c->nrow - edge), c->nrow);
    }
    if (!whole)
        if (int e = run_combine(c, ins, edge, c->nrow - edge)) return e;
    c->next_instr = i + 1;
    return 0;
}

int astrea_finish_step(astrea_ctx* c) {
    if (!c) return ASTREA_E_ARG;
    if (c->next_instr != (int)c->prog.size()) return fail(c, ASTREA_E_STATE, "astrea_finish_step: the step program has not run to its end");
    c->grid_reg = c->final_reg;
    c->parity ^= 1;
    c->next_instr = 0;
    return 0;
}

int astrea_read_eigmax(astrea_ctx* c, double* eigmax) {
    if (!c || !eigmax) return fail(c, ASTREA_E_ARG, "astrea_read_eigmax: NULL argument");
    ASTREA_ON_DEVICE(c);
    unsigned long long bits[3] = {0, 0, 0};
    ASTREA_TRY(copy_d2h(bits, c->eig_bits, sizeof(bits), c->st));
    if (stream_sync(c->st) != 0) return fail(c, ASTREA_E_CUDA, "astrea_read_eigmax: stream sync failed");
    const bool flag = bits[2] != 0;
    for (int a = 0; a < c->cfg.dimension; ++a) {
        const int slot = c->cfg.dimension == 1 ? 0 : a;
        std::memcpy(&eigmax[a], &bits[slot], sizeof(double));
    }
    if (flag) {     // sticky until read: reported once, then cleared
        dev_zero(c->flag, sizeof(unsigned long long), c->st);
        return fail(c, ASTREA_E_NONFINITE, "non-finite wave speed (the reference raises LinAlgError: Array must not contain infs or NaNs, fv.py:158)");
    }
    return 0;
}

int astrea_evolve_space(astrea_ctx* c, int step_parity, double* eigmax) {
    if (!c) return ASTREA_E_ARG;
    ASTREA_ON_DEVICE(c);
    c->parity = step_parity & 1;
    c->next_instr = 0;
    if (int e = astrea_run_instr(c, 0, 0)) return e;
    return astrea_read_eigmax(c, eigmax);
}

int astrea_evolve_time(astrea_ctx* c, double dt) {
    if (!c) return ASTREA_E_ARG;
    ASTREA_ON_DEVICE(c);
    if (c->next_instr != 1) return fail(c, ASTREA_E_STATE, "astrea_evolve_time: call astrea_evolve_space first");
    if (int e = astrea_set_dt(c, dt)) return e;
    for (int i = 1; i < (int)c->prog.size(); ++i)
        if (int e = astrea_run_instr(c, i, 0)) return e;
    // astrea.py:81 rebinds grid; the permutation reversal (astrea.py:85) is the caller's (astrea_step does both)
    c->grid_reg = c->final_reg;
    c->next_instr = 0;
    unsigned long long flag = 0;
    ASTREA_TRY(copy_d2h(&flag, c->flag, sizeof(flag), c->st));
    if (stream_sync(c->st) != 0) return fail(c, ASTREA_E_CUDA, "astrea_evolve_time: stream sync failed");
    if (flag) {
        dev_zero(c->flag, sizeof(unsigned long long), c->st);
        return fail(c, ASTREA_E_NONFINITE, "non-finite wave speed in a Runge-Kutta stage (fv.py:158 raises LinAlgError)");
    }
    return 0;
}

int astrea_step(astrea_ctx* c, double t, double t_stop, double* dt_out) {
    if (!c) return ASTREA_E_ARG;
    ASTREA_ON_DEVICE(c);
    double eig[2] = {0, 0};
    if (int e = astrea_evolve_space(c, c->parity, eig)) return e;
    double dt = c->cfg.cfl * (c->cfg.dx / eig[0]);
    if (c->cfg.dimension == 2) dt = std::min(dt, c->cfg.cfl * (c->cfg.dx / eig[1]));
    if (t_stop > t && t + dt >= t_stop) dt = t_stop - t;
    if (dt_out) *dt_out = dt;
    if (int e = astrea_set_dt(c, dt)) return e;
    for (int i = 1; i < (int)c->prog.size(); ++i)
        if (int e = astrea_run_instr(c, i, 0)) return e;
    return astrea_finish_step(c);
}

int astrea_set_time(astrea_ctx* c, double t, double t_stop) {
    if (!c) return ASTREA_E_ARG;
    ASTREA_ON_DEVICE(c);
    ClockParams k{c->clock, c->dt_dev, c->eig_bits, c->dt_history, DT_HISTORY, 0, c->cfg.dimension, c->cfg.cfl, c->cfg.dx, t, t_stop};
    Timed timed(c, CLS_HALO);
    ASTREA_TRY(launch<ClockKernel>(k, 1, 1, 32, 0, c->st));
    return 0;
}

int astrea_dt_async(astrea_ctx* c) {
    if (!c) return ASTREA_E_ARG;
    ASTREA_ON_DEVICE(c);
    if (c->next_instr != 1) return fail(c, ASTREA_E_STATE, "astrea_dt_async: run instruction 0 (the operator on the grid) first");
    ClockParams k{c->clock, c->dt_dev, c->eig_bits, c->dt_history, DT_HISTORY, 1, c->cfg.dimension, c->cfg.cfl, c->cfg.dx, 0.0, 0.0};
    Timed timed(c, CLS_HALO);
    ASTREA_TRY(launch<ClockKernel>(k, 1, 1, 32, 0, c->st));
    return 0;
}

static int enqueue_step(astrea_ctx* c) {
    c->next_instr = 0;
    if (int e = astrea_run_instr(c, 0, 0)) return e;
    if (int e = astrea_dt_async(c)) return e;
    for (int i = 1; i < (int)c->prog.size(); ++i)
        if (int e = astrea_run_instr(c, i, 0)) return e;
    return 0;
}

int astrea_step_async(astrea_ctx* c) {
    if (!c) return ASTREA_E_ARG;
    ASTREA_ON_DEVICE(c);
#ifdef ASTREA_DEVICE_BUILD
    // Launch-bound regime (the 1D configurations, small 2D grids): the ~10-40 launches of a step are captured once
    // per (step parity, hydro) and replayed as one CUDA graph.  Large grids gain nothing and launch eagerly.
    const bool small = (int64_t)c->nrow * c->ncol <= GRAPH_MAX_CELLS && !(c->cfg.flags & 2) && !c->profiling;
    if (small) {
        const int a = c->parity & 1, b = c->hydro ? 1 : 0;
        if (c->step_graph[a][b] == nullptr && c->eager_steps[a][b] >= 1) {     // first step eager: sets kernel attributes
            cudaGraph_t graph = nullptr;
            cudaError_t err = cudaStreamBeginCapture(c->st.s, cudaStreamCaptureModeThreadLocal);
            if (err != cudaSuccess) return fail(c, ASTREA_E_CUDA, std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(err));
            const int64_t launches0 = c->launches;
            const int e = enqueue_step(c);
            err = cudaStreamEndCapture(c->st.s, &graph);
            if (e) { if (graph) cudaGraphDestroy(graph); return e; }
            if (err != cudaSuccess) return fail(c, ASTREA_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(err));
            err = cudaGraphInstantiate(&c->step_graph[a][b], graph, 0);
            cudaGraphDestroy(graph);
            if (err != cudaSuccess) return fail(c, ASTREA_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(err));
            c->graph_launches[a][b] = c->launches - launches0;
            c->launches = launches0;
        }
        if (c->step_graph[a][b] != nullptr) {
            const cudaError_t err = cudaGraphLaunch(c->step_graph[a][b], c->st.s);
            if (err != cudaSuccess) return fail(c, ASTREA_E_CUDA, std::string("cudaGraphLaunch: ") + cudaGetErrorString(err));
            c->launches += c->graph_launches[a][b];
            c->next_instr = (int)c->prog.size();
            return astrea_finish_step(c);
        }
        c->eager_steps[a][b]++;
    }
#endif
    if (int e = enqueue_step(c)) return e;
    return astrea_finish_step(c);
}

int astrea_run_steps(astrea_ctx* c, int64_t nsteps) {
    if (!c || nsteps < 0) return fail(c, ASTREA_E_ARG, "astrea_run_steps: bad argument");
    for (int64_t k = 0; k < nsteps; ++k)
        if (int e = astrea_step_async(c)) return e;
    return 0;
}

int astrea_get_time(astrea_ctx* c, double* t, int64_t* steps, double* last_dt) {
    if (!c) return ASTREA_E_ARG;
    ASTREA_ON_DEVICE(c);
    double clock[4] = {0, 0, 0, 0};
    unsigned long long flag = 0;
    ASTREA_TRY(copy_d2h(clock, c->clock, sizeof(clock), c->st));
    ASTREA_TRY(copy_d2h(&flag, c->flag, sizeof(flag), c->st));
    if (stream_sync(c->st) != 0) return fail(c, ASTREA_E_CUDA, "astrea_get_time: stream sync failed");
    if (t) *t = clock[0];
    if (steps) *steps = (int64_t)clock[2];
    if (last_dt) *last_dt = clock[3];
    if (flag) {
        dev_zero(c->flag, sizeof(unsigned long long), c->st);
        return fail(c, ASTREA_E_NONFINITE, "non-finite wave speed in a step since the last check (fv.py:158 raises LinAlgError)");
    }
    return 0;
}

int astrea_dt_history(astrea_ctx* c, double* out, int n) {
    if (!c || !out || n < 0 || n > DT_HISTORY) return fail(c, ASTREA_E_ARG, "astrea_dt_history: n must be within 0..1024");
    ASTREA_ON_DEVICE(c);
    double clock[4];
    std::vector<double> hist(DT_HISTORY);
    ASTREA_TRY(copy_d2h(clock, c->clock, sizeof(clock), c->st));
    ASTREA_TRY(copy_d2h(hist.data(), c->dt_history, DT_HISTORY * sizeof(double), c->st));
    if (stream_sync(c->st) != 0) return fail(c, ASTREA_E_CUDA, "astrea_dt_history: stream sync failed");
    const long long steps = (long long)clock[2];
    if (n > steps) return fail(c, ASTREA_E_ARG, "astrea_dt_history: fewer steps taken than requested");
    for (int k = 0; k < n; ++k) out[k] = hist[(size_t)((steps - n + k) % DT_HISTORY)];
    return 0;
}

int astrea_get_parity(const astrea_ctx* c) { return c ? c->parity : ASTREA_E_ARG; }
int astrea_set_parity(astrea_ctx* c, int p) { if (!c) return ASTREA_E_ARG; c->parity = p & 1; return 0; }

int astrea_download_face_field(astrea_ctx* c, double* bxy_aos) {
    if (!c || !bxy_aos) return fail(c, ASTREA_E_ARG, "astrea_download_face_field: NULL argument");
    ASTREA_ON_DEVICE(c);
    if (!c->cfg.magnetic_2d) return fail(c, ASTREA_E_ARG, "astrea_download_face_field: magnetic_2d is off");
    // face averages of the last operator, assembled in a scratch plane and unpacked on the host
    Plane tmp = make_plane(c->ws.mem, c->ncol, GHOST);
    FaceFieldParams fp{tmp, make_plane(c->wfx.mem, c->nrow, GHOST), make_plane(c->wfy.mem, c->ncol, GHOST), c->nrow, c->ncol};
    { Timed timed(c, CLS_UPDATE); ASTREA_TRY(launch<FaceFieldKernel>(fp, (int)((c->ncol + 31) / 32), (int)((c->nrow + 31) / 32), 256, FaceFieldKernel::smem_bytes(), c->st)); }
    const size_t n = (size_t)c->nrow * c->ncol;
    std::vector<double> host(n * NVAR);
    double* staging = c->wp.mem;
    PackParams p{tmp, staging, c->nrow, c->ncol, 0, nullptr};
    { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<PackKernel>(p, (int)((c->ncol + 255) / 256), (int)c->nrow, 256, 0, c->st)); }
    ASTREA_TRY(copy_d2h(host.data(), staging, n * NVAR * sizeof(double), c->st));
    if (stream_sync(c->st) != 0) return fail(c, ASTREA_E_CUDA, "astrea_download_face_field: stream sync failed");
    for (size_t k = 0; k < n; ++k) { bxy_aos[2 * k] = host[NVAR * k + 5]; bxy_aos[2 * k + 1] = host[NVAR * k + 6]; }
    return 0;
}

int astrea_halo_info(const astrea_ctx* c, int64_t* ghost_rows, int64_t* doubles_per_block) {
    if (!c) return ASTREA_E_ARG;
    if (ghost_rows) *ghost_rows = c->ghost_r;
    if (doubles_per_block) *doubles_per_block = (int64_t)c->ghost_r * NVAR * pitch_of(c->ncol);
    return 0;
}

int astrea_halo_ptrs(astrea_ctx* c, int i, double** send_lo, double** send_hi, double** recv_lo, double** recv_hi) {
    if (!c || i < 0 || i >= (int)c->prog.size() || !needs_ghost_rows(c->prog[i])) return fail(c, ASTREA_E_ARG, "astrea_halo_ptrs: the instruction reads no ghost rows");
    const Plane& p = c->regs[halo_register(c->prog[i])].plane;
    // whole padded rows (ghost columns included): the receiver's corner ghosts come along for free; the ghost
    // columns of the interior rows sent here are filled by astrea_halo_prepare()
    double* row0 = p.base - GHOST;
    if (send_lo) *send_lo = row0;                                                  // rows 0 .. G-1
    if (send_hi) *send_hi = row0 + (c->nrow - c->ghost_r) * p.row_pitch;           // rows n-G .. n-1
    if (recv_lo) *recv_lo = row0 - (int64_t)c->ghost_r * p.row_pitch;              // rows -G .. -1
    if (recv_hi) *recv_hi = row0 + c->nrow * p.row_pitch;                          // rows n .. n+G-1
    return 0;
}

int astrea_halo_prepare(astrea_ctx* c, int i) {
    if (!c || i < 0 || i >= (int)c->prog.size() || !needs_ghost_rows(c->prog[i])) return fail(c, ASTREA_E_ARG, "astrea_halo_prepare: the instruction reads no ghost rows");
    ASTREA_ON_DEVICE(c);
    // only the first / last ghost_r interior rows travel: fill their ghost columns
    HaloParams h{c->regs[halo_register(c->prog[i])].plane, c->nrow, c->ncol, c->cfg.boundary, 0, 0, 0, c->vars(), 0, 0, 0};
    const int rows = (int)std::min<int64_t>(c->ghost_r, c->nrow);
    { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<HaloKernel>(h, 1, rows, 64, 0, c->st)); }
    h.row0 = c->nrow - rows;
    { Timed timed(c, CLS_HALO); ASTREA_TRY(launch<HaloKernel>(h, 1, rows, 64, 0, c->st)); }
    return 0;
}

int astrea_eigmax_device(astrea_ctx* c, double** p) {
    if (!c || !p) return ASTREA_E_ARG;
    *p = reinterpret_cast<double*>(c->eig_bits);
    return 0;
}

int astrea_sync(astrea_ctx* c) {
    if (!c) return ASTREA_E_ARG;
    ASTREA_ON_DEVICE(c);
    return stream_sync(c->st) == 0 ? 0 : fail(c, ASTREA_E_CUDA, "astrea_sync: stream sync failed");
}

uint64_t astrea_stream_handle(const astrea_ctx* c) {
#ifdef ASTREA_DEVICE_BUILD
    return c ? (uint64_t)(uintptr_t)c->st.s : 0;
#else
    (void)c;
    return 0;
#endif
}

int64_t astrea_launch_count(const astrea_ctx* c) { return c ? c->launches : 0; }

int astrea_save_state(astrea_ctx* c) {
    if (!c) return ASTREA_E_ARG;
    ASTREA_ON_DEVICE(c);
    if (c->next_instr != 0) return fail(c, ASTREA_E_STATE, "astrea_save_state: a step is in flight");
    if (!c->saved.mem && !alloc_reg(c, c->saved, c->ncol)) return fail(c, ASTREA_E_CUDA, "astrea_save_state: device allocation failed");
    ASTREA_TRY(copy_d2d(c->saved.mem, c->regs[c->grid_reg].mem, c->plane_doubles * sizeof(double), c->st));
    c->saved_parity = c->parity;
    c->saved_hydro = c->hydro;
    c->saved_field_free = c->field_free;
    return 0;
}

int astrea_restore_state(astrea_ctx* c) {
    if (!c) return ASTREA_E_ARG;
    ASTREA_ON_DEVICE(c);
    if (!c->saved.mem) return fail(c, ASTREA_E_STATE, "astrea_restore_state: nothing saved");
    ASTREA_TRY(copy_d2d(c->regs[c->grid_reg].mem, c->saved.mem, c->plane_doubles * sizeof(double), c->st));
    ASTREA_TRY(dev_zero(c->flag, sizeof(unsigned long long), c->st));
    c->parity = c->saved_parity;
    c->hydro = c->saved_hydro;
    c->field_free = c->saved_field_free;
    c->next_instr = 0;
    return 0;
}

int astrea_profile(astrea_ctx* c, int enable) {
    if (!c) return ASTREA_E_ARG;
    c->profiling = enable ? 1 : 0;
    return 0;
}

int astrea_profile_read(astrea_ctx* c, double* ms_by_class, int64_t* launches_by_class) {
    if (!c || !ms_by_class || !launches_by_class) return fail(c, ASTREA_E_ARG, "astrea_profile_read: NULL argument");
    ASTREA_ON_DEVICE(c);
    for (int k = 0; k < CLS_COUNT; ++k) { ms_by_class[k] = 0.0; launches_by_class[k] = 0; }
#ifdef ASTREA_DEVICE_BUILD
    if (stream_sync(c->st) != 0) return fail(c, ASTREA_E_CUDA, "astrea_profile_read: stream sync failed");
    for (auto& sp : c->spans) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, sp.a, sp.b);
        ms_by_class[sp.cls] += ms;
        launches_by_class[sp.cls] += 1;
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    c->spans.clear();
#endif
    return 0;
}

void* astrea_host_alloc(int device, uint64_t bytes) {
    if (bytes == 0) return nullptr;
#ifdef ASTREA_DEVICE_BUILD
    DeviceGuard guard(device);
    void* p = nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
#else
    (void)device;
    return std::malloc((size_t)bytes);
#endif
}

void astrea_host_free(void* p) {
    if (!p) return;
#ifdef ASTREA_DEVICE_BUILD
    cudaFreeHost(p);
#else
    std::free(p);
#endif
}

int astrea_is_device_build(void) {
#ifdef ASTREA_DEVICE_BUILD
    return 1;
#else
    return 0;
#endif
}

}  // extern "C"
