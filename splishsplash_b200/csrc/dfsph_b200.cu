// C-ABI implementation + step orchestration of the B200 DFSPH hot path (see include/dfsph_b200.h).
#ifndef _GNU_SOURCE
#define _GNU_SOURCE 1   /* sched_setaffinity, CPU_SET (dfsph_b200_bind_host_numa) */
#endif
#include <sched.h>
#include <unistd.h>
#include <sys/syscall.h>
// Build: nvcc -gencode arch=compute_100a,code=sm_100a [-DDFSPH_DOUBLE] -> libdfsph_b200_f32.so / libdfsph_b200_f64.so
#include "../../include/dfsph_b200.h"
#include "common.cuh"
#include "sph_kernels.cuh"
#include "search_kernels.cuh"
#include "solver_kernels.cuh"
#include "multi_gpu.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#ifndef M_PI
#include <sched.h>
#include <unistd.h>
#include <sys/syscall.h>
#include <ctype.h>
#define M_PI 3.14159265358979323846
#endif

static thread_local std::string g_create_error;

struct dfsph_b200_ctx {
    dfsph_b200_config cfg;
    dfsph_b200_params par;
    cudaStream_t stream = nullptr;
    std::string err;
    int sticky = 0;

    unsigned n = 0, cap = 0, ntiles_cap = 0;
    unsigned Kf = 64, Kb = 64;
    double density0 = 1000.0, volume = 0.0;

    // fluid: double-buffered persistent state (search reorders from [cur] into [1-cur])
    Real4* pos[2] = {nullptr, nullptr};
    Real4* vel[2] = {nullptr, nullptr};
    Real* kappa[2] = {nullptr, nullptr};
    Real* kappa_v[2] = {nullptr, nullptr};
    unsigned* id[2] = {nullptr, nullptr};
    unsigned* state[2] = {nullptr, nullptr};
    int cur = 0;        // buffers holding vel/kappa/kappa_v/id/state
    int cur_pos = 0;    // buffer holding the current positions
    Real4 *acc = nullptr, *bgrad = nullptr;
    Real *density = nullptr, *factor = nullptr, *density_adv = nullptr;
    unsigned *nnbr = nullptr, *cnt_f = nullptr, *cnt_b = nullptr, *tab_f = nullptr, *tab_b = nullptr, *tcnt_f = nullptr, *tcnt_b = nullptr;
    unsigned *cell_key = nullptr, *cell_rank = nullptr, *cell_fine = nullptr, *sorted_idx = nullptr, *block_rank = nullptr, *block_of_rank = nullptr;
    unsigned nblocks = 0, nbx = 0;
    unsigned char* bpart_near = nullptr;   // per block part: boundary points in reach (k_mark_boundary_parts)
    bool tile_build = true, tile_attr_set = false;   // fluid-fluid table from shared-memory tiles (DFSPH_B200_TILE_BUILD=0: one-thread walk)
    unsigned *cell_count = nullptr, *cell_start = nullptr, *scan_partial = nullptr;
    unsigned long long* scan_status = nullptr;   // look-back scan: one word per chunk, tagged with scan_epoch
    unsigned* scan_ticket = nullptr;
    unsigned scan_epoch = 0;
    bool fused_reorder = true;      // single GPU: k_fix_reorder instead of k_cell_fix_order + k_reorder (DFSPH_B200_FUSED_REORDER=0)
    bool scan_single_pass = true;   // DFSPH_B200_SCAN=3: the three-kernel scan
    unsigned keys_cap = 0, scratch_cap = 0;
    bool tables_valid = false;   // neighbour table matches pos[cur_pos]
    cudaTextureObject_t acc_tex = 0, pos_tex[2] = {0, 0}, vel_tex[2] = {0, 0};

    // boundary (static Akinci2012 particles, all bodies concatenated)
    std::vector<Real4> h_bpos;
    bool have_bvol = true;
    unsigned nb = 0;
    Real4* bpos = nullptr;
    unsigned* borig = nullptr;
    unsigned* bcell_start = nullptr;
    unsigned char* bnear = nullptr;   // per cell: a boundary particle in the 3x3x3 neighbourhood
    bool boundary_dirty = true;

    double bb_min[3], bb_max[3];
    bool bb_valid = false;
    GridDesc grid;
    bool grid_valid = false;

    SphConst sph;        // solver kernel
    SphConst sph_bv;     // scalar kernel used by the boundary volume initialisation
    Real *lutW = nullptr, *lutGradW = nullptr;
    int solver_mode = KM_CUBIC_AVX, bv_mode = KM_LUT;

    Ctrl* ctrl = nullptr;
    Ctrl* h_ctrl = nullptr;      // pinned
    bool capacity_error = false; // a neighbour list overflowed (Ctrl::fatal seen on the host): sticky until set_fluid
    // solver loops as CUDA graphs: k_solve_begin -> WHILE(not done) { pass A, pass B (+ loop control), k_loop_cond }.
    // One executable graph per (solve, buffer parities); rebuilt when anything baked into the kernel arguments changes.
    struct SolveGraph { cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr; unsigned n = 0; };
    SolveGraph sgraph[2][4];   // [solve][cur * 2 + cur_pos]
    SolveGraph sgraph_slab[2][4];   // the same for slab contexts with peer memory (pushes and waits are nodes of the loop body)
    cudaStream_t capture_stream = nullptr;
    bool use_graph = true;
    double* partial = nullptr;
    unsigned pred_iter = 2, pred_iter_v = 1;
    unsigned launches = 0;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    // step_host overlap: v arrives on copy_stream while the search runs; density leaves as soon as the init sweep wrote it
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy_in = nullptr, ev_density = nullptr;
    const Real* late_vel_stage = nullptr;     // staged host velocities (id order), landed after the reorder
    Real* early_density_stage = nullptr;      // device staging + host destination of the early density download
    void* early_density_host = nullptr;
    Real* final_x_stage = nullptr;            // step_host: the finaliser leaves packed x / v (host row order) here ...
    Real* final_v_stage = nullptr;
    bool final_stage_written = false;         // ... and says so (a failed or empty step does not)
    void* stage = nullptr;       // device staging for AoS transfers
    size_t stage_bytes = 0;

    // multi-GPU slab decomposition (csrc/multi_gpu.cuh)
    bool multi = false;
    int rank = 0, world = 1;
    bool has_left = false, has_right = false;
    double slab_lo = -1e300, slab_hi = 1e300;
    int slab_axis = 0;
    NcclApi nccl;
    ncclComm_t comm = nullptr, comm2 = nullptr;      // comm: reductions, counts, migration (main stream); comm2: halo refresh (stream2)
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_a = nullptr, ev_a2 = nullptr, ev_k = nullptr, ev_k2 = nullptr;
    unsigned *exp_all = nullptr; unsigned n_exp_all = 0;
    int dbg_skip = 0;      // timing experiments only (DFSPH_B200_DEBUG_SKIP bit 1: ghost refresh inside iterations, bit 2: error all-reduce)
    // NVLink P2P ghost refresh (peer buffers mapped through CUDA IPC)
    bool p2p = false;
    unsigned* flags = nullptr;          // [3 arrays][2 sides] sequence numbers written by the neighbours; [6] = push ticket; [8 + kind] = my sequence numbers
    PushDesc* push_desc = nullptr;      // device: this step's export counts, slot offsets in the peers' arrays, owned count
    struct Peer { Real4* pos[2] = {nullptr, nullptr}; Real4* vel[2] = {nullptr, nullptr}; Real4* acc = nullptr; unsigned* flags = nullptr; bool open = false; } peer[2];
    unsigned peer_n[2] = {0, 0}, peer_ngl[2] = {0, 0};
    double* red_val = nullptr; unsigned* red_seq = nullptr;     // my all-reduce table: val [2][MAX_RANKS], seq [MAX_RANKS]
    PeerReduce pr;                                              // every rank's table (peer-mapped)
    std::vector<void*> red_opened;
    bool overlap = true;   // boundary-first overlap of the halo refresh (DFSPH_B200_NO_OVERLAP=1 selects the serial refresh)
    unsigned char* is_export = nullptr;
    Real4* ghost_stage = nullptr;
    unsigned ghost_cap = 0, ng = 0, ng_l = 0, ng_r = 0, n_exp_l = 0, n_exp_r = 0;
    unsigned *exp_l = nullptr, *exp_r = nullptr, *gcell_start = nullptr;
    unsigned* gblock_rank = nullptr;   // ghost cell table: block -> rank among the blocks in ghost reach of the slab faces (all others: one shared empty block)
    GridDesc ggrid;                    // = grid with block_rank = gblock_rank and the (small) number of entries of the ghost table
    Real4 *send_l = nullptr, *send_r = nullptr, *send_l2 = nullptr, *send_r2 = nullptr;
    MigrantAux *aux_sl = nullptr, *aux_sr = nullptr, *aux_rl = nullptr, *aux_rr = nullptr;
    ExchangeCounts* xcnt = nullptr;       // device: [0] mine, [1] from left, [2] from right
    ExchangeCounts* h_xcnt = nullptr;     // pinned mirror
    unsigned migrated_in = 0, migrated_out = 0;

    // optional per-kernel-class device timing (CUDA events on the launching stream)
    bool profiling = false;
    struct ProfRec { int cls; int seq; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    // DFSPH_B200_TRACE=1: GPU-timeline time between phase marks of a step (host gaps included), printed at destroy
    bool trace = false;
    std::vector<std::pair<const char*, cudaEvent_t>> trace_marks;
    std::vector<std::pair<const char*, double>> trace_sum;
    unsigned trace_steps = 0, trace_seen = 0;
    double prof_ms[DFSPH_B200_PROF_CLASSES] = {0};
    uint64_t prof_count[DFSPH_B200_PROF_CLASSES] = {0};
    cudaEvent_t timer_a = nullptr, timer_b = nullptr;
};

struct ProfScope {
    dfsph_b200_ctx* c; cudaEvent_t b = nullptr;
    ProfScope(dfsph_b200_ctx* c_, int cls, int seq = -1) : c(c_)
    {
        if (!c->profiling) return;
        cudaEvent_t e[2];
        for (int k = 0; k < 2; ++k) {
            if (!c->prof_pool.empty()) { e[k] = c->prof_pool.back(); c->prof_pool.pop_back(); }
            else cudaEventCreate(&e[k]);
        }
        cudaEventRecord(e[0], c->stream);
        b = e[1];
        c->prof_recs.push_back({cls, seq, e[0], e[1]});
    }
    ~ProfScope() { if (b) cudaEventRecord(b, c->stream); }
};

static void trace_mark(dfsph_b200_ctx* c, const char* name)
{
    if (!c->trace) return;
    cudaEvent_t e;
    if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); } else cudaEventCreate(&e);
    cudaEventRecord(e, c->stream);
    c->trace_marks.push_back({name, e});
}
static void trace_collect(dfsph_b200_ctx* c)   // after a stream synchronise; the time up to a mark is booked under its name
{
    if (!c->trace || c->trace_marks.empty()) return;
    const bool skip = c->trace_seen++ < 5u;   // the first steps carry one-time costs (NCCL connections, graph instantiation)
    for (size_t k = 1; !skip && k < c->trace_marks.size(); ++k) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->trace_marks[k - 1].second, c->trace_marks[k].second);
        const char* nm = c->trace_marks[k].first;
        bool found = false;
        for (auto& t : c->trace_sum) if (t.first == nm) { t.second += ms; found = true; break; }
        if (!found) c->trace_sum.push_back({nm, (double)ms});
    }
    for (auto& m : c->trace_marks) c->prof_pool.push_back(m.second);
    c->trace_marks.clear();
    if (!skip) c->trace_steps++;
}
static void trace_print(dfsph_b200_ctx* c)
{
    if (!c->trace || c->trace_steps == 0) return;
    double tot = 0; for (auto& t : c->trace_sum) tot += t.second;
    fprintf(stderr, "[dfsph_b200 trace rank %d] %u steps, %.3f ms/step between the first and the last mark\n", c->rank, c->trace_steps, tot / c->trace_steps);
    for (auto& t : c->trace_sum) fprintf(stderr, "[dfsph_b200 trace rank %d]   %-34s %8.3f ms/step\n", c->rank, t.first, t.second / c->trace_steps);
}

static void prof_collect(dfsph_b200_ctx* c)   // call after a stream synchronise
{
    for (auto& r : c->prof_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { c->prof_ms[r.cls] += ms; c->prof_count[r.cls]++; }
        c->prof_pool.push_back(r.a); c->prof_pool.push_back(r.b);
    }
    c->prof_recs.clear();
}

#define CTX_FAIL(ctx, code, ...) do { char _b[512]; snprintf(_b, sizeof(_b), __VA_ARGS__); (ctx)->err = _b; return (code); } while (0)
#define CUDA_TRY(ctx, expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { (ctx)->sticky = 1; \
    CTX_FAIL(ctx, DFSPH_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); } } while (0)
#define CHECK_CTX(ctx) do { if (!(ctx)) return DFSPH_B200_ERR_INVALID; if ((ctx)->sticky) return DFSPH_B200_ERR_CUDA; } while (0)

static inline unsigned div_up(unsigned a, unsigned b) { return (a + b - 1) / b; }
static void destroy_textures(dfsph_b200_ctx* c);
static void invalidate_graphs(dfsph_b200_ctx* c)
{
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 4; ++b) {
      for (int which = 0; which < 2; ++which) {
        dfsph_b200_ctx::SolveGraph& g = which ? c->sgraph_slab[a][b] : c->sgraph[a][b];
        if (g.exec) cudaGraphExecDestroy(g.exec);
        if (g.graph) cudaGraphDestroy(g.graph);
        g.exec = nullptr; g.graph = nullptr; g.n = 0;
      }
    }
}

template <typename T>
static int dev_alloc(dfsph_b200_ctx* c, T** p, size_t count)
{
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (count == 0) count = 1;
    CUDA_TRY(c, cudaMalloc((void**)p, count * sizeof(T)));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel constants, computed on the host in Real with the reference's own expression order
// (CubicKernel::setRadius SPHKernels.h:25-35, CubicKernel_AVX::setRadius :713-740, PrecomputedKernel::setRadius :625-644)
// ---------------------------------------------------------------------------------------------------------------
static Real host_cubic_W(Real r, Real radius, Real k)
{
    Real res = 0.0;
    const Real q = r / radius;
    if (q <= 1.0) {
        if (q <= 0.5) {
            const Real q2 = q * q;
            const Real q3 = q2 * q;
            res = k * (static_cast<Real>(6.0) * q3 - static_cast<Real>(6.0) * q2 + static_cast<Real>(1.0));
        } else {
            res = k * (static_cast<Real>(2.0) * std::pow(static_cast<Real>(1.0) - q, static_cast<Real>(3.0)));
        }
    }
    return res;
}
static Real host_cubic_gradW_x(Real rl, Real radius, Real l)   // x component of gradW((rl,0,0))
{
    Real res = 0.0;
    const Real q = rl / radius;
    if ((rl > 1.0e-9) && (q <= 1.0)) {
        Real gradq = rl / rl;
        gradq /= radius;
        if (q <= 0.5) res = l * q * ((Real)3.0 * q - static_cast<Real>(2.0)) * gradq;
        else { const Real factor = static_cast<Real>(1.0) - q; res = l * (-factor * factor) * gradq; }
    }
    return res;
}

static int setup_constants(dfsph_b200_ctx* c)
{
    const Real radius = static_cast<Real>(4.0) * static_cast<Real>(c->cfg.particle_radius);   // Simulation.cpp:283
    const Real pi = static_cast<Real>(M_PI);
    const Real h3 = radius * radius * radius;
    SphConst s;
    memset(&s, 0, sizeof(s));
    s.R = radius;
    s.R2 = radius * radius;
#if DFSPH_REAL_IS_DOUBLE
    s.invR = 1.0 / radius;
    s.k = static_cast<Real>(8.0) / (pi * h3);
    s.l = static_cast<Real>(48.0) / (pi * h3);
#else
    s.invR = 1.0f / radius;
    s.k = 8.0f / static_cast<float>(pi * h3);
    s.l = 48.0f / static_cast<float>(pi * h3);
#endif
    s.invR2 = s.invR * s.invR;
    s.g_a = s.l * s.invR2 * static_cast<Real>(3.0);
    s.g_b = -(s.l * s.invR2 * static_cast<Real>(2.0));
    s.g_ml = -s.l;
    s.W_zero = host_cubic_W(0, radius, s.k);
    s.V = static_cast<Real>(c->volume);
    s.gV_a = s.g_a * s.V; s.gV_b = s.g_b * s.V; s.gV_ml = s.g_ml * s.V;
    s.density0 = static_cast<Real>(c->density0);

    // lookup tables (PrecomputedKernel<CubicKernel, 10000>)
    std::vector<Real> W(LUT_RESOLUTION), G(LUT_RESOLUTION + 1);
    const Real stepSize = radius / (Real)(LUT_RESOLUTION - 1);
    s.lut_inv_step = static_cast<Real>(1.0) / stepSize;
    for (unsigned i = 0; i < LUT_RESOLUTION; i++) {
        const Real posX = stepSize * (Real)i;
        W[i] = host_cubic_W(posX, radius, s.k);
        if (posX > 1.0e-9) G[i] = host_cubic_gradW_x(posX, radius, s.l) / posX;
        else G[i] = 0.0;
    }
    G[LUT_RESOLUTION] = 0.0;
    // The reference evaluates 0.5*(m_W[pos] + m_W[pos+1]) per call (SPHKernels.h:657,682) with pos <= resolution-2.
    // The same expression is evaluated here once per table slot, so the device needs one load instead of two.
    std::vector<Real> Wavg(LUT_RESOLUTION, (Real)0.0), Gavg(LUT_RESOLUTION, (Real)0.0);   // slot 9999 = 0: outside the support
    for (unsigned i = 0; i + 1 < LUT_RESOLUTION; i++) {
        Wavg[i] = static_cast<Real>(0.5) * (W[i] + W[i + 1]);
        Gavg[i] = static_cast<Real>(0.5) * (G[i] + G[i + 1]);
    }
    if (dev_alloc(c, &c->lutW, LUT_RESOLUTION)) return DFSPH_B200_ERR_CUDA;
    if (dev_alloc(c, &c->lutGradW, LUT_RESOLUTION)) return DFSPH_B200_ERR_CUDA;
    CUDA_TRY(c, cudaMemcpy(c->lutW, Wavg.data(), Wavg.size() * sizeof(Real), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->lutGradW, Gavg.data(), Gavg.size() * sizeof(Real), cudaMemcpyHostToDevice));
    s.lutW = c->lutW;
    s.lutGradW = c->lutGradW;

    // constants of the other selectable kernels, computed in Real as their setRadius does (next-row f2)
    {
        const Real h9 = std::pow(radius, static_cast<Real>(9.0)), h6 = std::pow(radius, static_cast<Real>(6.0));
        s.gen_k[0] = s.k; s.gen_l[0] = s.l;
        s.gen_k[1] = static_cast<Real>(21.0) / (static_cast<Real>(2.0) * pi * h3);    // SPHKernels.h:286-287
        s.gen_l[1] = -static_cast<Real>(210.0) / (pi * h3);
        s.gen_k[2] = static_cast<Real>(315.0) / (static_cast<Real>(64.0) * pi * h9);  // SPHKernels.h:112-113
        s.gen_l[2] = -static_cast<Real>(945.0) / (static_cast<Real>(32.0) * pi * h9);
        s.gen_k[3] = static_cast<Real>(15.0) / (pi * h6);                              // SPHKernels.h:209-210
        s.gen_l[3] = -static_cast<Real>(45.0) / (pi * h6);
    }
    const int wk = c->cfg.kernel, gk = c->cfg.grad_kernel < 0 ? c->cfg.kernel : c->cfg.grad_kernel;
    s.w_kind = wk; s.g_kind = gk;
    const int scalar_mode = (wk == 4 && gk == 4) ? KM_LUT : (wk == 0 && gk == 0) ? KM_CUBIC : KM_GENERIC;
    // W(0) of the configured kernel (Simulation::setKernel, Simulation.cpp:338-393)
    Real w_zero = s.W_zero;
    if (wk == 1) w_zero = s.gen_k[1];
    else if (wk == 2) w_zero = std::pow(radius * radius, static_cast<Real>(3.0)) * s.gen_k[2];
    else if (wk == 3) w_zero = s.gen_k[3] * std::pow(radius, static_cast<Real>(3.0));
#if DFSPH_REAL_IS_DOUBLE
    c->solver_mode = scalar_mode;
    s.W_zero = w_zero;
#else
    c->solver_mode = KM_CUBIC_AVX;   // the AVX solver ignores the kernel setting
#endif
    c->bv_mode = scalar_mode;
    s.mode = c->solver_mode;
    c->sph = s;
    c->sph_bv = s;
    c->sph_bv.mode = c->bv_mode;
    c->sph_bv.W_zero = w_zero;
    return 0;
}

static SolverParams make_solver_params(const dfsph_b200_ctx* c)
{
    SolverParams sp;
    sp.gx = (Real)c->par.gravitation[0]; sp.gy = (Real)c->par.gravitation[1]; sp.gz = (Real)c->par.gravitation[2];
    sp.max_error = (Real)c->par.max_error; sp.max_error_v = (Real)c->par.max_error_v;
    sp.min_iter = c->par.min_iterations; sp.max_iter = c->par.max_iterations; sp.max_iter_v = c->par.max_iterations_v;
    sp.cfl_method = c->par.cfl_method;
    sp.cfl_factor = (Real)c->par.cfl_factor; sp.cfl_min = (Real)c->par.cfl_min_time_step_size; sp.cfl_max = (Real)c->par.cfl_max_time_step_size;
    sp.radius = (Real)c->cfg.particle_radius;
    sp.viscosity_method = c->par.viscosity_method;
    sp.viscosity = (Real)c->par.viscosity; sp.viscosity_boundary = (Real)c->par.viscosity_boundary;
    return sp;
}

// ---------------------------------------------------------------------------------------------------------------
extern "C" {

int dfsph_b200_sizeof_real(void) { return (int)sizeof(Real); }
const char* dfsph_b200_version(void) { return DFSPH_REAL_IS_DOUBLE ? "dfsph_b200 0.1 (f64, sm_100a)" : "dfsph_b200 0.1 (f32, sm_100a)"; }

void dfsph_b200_default_config(dfsph_b200_config* cfg)
{
    memset(cfg, 0, sizeof(*cfg));
    cfg->device = 0;
    cfg->kernel = DFSPH_B200_KERNEL_PRECOMPUTED_CUBIC;
    cfg->grad_kernel = -1;
    cfg->particle_radius = 0.025;
    cfg->max_fluid_neighbors = 64;
    cfg->max_boundary_neighbors = 64;
    cfg->world_size = 1;
}

void dfsph_b200_default_params(dfsph_b200_params* p)
{
    memset(p, 0, sizeof(*p));
    p->time_step_size = 0.001;
    p->gravitation[1] = -9.81;
    p->min_iterations = 2;
    p->max_iterations = 100;
    p->max_error = 0.01;
    p->max_iterations_v = 100;
    p->max_error_v = 0.1;
    p->enable_divergence_solver = 1;
    p->cfl_method = 1;
    p->cfl_factor = 0.5;
    p->cfl_min_time_step_size = 0.0001;
    p->cfl_max_time_step_size = 0.005;
    p->viscosity_method = 0;            // the reference's FluidModel default is 1 (Standard, FluidModel.cpp:98); explicit here
    p->viscosity = 0.01;                // Viscosity_Standard.cpp:21
    p->viscosity_boundary = 0.0;
}

const char* dfsph_b200_last_error(const dfsph_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int dfsph_b200_create(const dfsph_b200_config* cfg, dfsph_b200_ctx** out)
{
    if (!cfg || !out) { g_create_error = "null argument"; return DFSPH_B200_ERR_INVALID; }
    *out = nullptr;
    if (cfg->kernel < 0 || cfg->kernel > 4 || cfg->grad_kernel < -1 || cfg->grad_kernel > 4) {
        g_create_error = "unsupported kernel (0 cubic, 1 Wendland quintic C2, 2 Poly6, 3 Spiky, 4 precomputed cubic; grad_kernel -1 = same as kernel)";
        return DFSPH_B200_ERR_UNSUPPORTED;
    }
    if (!(cfg->particle_radius > 0.0)) { g_create_error = "particle_radius must be > 0"; return DFSPH_B200_ERR_INVALID; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)";
        return DFSPH_B200_ERR_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "bad device ordinal"; return DFSPH_B200_ERR_INVALID; }
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return DFSPH_B200_ERR_CUDA; }
    dfsph_b200_ctx* c = new dfsph_b200_ctx();
    c->cfg = *cfg;
    dfsph_b200_default_params(&c->par);
    c->use_graph = getenv("DFSPH_B200_NO_GRAPH") == nullptr;
    c->trace = getenv("DFSPH_B200_TRACE") != nullptr;
    if (const char* e = getenv("DFSPH_B200_TILE_BUILD")) c->tile_build = e[0] != '0';
    if (const char* e = getenv("DFSPH_B200_SCAN")) c->scan_single_pass = e[0] != '3';
    if (const char* e = getenv("DFSPH_B200_FUSED_REORDER")) c->fused_reorder = e[0] != '0';
    c->Kf = cfg->max_fluid_neighbors > 0 ? (unsigned)cfg->max_fluid_neighbors : 64u;
    c->Kb = cfg->max_boundary_neighbors > 0 ? (unsigned)cfg->max_boundary_neighbors : 64u;
    c->Kf = (c->Kf + DFSPH_PAD - 1u) & ~(DFSPH_PAD - 1u);
    c->Kb = (c->Kb + DFSPH_PAD - 1u) & ~(DFSPH_PAD - 1u);
    int rc = 0;
    do {
        if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) { rc = -1; break; }
        for (int k = 0; k < 3; ++k) if ((e = cudaEventCreate(&c->ev[k])) != cudaSuccess) { rc = -1; break; }
        if (rc) break;
        if ((e = cudaMalloc((void**)&c->ctrl, sizeof(Ctrl))) != cudaSuccess) { rc = -1; break; }
        if ((e = cudaMemset(c->ctrl, 0, sizeof(Ctrl))) != cudaSuccess) { rc = -1; break; }
        if ((e = cudaMallocHost((void**)&c->h_ctrl, sizeof(Ctrl))) != cudaSuccess) { rc = -1; break; }
        memset(c->h_ctrl, 0, sizeof(Ctrl));
    } while (0);
    if (rc) { g_create_error = cudaGetErrorString(e); dfsph_b200_destroy(c); return DFSPH_B200_ERR_CUDA; }
    *out = c;
    return DFSPH_B200_OK;
}

int dfsph_b200_destroy(dfsph_b200_ctx* c)
{
    if (!c) return DFSPH_B200_OK;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    trace_print(c);
    for (int k = 0; k < 2; ++k) {
        cudaFree(c->pos[k]); cudaFree(c->vel[k]); cudaFree(c->kappa[k]); cudaFree(c->kappa_v[k]); cudaFree(c->id[k]); cudaFree(c->state[k]);
    }
    destroy_textures(c);
    invalidate_graphs(c);
    if (c->capture_stream) cudaStreamDestroy(c->capture_stream);
    cudaFree(c->acc); cudaFree(c->bgrad); cudaFree(c->density); cudaFree(c->factor); cudaFree(c->density_adv);
    cudaFree(c->nnbr); cudaFree(c->cnt_f); cudaFree(c->cnt_b); cudaFree(c->tab_f); cudaFree(c->tab_b); cudaFree(c->tcnt_f); cudaFree(c->tcnt_b);
    cudaFree(c->cell_key); cudaFree(c->cell_rank); cudaFree(c->cell_fine); cudaFree(c->block_rank); cudaFree(c->block_of_rank); cudaFree(c->bpart_near); cudaFree(c->sorted_idx); cudaFree(c->cell_count); cudaFree(c->cell_start);
    cudaFree(c->scan_partial); cudaFree(c->scan_status); cudaFree(c->scan_ticket); cudaFree(c->bpos); cudaFree(c->borig); cudaFree(c->bcell_start); cudaFree(c->bnear);
    cudaFree(c->lutW); cudaFree(c->lutGradW); cudaFree(c->ctrl); cudaFree(c->partial); cudaFree(c->stage);
    cudaFree(c->exp_l); cudaFree(c->exp_r); cudaFree(c->gcell_start); cudaFree(c->gblock_rank); cudaFree(c->send_l); cudaFree(c->send_r);
    cudaFree(c->send_l2); cudaFree(c->send_r2); cudaFree(c->aux_sl); cudaFree(c->aux_sr); cudaFree(c->aux_rl); cudaFree(c->aux_rr);
    for (int s = 0; s < 2; ++s) if (c->peer[s].open) {
        for (int b = 0; b < 2; ++b) { cudaIpcCloseMemHandle(c->peer[s].pos[b]); cudaIpcCloseMemHandle(c->peer[s].vel[b]); }
        cudaIpcCloseMemHandle(c->peer[s].acc); cudaIpcCloseMemHandle(c->peer[s].flags);
    }
    for (void* p : c->red_opened) cudaIpcCloseMemHandle(p);
    cudaFree(c->flags); cudaFree(c->push_desc); cudaFree(c->red_val); cudaFree(c->red_seq);
    cudaFree(c->xcnt); cudaFree(c->exp_all); cudaFree(c->is_export); cudaFree(c->ghost_stage);
    if (c->comm2 && c->nccl.CommDestroy) c->nccl.CommDestroy(c->comm2);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->ev_copy_in) cudaEventDestroy(c->ev_copy_in);
    if (c->ev_density) cudaEventDestroy(c->ev_density);
    if (c->ev_a) { cudaEventDestroy(c->ev_a); cudaEventDestroy(c->ev_a2); cudaEventDestroy(c->ev_k); cudaEventDestroy(c->ev_k2); }
    if (c->h_xcnt) cudaFreeHost(c->h_xcnt);
    if (c->comm && c->nccl.CommDestroy) c->nccl.CommDestroy(c->comm);
    if (c->h_ctrl) cudaFreeHost(c->h_ctrl);
    for (int k = 0; k < 3; ++k) if (c->ev[k]) cudaEventDestroy(c->ev[k]);
    for (auto& r : c->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : c->prof_pool) cudaEventDestroy(e);
    if (c->timer_a) cudaEventDestroy(c->timer_a);
    if (c->timer_b) cudaEventDestroy(c->timer_b);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return DFSPH_B200_OK;
}

uint64_t dfsph_b200_num_particles(const dfsph_b200_ctx* c) { return c ? c->n : 0; }
uint64_t dfsph_b200_num_boundary_particles(const dfsph_b200_ctx* c) { return c ? c->nb : 0; }

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
static void bbox_extend(dfsph_b200_ctx* c, const Real* x, uint64_t n)
{
    if (!c->bb_valid) { for (int k = 0; k < 3; ++k) { c->bb_min[k] = 1e300; c->bb_max[k] = -1e300; } c->bb_valid = true; }
    for (uint64_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            const double v = (double)x[3 * i + k];
            if (v < c->bb_min[k]) c->bb_min[k] = v;
            if (v > c->bb_max[k]) c->bb_max[k] = v;
        }
}

static int ensure_stage(dfsph_b200_ctx* c, size_t bytes)
{
    if (bytes <= c->stage_bytes) return 0;
    if (c->stage) cudaFree(c->stage);
    c->stage = nullptr; c->stage_bytes = 0;
    CUDA_TRY(c, cudaMalloc(&c->stage, bytes));
    c->stage_bytes = bytes;
    return 0;
}

static int ensure_scratch(dfsph_b200_ctx* c, unsigned count)
{
    if (count <= c->scratch_cap) return 0;
    if (dev_alloc(c, &c->cell_key, count)) return DFSPH_B200_ERR_CUDA;
    if (dev_alloc(c, &c->cell_rank, count)) return DFSPH_B200_ERR_CUDA;
    if (dev_alloc(c, &c->cell_fine, count)) return DFSPH_B200_ERR_CUDA;
    if (dev_alloc(c, &c->sorted_idx, count)) return DFSPH_B200_ERR_CUDA;
    c->scratch_cap = count;
    return 0;
}

static int setup_grid(dfsph_b200_ctx* c)
{
    double lo[3], hi[3];
    bool given = false;
    for (int k = 0; k < 3; ++k) if (c->cfg.domain_max[k] > c->cfg.domain_min[k]) given = true;
    const double S = (double)c->sph.R * (1.0 + 1.0e-5);
    if (given) {
        for (int k = 0; k < 3; ++k) { lo[k] = c->cfg.domain_min[k]; hi[k] = c->cfg.domain_max[k]; }
        if (c->multi) {
            // every rank only needs cells for its own slab plus the ghost / boundary halo (2 cells): the cell tables stay
            // proportional to the slab instead of the global domain (cell keys are rank-local)
            const int a = c->slab_axis;
            lo[a] = std::max(lo[a], c->slab_lo - 2.0 * S);
            hi[a] = std::min(hi[a], c->slab_hi + 2.0 * S);
        }
    }
    else {
        if (!c->bb_valid) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "no particles: cannot derive the cell grid");
        for (int k = 0; k < 3; ++k) { lo[k] = c->bb_min[k] - S; hi[k] = c->bb_max[k] + S; }
    }
    GridDesc g;
    g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2];
    g.inv_cell = 1.0 / S;
    int nc[3];
    for (int k = 0; k < 3; ++k) {
        const double cells = std::ceil((hi[k] - lo[k]) / S);
        if (cells > 8192.0) CTX_FAIL(c, DFSPH_B200_ERR_CAPACITY, "cell grid too large on axis %d (%g cells)", k, cells);
        nc[k] = std::max(1, (int)cells);
    }
    g.nx = nc[0]; g.ny = nc[1]; g.nz = nc[2];
    const unsigned nbx = div_up(g.nx, DFSPH_BX), nby = div_up(g.ny, DFSPH_BY), nbz = div_up(g.nz, DFSPH_BZ);
    g.nby = (int)nby; g.nbz = (int)nbz;
    const unsigned long long keys = (unsigned long long)nbx * nby * nbz * DFSPH_ENTRIES_PER_BLOCK;
    if (keys > 3000000000ull) CTX_FAIL(c, DFSPH_B200_ERR_CAPACITY, "cell table too large (%llu entries)", keys);
    g.num_keys = (unsigned)keys;
    {
        // z-order over the blocks: rank of every block's Morton code (z most significant, like the in-block code)
        const unsigned nblocks = nbx * nby * nbz;
        auto spread10 = [](unsigned long long v) {
            v &= 0x3ffull;
            v = (v | (v << 16)) & 0x30000ffull;
            v = (v | (v << 8)) & 0x300f00full;
            v = (v | (v << 4)) & 0x30c30c3ull;
            v = (v | (v << 2)) & 0x9249249ull;
            return v;
        };
        std::vector<std::pair<unsigned long long, unsigned>> codes(nblocks);
        for (unsigned bx = 0; bx < nbx; ++bx) for (unsigned by = 0; by < nby; ++by) for (unsigned bz = 0; bz < nbz; ++bz) {
            const unsigned lin = (bx * nby + by) * nbz + bz;
            codes[lin] = { spread10(bx) | (spread10(by) << 1) | (spread10(bz) << 2), lin };
        }
        std::sort(codes.begin(), codes.end());
        std::vector<unsigned> rank(nblocks);
        for (unsigned r = 0; r < nblocks; ++r) rank[codes[r].second] = r;
        if (dev_alloc(c, &c->block_rank, nblocks)) return DFSPH_B200_ERR_CUDA;
        CUDA_TRY(c, cudaMemcpy(c->block_rank, rank.data(), (size_t)nblocks * sizeof(unsigned), cudaMemcpyHostToDevice));
        g.block_rank = c->block_rank;
        std::vector<unsigned> inv(nblocks);
        for (unsigned r = 0; r < nblocks; ++r) inv[r] = codes[r].second;
        if (dev_alloc(c, &c->block_of_rank, nblocks)) return DFSPH_B200_ERR_CUDA;
        CUDA_TRY(c, cudaMemcpy(c->block_of_rank, inv.data(), (size_t)nblocks * sizeof(unsigned), cudaMemcpyHostToDevice));
        g.block_of_rank = c->block_of_rank;
        c->nblocks = nblocks; c->nbx = nbx;
        if (c->multi) {
            // Ghost particles sit within one cell of a slab face, outside the slab, i.e. in the 2 halo cells at either end of the
            // slab axis (3 are taken).  Only the blocks that hold such cells get rows in the ghost cell table, in curve order; every
            // other block maps to ONE shared block behind them that stays empty, so that the walk finds "no candidates" there
            // without a test.  The per-step ghost sort (memset, scan, entry fix-up) then runs over a table of a few percent of the
            // full size (at 10 M particles per slab: 0.3 ms -> 0.03 ms per step).
            const int a = c->slab_axis;
            const int na = a == 0 ? g.nx : (a == 1 ? g.ny : g.nz);
            const int bl = a == 0 ? DFSPH_BX_LOG2 : (a == 1 ? DFSPH_BY_LOG2 : DFSPH_BZ_LOG2);
            const int halo_cells = 4;   // 2 halo cells + the rounding of the cell count + one cell of margin
            std::vector<unsigned> grank(nblocks);
            unsigned nh = 0;
            for (unsigned r = 0; r < nblocks; ++r) {
                const unsigned lin = codes[r].second;
                const unsigned b3[3] = { lin / (nby * nbz), (lin / nbz) % nby, lin % nbz };
                const int c0 = (int)(b3[a] << bl), c1 = std::min(na, (int)((b3[a] + 1u) << bl));   // cells [c0, c1) of the block along the slab axis
                const bool halo = c0 < halo_cells || c1 > na - halo_cells;
                grank[lin] = halo ? nh++ : 0xffffffffu;
            }
            for (unsigned lin = 0; lin < nblocks; ++lin) if (grank[lin] == 0xffffffffu) grank[lin] = nh;
            if (dev_alloc(c, &c->gblock_rank, nblocks)) return DFSPH_B200_ERR_CUDA;
            CUDA_TRY(c, cudaMemcpy(c->gblock_rank, grank.data(), (size_t)nblocks * sizeof(unsigned), cudaMemcpyHostToDevice));
            c->ggrid = g;
            c->ggrid.block_rank = c->gblock_rank;
            c->ggrid.num_keys = (nh + 1u) * DFSPH_ENTRIES_PER_BLOCK;
        }
    }
    c->grid = g;
    const unsigned nfine = c->multi ? std::max(g.num_keys, c->ggrid.num_keys) : g.num_keys;   // table entries (+ the dump cell); the ghost
                                                                                              // table has one block more than the halo blocks
    if (nfine + 1 > c->keys_cap) {
        if (dev_alloc(c, &c->cell_count, (size_t)nfine + 8)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->cell_start, (size_t)nfine + 8)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->bcell_start, (size_t)nfine + 8)) return DFSPH_B200_ERR_CUDA;
        if (c->multi && dev_alloc(c, &c->gcell_start, (size_t)nfine + 8)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->scan_partial, div_up(nfine + 1, SCAN_CHUNK) + 8)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->scan_status, div_up(nfine + 1, SCAN1_CHUNK) + 8)) return DFSPH_B200_ERR_CUDA;
        CUDA_TRY(c, cudaMemsetAsync(c->scan_status, 0, ((size_t)div_up(nfine + 1, SCAN1_CHUNK) + 8) * sizeof(unsigned long long), c->stream));
        if (!c->scan_ticket && dev_alloc(c, &c->scan_ticket, 1)) return DFSPH_B200_ERR_CUDA;
        c->scan_epoch = 0;
        c->keys_cap = nfine + 1;
    }
    c->grid_valid = true;
    c->boundary_dirty = true;
    c->tables_valid = false;
    return 0;
}

// counting sort of `n` points at `pos` into the cell table `cell_start_out`; leaves the permutation in sorted_idx
// slab: particles outside the context's slab are filed under the dump key num_keys (multi-GPU migration)
// fix = false: the caller orders the entries itself (k_fix_reorder)
static int cell_sort(dfsph_b200_ctx* c, const Real4* pos, unsigned n, unsigned* cell_start_out, bool slab = false, bool fix = true, const GridDesc* grid = nullptr)
{
    const GridDesc& g = grid ? *grid : c->grid;
    cudaStream_t st = c->stream;
    const unsigned nk = g.num_keys + 1u;   // + dump cell
    CUDA_TRY(c, cudaMemsetAsync(c->cell_count, 0, (size_t)nk * sizeof(unsigned), st));
    if (n > 0) {
        if (slab) k_cell_hash_slab<<<div_up(n, DFSPH_BLOCK), DFSPH_BLOCK, 0, st>>>(pos, n, g, c->slab_lo, c->slab_hi, c->slab_axis, c->has_left, c->has_right,
                                                                                   c->cell_count, c->cell_key, c->cell_rank, c->cell_fine);
        else k_cell_hash<<<div_up(n, DFSPH_BLOCK), DFSPH_BLOCK, 0, st>>>(pos, n, g, c->cell_count, c->cell_key, c->cell_rank, c->cell_fine);
        c->launches++;
    }
    const unsigned nparts = div_up(nk, SCAN_CHUNK);
    if (c->scan_single_pass) {
        if (++c->scan_epoch >= (1u << 30)) {   // the tag has 30 bits: start over with a clean status array
            CUDA_TRY(c, cudaMemsetAsync(c->scan_status, 0, ((size_t)div_up(nk, SCAN1_CHUNK) + 8) * sizeof(unsigned long long), st));
            c->scan_epoch = 1;
        }
        CUDA_TRY(c, cudaMemsetAsync(c->scan_ticket, 0, sizeof(unsigned), st));
        k_scan_lookback<<<div_up(nk, SCAN1_CHUNK), SCAN_BLOCK, 0, st>>>(c->cell_count, nk, cell_start_out, c->scan_status, c->scan_ticket, c->scan_epoch);
        c->launches += 1;
    } else {
    k_scan_partials<<<nparts, SCAN_BLOCK, 0, st>>>(c->cell_count, nk, c->scan_partial);
    k_scan_spine<<<1, SCAN_BLOCK, 0, st>>>(c->scan_partial, nparts);
    k_scan_apply<<<nparts, SCAN_BLOCK, 0, st>>>(c->cell_count, nk, c->scan_partial, nparts, cell_start_out);
    c->launches += 3;
    }
    if (n > 0) {
        k_cell_scatter<<<div_up(n, DFSPH_BLOCK), DFSPH_BLOCK, 0, st>>>(c->cell_key, c->cell_rank, n, cell_start_out, c->sorted_idx);
        if (fix) k_cell_fix_order<<<div_up(nk, DFSPH_BLOCK), DFSPH_BLOCK, 0, st>>>(cell_start_out, nk, c->cell_fine, c->sorted_idx);
        c->launches += fix ? 2 : 1;
    }
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

__global__ void k_set_w(Real4* p, const Real* w, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i].w = w[i];
}
__global__ void k_iota(unsigned* p, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

static int finalize_boundary(dfsph_b200_ctx* c)
{
    if (!c->boundary_dirty) return 0;
    const unsigned nb = (unsigned)c->h_bpos.size();
    c->nb = nb;
    const unsigned need = nb + 1u;   // +1: sentinel boundary particle
    Real4* tmp = nullptr;
    unsigned* tmp_orig = nullptr;
    if (dev_alloc(c, &c->bpos, need)) return DFSPH_B200_ERR_CUDA;
    if (dev_alloc(c, &c->borig, need)) return DFSPH_B200_ERR_CUDA;
    if (nb > 0) {
        { int rs = ensure_scratch(c, nb); if (rs) return rs; }   // sorting scratch is sized for max(n, nb)
        CUDA_TRY(c, cudaMalloc((void**)&tmp, (size_t)nb * sizeof(Real4)));
        CUDA_TRY(c, cudaMalloc((void**)&tmp_orig, (size_t)nb * sizeof(unsigned)));
        CUDA_TRY(c, cudaMemcpyAsync(tmp, c->h_bpos.data(), (size_t)nb * sizeof(Real4), cudaMemcpyHostToDevice, c->stream));
        k_iota<<<div_up(nb, 256), 256, 0, c->stream>>>(tmp_orig, nb);
    }
    int rc = cell_sort(c, tmp, nb, c->bcell_start);
    if (rc == 0) {
        const size_t ncell = (size_t)c->grid.nx * c->grid.ny * c->grid.nz;
        if (dev_alloc(c, &c->bnear, ncell)) rc = DFSPH_B200_ERR_CUDA;
        else if (cudaMemsetAsync(c->bnear, 0, ncell, c->stream) != cudaSuccess) rc = DFSPH_B200_ERR_CUDA;
    }
    if (rc == 0) {
        const unsigned nparts = c->nblocks * TB_PARTS;
        if (dev_alloc(c, &c->bpart_near, nparts)) rc = DFSPH_B200_ERR_CUDA;
        else k_mark_boundary_parts<<<div_up(nparts, 4), 128, 0, c->stream>>>(c->grid, c->nbx, nparts, c->bcell_start, c->bpart_near);
    }
    if (rc == 0 && nb > 0) {
        k_mark_boundary_cells<<<div_up(nb, DFSPH_BLOCK), DFSPH_BLOCK, 0, c->stream>>>(nb, c->grid, tmp, c->bnear);
        k_reorder_boundary<<<div_up(nb, DFSPH_BLOCK), DFSPH_BLOCK, 0, c->stream>>>(nb, c->sorted_idx, tmp, tmp_orig, c->bpos, c->borig);
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { c->sticky = 1; c->err = cudaGetErrorString(e); rc = DFSPH_B200_ERR_CUDA; }
    }
    if (tmp) cudaFree(tmp);
    if (tmp_orig) cudaFree(tmp_orig);
    if (rc) return rc;
    c->boundary_dirty = false;
    c->tables_valid = false;
    return 0;
}

static cudaTextureObject_t make_linear_texture(void* ptr, size_t texels)
{
    cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = ptr;
#if DFSPH_REAL_IS_DOUBLE
    rd.res.linear.desc = cudaCreateChannelDesc<int4>();
#else
    rd.res.linear.desc = cudaCreateChannelDesc<float4>();
#endif
    rd.res.linear.sizeInBytes = texels * 16;
    cudaTextureDesc td; memset(&td, 0, sizeof(td)); td.readMode = cudaReadModeElementType;
    cudaTextureObject_t t = 0;
    if (cudaCreateTextureObject(&t, &rd, &td, nullptr) != cudaSuccess) { t = 0; cudaGetLastError(); }
    return t;
}
static void destroy_textures(dfsph_b200_ctx* c)
{
    if (c->acc_tex) { cudaDestroyTextureObject(c->acc_tex); c->acc_tex = 0; }
    for (int k = 0; k < 2; ++k) {
        if (c->pos_tex[k]) { cudaDestroyTextureObject(c->pos_tex[k]); c->pos_tex[k] = 0; }
        if (c->vel_tex[k]) { cudaDestroyTextureObject(c->vel_tex[k]); c->vel_tex[k] = 0; }
    }
}

static int alloc_fluid(dfsph_b200_ctx* c, unsigned cap)
{
    invalidate_graphs(c);   // device pointers are baked into the graphs' kernel arguments
    if (c->multi) {
        // room for one support radius of ghosts per face plus migrants; generous: an eighth of the slab, >= 256 Ki
        c->ghost_cap = std::max(cap / 8u, 262144u);
        const unsigned gc = c->ghost_cap;
        if (dev_alloc(c, &c->exp_l, gc)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->exp_r, gc)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->send_l, gc)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->send_r, gc)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->send_l2, gc)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->send_r2, gc)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->aux_sl, gc)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->aux_sr, gc)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->aux_rl, gc)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->aux_rr, gc)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->xcnt, 3)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->exp_all, 2 * gc)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->ghost_stage, gc)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->is_export, cap + gc)) return DFSPH_B200_ERR_CUDA;
        if (!c->h_xcnt) CUDA_TRY(c, cudaMallocHost((void**)&c->h_xcnt, 3 * sizeof(ExchangeCounts)));
        cap += gc;   // owned capacity includes room for arrivals
    }
    for (int k = 0; k < 2; ++k) {
        if (dev_alloc(c, &c->pos[k], cap + c->ghost_cap + 1)) return DFSPH_B200_ERR_CUDA;   // ghosts, then the sentinel particle
        if (dev_alloc(c, &c->vel[k], cap + c->ghost_cap + 1)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->kappa[k], cap)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->kappa_v[k], cap)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->id[k], cap)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->state[k], cap)) return DFSPH_B200_ERR_CUDA;
    }
    if (dev_alloc(c, &c->acc, cap + c->ghost_cap + 1)) return DFSPH_B200_ERR_CUDA;
    {
        // the sweeps gather neighbour records through the texture unit (16 B texels: one per particle in the float build,
        // two in the double build); without a texture the kernels fall back to plain loads
        destroy_textures(c);
        const size_t texels = ((size_t)cap + c->ghost_cap + 1) * (sizeof(Real4) / 16);
        if (texels < (1ull << 27) && !getenv("DFSPH_B200_NO_TEX")) {   // linear textures address at most 2^27 texels
            c->acc_tex = make_linear_texture(c->acc, texels);
            for (int k = 0; k < 2; ++k) { c->pos_tex[k] = make_linear_texture(c->pos[k], texels); c->vel_tex[k] = make_linear_texture(c->vel[k], texels); }
        }
    }
    if (dev_alloc(c, &c->bgrad, cap)) return DFSPH_B200_ERR_CUDA;
    if (dev_alloc(c, &c->density, cap)) return DFSPH_B200_ERR_CUDA;
    if (dev_alloc(c, &c->factor, cap)) return DFSPH_B200_ERR_CUDA;
    if (dev_alloc(c, &c->density_adv, cap)) return DFSPH_B200_ERR_CUDA;
    if (dev_alloc(c, &c->nnbr, cap)) return DFSPH_B200_ERR_CUDA;
    if (dev_alloc(c, &c->cnt_f, cap)) return DFSPH_B200_ERR_CUDA;
    if (dev_alloc(c, &c->cnt_b, cap)) return DFSPH_B200_ERR_CUDA;
    const unsigned ntiles = div_up(cap, DFSPH_TILE);
    c->ntiles_cap = ntiles;
    if (dev_alloc(c, &c->tab_f, (size_t)ntiles * c->Kf * DFSPH_TILE)) return DFSPH_B200_ERR_CUDA;
    if (dev_alloc(c, &c->tab_b, (size_t)ntiles * c->Kb * DFSPH_TILE)) return DFSPH_B200_ERR_CUDA;
    if (dev_alloc(c, &c->tcnt_f, ntiles)) return DFSPH_B200_ERR_CUDA;
    if (dev_alloc(c, &c->tcnt_b, ntiles)) return DFSPH_B200_ERR_CUDA;
    { int rs = ensure_scratch(c, cap); if (rs) return rs; }
    if (dev_alloc(c, &c->partial, 2 * (div_up(cap, DFSPH_BLOCK) + 1))) return DFSPH_B200_ERR_CUDA;
    c->cap = cap;
    return 0;
}

// AoS Real[3] <-> Real4 conversions through the device staging buffer
__global__ void k_unpack3(const Real* __restrict__ src, Real4* __restrict__ dst, unsigned n, const unsigned* __restrict__ id, int keep_w)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned s = id ? id[i] : i;
    Real4 v = make_real4(src[3 * (size_t)s], src[3 * (size_t)s + 1], src[3 * (size_t)s + 2], (Real)0.0);
    if (keep_w) v.w = dst[i].w;
    st_real4(dst + i, v);
}
__global__ void k_pack3(const Real4* __restrict__ src, Real* __restrict__ dst, unsigned n, const unsigned* __restrict__ id)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned d = id ? id[i] : i;
    const Real4 v = ld_gather(src + i);
    dst[3 * (size_t)d] = v.x; dst[3 * (size_t)d + 1] = v.y; dst[3 * (size_t)d + 2] = v.z;
}
template <typename T>
__global__ void k_pack1(const T* __restrict__ src, T* __restrict__ dst, unsigned n, const unsigned* __restrict__ id)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[id ? id[i] : i] = src[i];
}
template <typename T>
__global__ void k_unpack1(const T* __restrict__ src, T* __restrict__ dst, unsigned n, const unsigned* __restrict__ id)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[i] = src[id ? id[i] : i];
}
__global__ void k_pack_w(const Real4* __restrict__ src, Real* __restrict__ dst, unsigned n, const unsigned* __restrict__ id)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[id ? id[i] : i] = src[i].w;
}

extern "C" {

int dfsph_b200_set_fluid(dfsph_b200_ctx* c, uint64_t n64, const void* x_, const void* v_, const uint32_t* id, const uint32_t* state,
                         double density0, double volume)
{
    CHECK_CTX(c);
    cudaSetDevice(c->cfg.device);
    if (n64 > 0 && !x_) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "x is NULL");
    if (n64 > 0xfffffff0ull) CTX_FAIL(c, DFSPH_B200_ERR_CAPACITY, "too many particles for 32-bit indices");
    if (!(density0 > 0.0) || !(volume > 0.0)) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "density0 and volume must be > 0");
    const unsigned n = (unsigned)n64;
    unsigned cap = (unsigned)std::max<uint64_t>(std::max<uint64_t>(c->cfg.max_fluid_particles, n), 1);
    // Peer-memory contexts: the neighbour ranks hold CUDA-IPC mappings of this rank's particle arrays; growing them would
    // free the memory those mappings point at.  A restart on the same communicator has to fit the exported capacity.
    if (c->p2p && cap > c->cap)
        CTX_FAIL(c, DFSPH_B200_ERR_CAPACITY, "set_fluid with %llu particles exceeds the capacity (%u) whose buffers the peer ranks have mapped: create the "
                 "context with max_fluid_particles large enough, or build a new context and exchange the peer handles again", (unsigned long long)n64, c->cap);
    if (!c->multi && id) {
        // single GPU: by-id transfers scatter to host row id[i], so the ids must be a permutation of 0..n-1
        std::vector<unsigned char> seen(n, 0);
        for (unsigned i = 0; i < n; ++i) {
            if (id[i] >= n || seen[id[i]]) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "ids must be a permutation of 0..n-1 (id[%u] = %u)", i, id[i]);
            seen[id[i]] = 1;
        }
    }
    if (cap > c->cap) { int rc = alloc_fluid(c, cap); if (rc) return rc; }
    c->n = n;
    c->density0 = density0;
    c->volume = volume;
    c->capacity_error = false;
    invalidate_graphs(c);
    { int rc = setup_constants(c); if (rc) return rc; }
    const Real* x = (const Real*)x_;
    const Real* v = (const Real*)v_;
    std::vector<Real4> hp(n), hv(n);
    std::vector<unsigned> hid(n), hst(n);
    for (unsigned i = 0; i < n; ++i) {
        hp[i] = make_real4(x[3 * i], x[3 * i + 1], x[3 * i + 2], (Real)0.0);
        hv[i] = v ? make_real4(v[3 * i], v[3 * i + 1], v[3 * i + 2], (Real)0.0) : make_real4(0, 0, 0, 0);
        hid[i] = id ? id[i] : i;
        hst[i] = state ? state[i] : 0u;
    }
    c->cur = 0; c->cur_pos = 0;
    if (n > 0) {
        CUDA_TRY(c, cudaMemcpy(c->pos[0], hp.data(), (size_t)n * sizeof(Real4), cudaMemcpyHostToDevice));
        CUDA_TRY(c, cudaMemcpy(c->vel[0], hv.data(), (size_t)n * sizeof(Real4), cudaMemcpyHostToDevice));
        CUDA_TRY(c, cudaMemcpy(c->id[0], hid.data(), (size_t)n * sizeof(unsigned), cudaMemcpyHostToDevice));
        CUDA_TRY(c, cudaMemcpy(c->state[0], hst.data(), (size_t)n * sizeof(unsigned), cudaMemcpyHostToDevice));
        CUDA_TRY(c, cudaMemset(c->kappa[0], 0, (size_t)n * sizeof(Real)));
        CUDA_TRY(c, cudaMemset(c->kappa_v[0], 0, (size_t)n * sizeof(Real)));
        CUDA_TRY(c, cudaMemset(c->density, 0, (size_t)n * sizeof(Real)));
        CUDA_TRY(c, cudaMemset(c->factor, 0, (size_t)n * sizeof(Real)));
        CUDA_TRY(c, cudaMemset(c->density_adv, 0, (size_t)n * sizeof(Real)));
        CUDA_TRY(c, cudaMemset(c->acc, 0, (size_t)n * sizeof(Real4)));
        CUDA_TRY(c, cudaMemset(c->nnbr, 0, (size_t)n * sizeof(unsigned)));
        bbox_extend(c, x, n);
    }
    // time step size / time live on the device
    Ctrl hc;
    memset(&hc, 0, sizeof(hc));
    hc.h = (Real)c->par.time_step_size;
    hc.h_step = hc.h;
    hc.multi = c->multi ? 1u : 0u;
    hc.n_global = n;
    if (c->multi) {
        // a repeated set_fluid (restart of a run on the same communicator) must not rewind the sequence number of the
        // peer-memory all-reduce: the peers' tables still hold the values published so far
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        CUDA_TRY(c, cudaMemcpy(&hc.red_seq, &c->ctrl->red_seq, sizeof(unsigned), cudaMemcpyDeviceToHost));
    }
    CUDA_TRY(c, cudaMemcpy(c->ctrl, &hc, sizeof(hc), cudaMemcpyHostToDevice));
    c->grid_valid = false;
    c->tables_valid = false;
    c->pred_iter = 2; c->pred_iter_v = 1;
    return DFSPH_B200_OK;
}

int dfsph_b200_add_boundary(dfsph_b200_ctx* c, uint64_t n, const void* x_, const void* V_, int is_dynamic)
{
    CHECK_CTX(c);
    if (is_dynamic) CTX_FAIL(c, DFSPH_B200_ERR_UNSUPPORTED, "dynamic / animated rigid bodies are outside the hot-path scope (static Akinci2012 boundaries only)");
    if (n > 0 && !x_) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "x is NULL");
    if (c->h_bpos.size() + n > 0xfffffff0ull) CTX_FAIL(c, DFSPH_B200_ERR_CAPACITY, "too many boundary particles");
    const Real* x = (const Real*)x_;
    const Real* V = (const Real*)V_;
    if (!V) c->have_bvol = false;
    for (uint64_t i = 0; i < n; ++i) c->h_bpos.push_back(make_real4(x[3 * i], x[3 * i + 1], x[3 * i + 2], V ? V[i] : (Real)0.0));
    bbox_extend(c, x, n);
    c->boundary_dirty = true;
    c->grid_valid = false;
    c->tables_valid = false;
    return DFSPH_B200_OK;
}

static int prepare(dfsph_b200_ctx* c)
{
    cudaSetDevice(c->cfg.device);
    if (c->cap == 0) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "set_fluid has not been called");
    if (!c->grid_valid) { int rc = setup_grid(c); if (rc) return rc; }
    if (c->boundary_dirty) { int rc = finalize_boundary(c); if (rc) return rc; }
    return 0;
}

int dfsph_b200_compute_boundary_volume(dfsph_b200_ctx* c)
{
    CHECK_CTX(c);
    if (c->volume <= 0.0) {   // constants need the support radius only; allow calling before set_fluid
        c->volume = 1.0;
        int rc0 = setup_constants(c); c->volume = 0.0; if (rc0) return rc0;
    }
    if (c->cap == 0) { int rc = alloc_fluid(c, (unsigned)std::max<uint64_t>(c->cfg.max_fluid_particles, 1)); if (rc) return rc; }
    int rc = prepare(c);
    if (rc) return rc;
    const unsigned nb = c->nb;
    if (nb == 0) { c->have_bvol = true; return DFSPH_B200_OK; }
    Real* vol = nullptr;
    CUDA_TRY(c, cudaMalloc((void**)&vol, (size_t)nb * sizeof(Real)));
    const Real W0 = c->sph_bv.W_zero;   // sim->W_zero() (BoundaryModel_Akinci2012.cpp:61)
    if (c->bv_mode == KM_LUT) k_boundary_volume<KM_LUT><<<div_up(nb, DFSPH_BLOCK), DFSPH_BLOCK, 0, c->stream>>>(nb, c->grid, c->sph_bv, W0, c->bpos, c->bcell_start, vol);
    else if (c->bv_mode == KM_GENERIC) k_boundary_volume<KM_GENERIC><<<div_up(nb, DFSPH_BLOCK), DFSPH_BLOCK, 0, c->stream>>>(nb, c->grid, c->sph_bv, W0, c->bpos, c->bcell_start, vol);
    else k_boundary_volume<KM_CUBIC><<<div_up(nb, DFSPH_BLOCK), DFSPH_BLOCK, 0, c->stream>>>(nb, c->grid, c->sph_bv, W0, c->bpos, c->bcell_start, vol);
    k_set_w<<<div_up(nb, 256), 256, 0, c->stream>>>(c->bpos, vol, nb);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    // keep the host copy in sync (the boundary may be re-sorted if the grid changes)
    std::vector<Real> hv(nb);
    std::vector<unsigned> ho(nb);
    if (e == cudaSuccess) e = cudaMemcpy(hv.data(), vol, (size_t)nb * sizeof(Real), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(ho.data(), c->borig, (size_t)nb * sizeof(unsigned), cudaMemcpyDeviceToHost);
    cudaFree(vol);
    if (e != cudaSuccess) { c->sticky = 1; CTX_FAIL(c, DFSPH_B200_ERR_CUDA, "boundary volume: %s", cudaGetErrorString(e)); }
    for (unsigned i = 0; i < nb; ++i) c->h_bpos[ho[i]].w = hv[i];
    c->have_bvol = true;
    return DFSPH_B200_OK;
}

int dfsph_b200_set_params(dfsph_b200_ctx* c, const dfsph_b200_params* p)
{
    CHECK_CTX(c);
    if (!p) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "null params");
    dfsph_b200_params q = *p;
    // GenParam min-value clamps of the reference (TimeStepDFSPH.cpp:86-109)
    if (q.max_iterations < 1) q.max_iterations = 1;
    if (q.max_iterations_v < 1) q.max_iterations_v = 1;
    if (q.max_error < 1e-6) q.max_error = 1e-6;
    if (q.max_error_v < 1e-6) q.max_error_v = 1e-6;
    if (!(q.time_step_size > 0.0)) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "time_step_size must be > 0");
    if (q.cfl_method < 0 || q.cfl_method > 2) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "cfl_method must be 0, 1 or 2");
    if (q.viscosity_method != 0 && q.viscosity_method != 1) CTX_FAIL(c, DFSPH_B200_ERR_UNSUPPORTED, "viscosity_method: only 0 (none) and 1 (Standard viscosity) run on the B200 path");
    if (q.viscosity < 0.0) q.viscosity = 0.0;
    if (q.viscosity_boundary < 0.0) q.viscosity_boundary = 0.0;
    c->par = q;
    invalidate_graphs(c);   // solver parameters are kernel arguments
    if (c->ctrl) {
        // TimeManager::setTimeStepSize always takes effect: compare with the step size the DEVICE holds (the CFL kernel
        // adapts it), not with the last requested value
        cudaSetDevice(c->cfg.device);
        const Real h = (Real)q.time_step_size;
        Real hd = (Real)0.0;
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        CUDA_TRY(c, cudaMemcpy(&hd, &c->ctrl->h, sizeof(Real), cudaMemcpyDeviceToHost));
        if (hd != h) CUDA_TRY(c, cudaMemcpy(&c->ctrl->h, &h, sizeof(Real), cudaMemcpyHostToDevice));
    }
    return DFSPH_B200_OK;
}

int dfsph_b200_get_params(const dfsph_b200_ctx* c, dfsph_b200_params* p)
{
    if (!c || !p) return DFSPH_B200_ERR_INVALID;
    *p = c->par;
    return DFSPH_B200_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
static FluidArrays fluid_arrays(dfsph_b200_ctx* c)
{
    FluidArrays f;
    f.pos = c->pos[c->cur_pos];
    f.vel = c->vel[c->cur];
    f.acc = c->acc; f.bgrad = c->bgrad;
    f.density = c->density; f.factor = c->factor; f.density_adv = c->density_adv;
    f.kappa = c->kappa[c->cur]; f.kappa_v = c->kappa_v[c->cur];
    f.state = c->state[c->cur]; f.nnbr = c->nnbr;
    f.tab_f = c->tab_f; f.cnt_f = c->cnt_f; f.tab_b = c->tab_b; f.cnt_b = c->cnt_b; f.tcnt_f = c->tcnt_f; f.tcnt_b = c->tcnt_b;
    f.Kf = c->Kf; f.Kb = c->Kb; f.n = c->n; f.n_dev = nullptr;
    f.acc_tex = c->acc_tex; f.pos_tex = c->pos_tex[c->cur_pos]; f.vel_tex = c->vel_tex[c->cur];
    return f;
}

#define NCCL_TRY(ctx, expr) do { ncclResult_t _r = (expr); if (_r != ncclSuccess) { (ctx)->sticky = 1; \
    CTX_FAIL(ctx, DFSPH_B200_ERR_COMM, "%s failed: %s", #expr, (ctx)->nccl.GetErrorString(_r)); } } while (0)

// exchange the per-rank counters with both slab neighbours and bring all three to the host (one synchronisation)
static int exchange_counts(dfsph_b200_ctx* c)
{
    cudaStream_t st = c->stream;
    const size_t b = sizeof(ExchangeCounts);
    CUDA_TRY(c, cudaMemsetAsync(c->xcnt + 1, 0, 2 * b, st));
    NCCL_TRY(c, c->nccl.GroupStart());
    if (c->has_left) { NCCL_TRY(c, c->nccl.Send(c->xcnt, b, ncclChar, c->rank - 1, c->comm, st)); NCCL_TRY(c, c->nccl.Recv(c->xcnt + 1, b, ncclChar, c->rank - 1, c->comm, st)); }
    if (c->has_right) { NCCL_TRY(c, c->nccl.Send(c->xcnt, b, ncclChar, c->rank + 1, c->comm, st)); NCCL_TRY(c, c->nccl.Recv(c->xcnt + 2, b, ncclChar, c->rank + 1, c->comm, st)); }
    NCCL_TRY(c, c->nccl.GroupEnd());
    CUDA_TRY(c, cudaMemcpyAsync(c->h_xcnt, c->xcnt, 3 * b, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    return 0;
}

// refresh one Real4 field of the ghost particles from the owning ranks (ghost slots [n, n+ng) of `arr`)
enum { GH_POS = 0, GH_VEL = 1, GH_ACC = 2 };

// P2P version: push the export values into the neighbours' ghost slots, then wait for theirs.  Counts, slot offsets and
// sequence numbers are read from device memory (PushDesc, flags + 8 + kind), so the same launches work inside a graph.
#define DFSPH_PUSH_BLOCKS 96u
struct PushArgs {
    const Real4* src; Real4 *dst_l, *dst_r; unsigned *flag_l, *flag_r, *seq_word; const unsigned *wait_l, *wait_r;
};
static PushArgs push_args(const dfsph_b200_ctx* c, int kind)
{
    PushArgs a;
    a.src = kind == GH_POS ? c->pos[c->cur_pos] : (kind == GH_VEL ? c->vel[c->cur] : c->acc);
    auto peer_arr = [&](int side) -> Real4* {
        const dfsph_b200_ctx::Peer& p = c->peer[side];
        return kind == GH_POS ? p.pos[c->cur_pos] : (kind == GH_VEL ? p.vel[c->cur] : p.acc);
    };
    // my exports to the left neighbour are ITS right ghosts (behind its left ghosts); to the right neighbour its left ghosts:
    // the offsets inside the peers' arrays are in the PushDesc
    a.dst_l = c->has_left ? peer_arr(0) : nullptr;
    a.dst_r = c->has_right ? peer_arr(1) : nullptr;
    a.flag_l = c->has_left ? c->peer[0].flags + kind * 2 + 1 : nullptr;    // I am the left neighbour's RIGHT side
    a.flag_r = c->has_right ? c->peer[1].flags + kind * 2 + 0 : nullptr;   // and the right neighbour's LEFT side
    a.seq_word = c->flags + 8 + kind;
    a.wait_l = c->has_left ? c->flags + kind * 2 + 0 : nullptr;
    a.wait_r = c->has_right ? c->flags + kind * 2 + 1 : nullptr;
    return a;
}
static void launch_push(dfsph_b200_ctx* c, const PushArgs& a, cudaStream_t st, unsigned blocks)
{
    k_push_exports<<<blocks, DFSPH_BLOCK, 0, st>>>(a.src, c->exp_l, a.dst_l, a.flag_l, c->exp_r, a.dst_r, a.flag_r, c->push_desc, a.seq_word, c->flags + 6);
}
static int p2p_refresh(dfsph_b200_ctx* c, int kind, bool wait = true)
{
    cudaStream_t st = c->stream;
    const PushArgs a = push_args(c, kind);
    const unsigned tot = c->n_exp_l + c->n_exp_r;
    launch_push(c, a, st, std::min(std::max(div_up(tot, DFSPH_BLOCK), 1u), DFSPH_PUSH_BLOCKS));
    c->launches++;
    if (wait) {
        k_wait_flags<<<1, 1, 0, st>>>(a.wait_l, a.wait_r, a.seq_word);
        c->launches++;
    }
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

static int exchange_ghosts(dfsph_b200_ctx* c, Real4* arr)
{
    if (!c->multi) return 0;
    if (c->p2p) return p2p_refresh(c, arr == c->acc ? GH_ACC : (arr == c->vel[c->cur] ? GH_VEL : GH_POS));
    cudaStream_t st = c->stream;
    const unsigned tot = c->n_exp_l + c->n_exp_r;
    if (tot > 0) { k_pack_exports<<<div_up(tot, DFSPH_BLOCK), DFSPH_BLOCK, 0, st>>>(arr, c->exp_l, c->n_exp_l, c->exp_r, c->n_exp_r, c->send_l, c->send_r); c->launches++; }
    const size_t e = sizeof(Real4);
    NCCL_TRY(c, c->nccl.GroupStart());
    if (c->has_left) {
        NCCL_TRY(c, c->nccl.Send(c->send_l, (size_t)c->n_exp_l * e, ncclChar, c->rank - 1, c->comm, st));
        NCCL_TRY(c, c->nccl.Recv(arr + c->n, (size_t)c->ng_l * e, ncclChar, c->rank - 1, c->comm, st));
    }
    if (c->has_right) {
        NCCL_TRY(c, c->nccl.Send(c->send_r, (size_t)c->n_exp_r * e, ncclChar, c->rank + 1, c->comm, st));
        NCCL_TRY(c, c->nccl.Recv(arr + c->n + c->ng_l, (size_t)c->ng_r * e, ncclChar, c->rank + 1, c->comm, st));
    }
    NCCL_TRY(c, c->nccl.GroupEnd());
    return 0;
}

// The same refresh on the communication stream (comm2 / stream2): starts once `after` has completed on the main
// stream, packs the export list of `arr`, receives into `dst` (ghost order: left face first), records `done`.
static int exchange_ghosts_async(dfsph_b200_ctx* c, const Real4* arr, Real4* dst, cudaEvent_t after, cudaEvent_t done)
{
    cudaStream_t s2 = c->stream2;
    CUDA_TRY(c, cudaStreamWaitEvent(s2, after, 0));
    const unsigned tot = c->n_exp_l + c->n_exp_r;
    if (tot > 0) { k_pack_exports<<<div_up(tot, DFSPH_BLOCK), DFSPH_BLOCK, 0, s2>>>(arr, c->exp_l, c->n_exp_l, c->exp_r, c->n_exp_r, c->send_l, c->send_r); c->launches++; }
    const size_t e = sizeof(Real4);
    NCCL_TRY(c, c->nccl.GroupStart());
    if (c->has_left) {
        NCCL_TRY(c, c->nccl.Send(c->send_l, (size_t)c->n_exp_l * e, ncclChar, c->rank - 1, c->comm2, s2));
        NCCL_TRY(c, c->nccl.Recv(dst, (size_t)c->ng_l * e, ncclChar, c->rank - 1, c->comm2, s2));
    }
    if (c->has_right) {
        NCCL_TRY(c, c->nccl.Send(c->send_r, (size_t)c->n_exp_r * e, ncclChar, c->rank + 1, c->comm2, s2));
        NCCL_TRY(c, c->nccl.Recv(dst + c->ng_l, (size_t)c->ng_r * e, ncclChar, c->rank + 1, c->comm2, s2));
    }
    NCCL_TRY(c, c->nccl.GroupEnd());
    CUDA_TRY(c, cudaEventRecord(done, s2));
    return 0;
}

// Simulation::performNeighborhoodSearch: sort + reorder + neighbour table
static int run_search(dfsph_b200_ctx* c)
{
    cudaStream_t st = c->stream;
    unsigned n = c->n, n_sorted = c->n;
    int rc;
    if (c->multi) {
        // ---- particle migration: owned particles that left the slab go to the neighbour rank --------------------------
        k_zero_counts<<<1, 1, 0, st>>>(c->xcnt);
        if (n > 0) k_pack_leavers<<<div_up(n, DFSPH_BLOCK), DFSPH_BLOCK, 0, st>>>(n, c->pos[c->cur_pos], c->vel[c->cur], c->kappa[c->cur], c->kappa_v[c->cur],
            c->id[c->cur], c->state[c->cur], c->slab_lo, c->slab_hi, c->slab_axis, c->has_left, c->has_right, c->ghost_cap,
            c->send_l, c->send_l2, c->aux_sl, c->send_r, c->send_r2, c->aux_sr, c->xcnt);
        c->launches += 2;
        rc = exchange_counts(c);
        if (rc) return rc;
        const unsigned out_l = c->h_xcnt[0].leave_l, out_r = c->h_xcnt[0].leave_r;
        const unsigned in_l = c->has_left ? c->h_xcnt[1].leave_r : 0u, in_r = c->has_right ? c->h_xcnt[2].leave_l : 0u;
        if (out_l > c->ghost_cap || out_r > c->ghost_cap || in_l + in_r > c->ghost_cap || (unsigned long long)n + in_l + in_r > c->cap)
            CTX_FAIL(c, DFSPH_B200_ERR_CAPACITY, "migration buffers too small (out %u/%u in %u/%u)", out_l, out_r, in_l, in_r);
        // Each exchange is decided PER PAIR of neighbours, from numbers both sides know (my out_l is the left rank's in_r):
        // a rank-wide test would let an interior rank post messages to a neighbour that has nothing to exchange and
        // therefore skips the group -- an unmatched send/recv, i.e. a hang (seen at 4 GPUs, never at 2).
        const bool xl = c->has_left && (out_l + in_l > 0), xr = c->has_right && (out_r + in_r > 0);
        if (xl || xr) {
            const size_t e = sizeof(Real4), a = sizeof(MigrantAux);
            Real4* pdst = c->pos[c->cur_pos] + n;
            Real4* vdst = c->vel[c->cur] + n;
            NCCL_TRY(c, c->nccl.GroupStart());
            if (xl) {
                NCCL_TRY(c, c->nccl.Send(c->send_l, out_l * e, ncclChar, c->rank - 1, c->comm, st));
                NCCL_TRY(c, c->nccl.Send(c->send_l2, out_l * e, ncclChar, c->rank - 1, c->comm, st));
                NCCL_TRY(c, c->nccl.Send(c->aux_sl, out_l * a, ncclChar, c->rank - 1, c->comm, st));
                NCCL_TRY(c, c->nccl.Recv(pdst, in_l * e, ncclChar, c->rank - 1, c->comm, st));
                NCCL_TRY(c, c->nccl.Recv(vdst, in_l * e, ncclChar, c->rank - 1, c->comm, st));
                NCCL_TRY(c, c->nccl.Recv(c->aux_rl, in_l * a, ncclChar, c->rank - 1, c->comm, st));
            }
            if (xr) {
                NCCL_TRY(c, c->nccl.Send(c->send_r, out_r * e, ncclChar, c->rank + 1, c->comm, st));
                NCCL_TRY(c, c->nccl.Send(c->send_r2, out_r * e, ncclChar, c->rank + 1, c->comm, st));
                NCCL_TRY(c, c->nccl.Send(c->aux_sr, out_r * a, ncclChar, c->rank + 1, c->comm, st));
                NCCL_TRY(c, c->nccl.Recv(pdst + in_l, in_r * e, ncclChar, c->rank + 1, c->comm, st));
                NCCL_TRY(c, c->nccl.Recv(vdst + in_l, in_r * e, ncclChar, c->rank + 1, c->comm, st));
                NCCL_TRY(c, c->nccl.Recv(c->aux_rr, in_r * a, ncclChar, c->rank + 1, c->comm, st));
            }
            NCCL_TRY(c, c->nccl.GroupEnd());
            if (in_l) k_unpack_arrivals<<<div_up(in_l, DFSPH_BLOCK), DFSPH_BLOCK, 0, st>>>(in_l, n, c->aux_rl, c->kappa[c->cur], c->kappa_v[c->cur], c->id[c->cur], c->state[c->cur]);
            if (in_r) k_unpack_arrivals<<<div_up(in_r, DFSPH_BLOCK), DFSPH_BLOCK, 0, st>>>(in_r, n + in_l, c->aux_rr, c->kappa[c->cur], c->kappa_v[c->cur], c->id[c->cur], c->state[c->cur]);
        }
        trace_mark(c, "search: leavers + count exchange + migration");
        c->migrated_in += in_l + in_r;
        c->migrated_out += out_l + out_r;
        const unsigned n1 = n + in_l + in_r;
        { ProfScope ps(c, DFSPH_B200_PROF_SORT); rc = cell_sort(c, c->pos[c->cur_pos], n1, c->cell_start, true, !c->fused_reorder); }
        if (rc) return rc;
        trace_mark(c, "search: cell sort");
        n_sorted = n1;
        n = n1 - out_l - out_r;   // leavers sit in the dump cell behind the kept particles
        c->n = n;
    } else {
        ProfScope ps(c, DFSPH_B200_PROF_SORT);
        rc = cell_sort(c, c->pos[c->cur_pos], n, c->cell_start, false, !c->fused_reorder);
        if (rc) return rc;
        n_sorted = n;
    }
    c->ng = c->ng_l = c->ng_r = 0;
    if (n > 0 || c->multi) {   // multi-GPU: every rank flips its buffers every step (peers address them by parity)
        const int src = c->cur, dst = 1 - c->cur;
        const int psrc = c->cur_pos, pdst = 1 - c->cur_pos;
        ProfScope ps(c, DFSPH_B200_PROF_SORT);
        if (n > 0 && c->fused_reorder)
            k_fix_reorder<<<div_up(n_sorted, DFSPH_BLOCK), DFSPH_BLOCK, 0, st>>>(n_sorted, n, c->sorted_idx, c->cell_key, c->cell_fine, c->cell_start, c->cell_rank,
                c->pos[psrc], c->vel[src], c->kappa[src], c->kappa_v[src], c->id[src], c->state[src],
                c->pos[pdst], c->vel[dst], c->kappa[dst], c->kappa_v[dst], c->id[dst], c->state[dst], c->acc);
        else if (n > 0) k_reorder<<<div_up(n, DFSPH_BLOCK), DFSPH_BLOCK, 0, st>>>(n, c->sorted_idx, c->pos[psrc], c->vel[src], c->kappa[src], c->kappa_v[src],
            c->id[src], c->state[src], c->pos[pdst], c->vel[dst], c->kappa[dst], c->kappa_v[dst], c->id[dst], c->state[dst], c->acc);
        c->cur = dst; c->cur_pos = pdst;
        c->launches++;
    }
    trace_mark(c, "search: reorder");
    if (c->multi) {
        // ---- ghost layer: one cell width of the neighbouring slabs, appended behind the owned particles ---------------
        k_zero_counts<<<1, 1, 0, st>>>(c->xcnt);
        const double width = 1.0 / c->grid.inv_cell;
        if (n > 0) k_select_exports<<<div_up(n, DFSPH_BLOCK), DFSPH_BLOCK, 0, st>>>(n, c->pos[c->cur_pos], c->slab_lo, c->slab_hi, width, c->slab_axis,
            c->has_left, c->has_right, c->ghost_cap, c->exp_l, c->exp_r, c->exp_all, c->is_export, c->xcnt);
        c->launches += 2;
        rc = exchange_counts(c);
        if (rc) return rc;
        c->n_exp_l = c->h_xcnt[0].exp_l; c->n_exp_r = c->h_xcnt[0].exp_r; c->n_exp_all = c->h_xcnt[0].exp_all;
        c->ng_l = c->has_left ? c->h_xcnt[1].exp_r : 0u;
        c->ng_r = c->has_right ? c->h_xcnt[2].exp_l : 0u;
        c->ng = c->ng_l + c->ng_r;
        if (c->n_exp_l > c->ghost_cap || c->n_exp_r > c->ghost_cap || c->ng > c->ghost_cap)
            CTX_FAIL(c, DFSPH_B200_ERR_CAPACITY, "ghost buffers too small (export %u/%u, ghosts %u, capacity %u)", c->n_exp_l, c->n_exp_r, c->ng, c->ghost_cap);
        if (c->p2p) {
            // the peers' owned counts and left-ghost counts locate my slots inside their ghost regions
            ExchangeCounts mine = c->h_xcnt[0];
            mine.pad0 = n; mine.pad1 = c->ng_l;
            CUDA_TRY(c, cudaMemcpyAsync(c->xcnt, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
            rc = exchange_counts(c);
            if (rc) return rc;
            c->peer_n[0] = c->h_xcnt[1].pad0; c->peer_ngl[0] = c->h_xcnt[1].pad1;
            c->peer_n[1] = c->h_xcnt[2].pad0; c->peer_ngl[1] = c->h_xcnt[2].pad1;
            k_set_push_desc<<<1, 1, 0, st>>>(c->push_desc, c->n_exp_l, c->n_exp_r, (unsigned long long)c->peer_n[0] + c->peer_ngl[0],
                                             (unsigned long long)c->peer_n[1], n);
            c->launches++;
        }
        trace_mark(c, "search: export lists + count exchanges");
        rc = exchange_ghosts(c, c->pos[c->cur_pos]); if (rc) return rc;
        rc = exchange_ghosts(c, c->vel[c->cur]); if (rc) return rc;
        trace_mark(c, "search: ghost x, v");
        k_write_sentinel<<<1, 1, 0, st>>>(c->pos[c->cur_pos], c->vel[c->cur], c->acc, n + c->ng);
        if (c->ng > 0) CUDA_TRY(c, cudaMemsetAsync(c->acc + n, 0, (size_t)c->ng * sizeof(Real4), st));
        c->launches++;
        // ghost cell table (ghosts stay in arrival order; the permutation is left in sorted_idx)
        rc = cell_sort(c, c->pos[c->cur_pos] + n, c->ng, c->gcell_start, false, true, &c->ggrid);
        if (rc) return rc;
        // global particle count = divisor of the average density error
        const unsigned long long nn = n;
        CUDA_TRY(c, cudaMemcpyAsync(&c->ctrl->n_global, &nn, sizeof(nn), cudaMemcpyHostToDevice, st));
        NCCL_TRY(c, c->nccl.AllReduce(&c->ctrl->n_global, &c->ctrl->n_global, 1, ncclUint64, ncclSum, c->comm, st));
        trace_mark(c, "search: ghost cell sort + global count");
    } else if (n == 0) {
        k_write_sentinel<<<1, 1, 0, st>>>(c->pos[c->cur_pos], c->vel[c->cur], c->acc, 0);
    }
    if (n > 0) {
        ProfScope ps(c, DFSPH_B200_PROF_BUILD);
        if (c->tile_build) {
            if (!c->tile_attr_set) {
                CUDA_TRY(c, cudaFuncSetAttribute(k_build_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TB_SMEM_BYTES));
                c->tile_attr_set = true;
            }
            const double gh_lo = c->has_left ? c->slab_lo + 1.001 / c->grid.inv_cell : -1e300, gh_hi = c->has_right ? c->slab_hi - 1.001 / c->grid.inv_cell : 1e300;
            k_build_tiles<<<c->nblocks * TB_PARTS, DFSPH_TB_THREADS, TB_SMEM_BYTES, st>>>(c->grid, c->nbx, c->sph.R2, n, 1,
                c->pos[c->cur_pos], c->cell_start, c->tab_f, c->Kf, c->cnt_f, c->tcnt_f,
                c->bpos, c->bcell_start, c->nb, c->tab_b, c->Kb, c->cnt_b, c->tcnt_b, n + c->ng, c->bpart_near, c->ctrl,
                c->multi && c->ng > 0 ? c->slab_axis : -1, gh_lo, gh_hi);
            k_build_neighbors<false><<<div_up(n, DFSPH_BLOCK), DFSPH_BLOCK, 0, st>>>(n, c->grid, c->sph.R2, c->pos[c->cur_pos], c->cell_start,
                c->bpos, c->bcell_start, c->nb, c->bnear, c->tab_f, c->Kf, c->tab_b, c->Kb, c->cnt_f, c->cnt_b, c->tcnt_f, c->tcnt_b, c->ctrl,
                c->ng, c->gcell_start, c->sorted_idx, c->slab_axis,
                c->has_left ? c->slab_lo + 1.001 / c->grid.inv_cell : -1e300, c->has_right ? c->slab_hi - 1.001 / c->grid.inv_cell : 1e300, c->gblock_rank);
            c->launches++;
        } else
        k_build_neighbors<true><<<div_up(n, DFSPH_BLOCK), DFSPH_BLOCK, 0, st>>>(n, c->grid, c->sph.R2, c->pos[c->cur_pos], c->cell_start,
            c->bpos, c->bcell_start, c->nb, c->bnear, c->tab_f, c->Kf, c->tab_b, c->Kb, c->cnt_f, c->cnt_b, c->tcnt_f, c->tcnt_b, c->ctrl,
            c->ng, c->gcell_start, c->sorted_idx, c->slab_axis,
            c->has_left ? c->slab_lo + 1.001 / c->grid.inv_cell : -1e300, c->has_right ? c->slab_hi - 1.001 / c->grid.inv_cell : 1e300, c->gblock_rank);
        k_check_capacity<<<1, 1, 0, st>>>(c->ctrl, c->Kf, c->Kb);
        c->launches += 2;
    }
    trace_mark(c, "search: neighbour table");
    CUDA_TRY(c, cudaGetLastError());
    c->tables_valid = true;
    return 0;
}

template <int MODE>
static int run_solver(dfsph_b200_ctx* c)
{
    const unsigned n = c->n;
    cudaStream_t st = c->stream;
    const unsigned grid = std::max(div_up(n, DFSPH_BLOCK), 1u);
    const SolverParams sp = make_solver_params(c);
    const bool div = c->par.enable_divergence_solver != 0;
    FluidArrays f = fluid_arrays(c);

    const bool multi = c->multi;
    Real4* gpos = c->pos[c->cur_pos];
    { ProfScope ps(c, DFSPH_B200_PROF_INIT);
    if (div) k_init_sweep<MODE, true><<<grid, DFSPH_BLOCK, 0, st>>>(f, c->sph, c->bpos, c->ctrl);
    else k_init_sweep<MODE, false><<<grid, DFSPH_BLOCK, 0, st>>>(f, c->sph, c->bpos, c->ctrl); }
    c->launches++;
    if (c->early_density_host && n > 0) {   // step_host: the density of this step is final here; send it home behind the solver
        CUDA_TRY(c, cudaEventRecord(c->ev_density, st));
        CUDA_TRY(c, cudaStreamWaitEvent(c->copy_stream, c->ev_density, 0));
        k_pack1<Real><<<div_up(n, 256), 256, 0, c->copy_stream>>>(c->density, c->early_density_stage, n, multi ? nullptr : c->id[c->cur]);
        CUDA_TRY(c, cudaMemcpyAsync(c->early_density_host, c->early_density_stage, (size_t)n * sizeof(Real), cudaMemcpyDeviceToHost, c->copy_stream));
        c->launches++;
        c->early_density_host = nullptr;
    }

    PeerReduce no_red; memset(&no_red, 0, sizeof(no_red));
    auto launch_accel = [&](const unsigned* list, unsigned list_n, const unsigned char* skip, int seq, GhostWait gw = GhostWait{nullptr, nullptr, 0u}) {
        const unsigned g = list ? std::max(div_up(list_n, DFSPH_BLOCK), 1u) : grid;
        const bool keep = c->profiling;
        if (list) c->profiling = false;   // the small export-list launch is not part of the per-kernel statistics
        ProfScope ps(c, DFSPH_B200_PROF_ACCEL, seq);
        c->profiling = keep;
        k_accel<MODE><<<g, DFSPH_BLOCK, 0, st>>>(f, c->sph, c->ctrl, list, list_n, skip, gw);
        c->launches++;
    };
    auto launch_jacobi = [&](int solve, const unsigned* list, unsigned list_n, const unsigned char* skip, unsigned base, int finalize, int seq,
                             GhostWait gw = GhostWait{nullptr, nullptr, 0u}, const PeerReduce* prp = nullptr) {
        const PeerReduce& pr = prp ? *prp : no_red;
        const unsigned g = std::max(div_up(list ? list_n : n, DFSPH_JACOBI_BLOCK), 1u);
        const bool keep = c->profiling;
        if (list) c->profiling = false;
        ProfScope ps(c, solve == SOLVE_DIV ? DFSPH_B200_PROF_JACOBI_DIV : DFSPH_B200_PROF_JACOBI_PRESS, seq);
        c->profiling = keep;
        if (solve == SOLVE_DIV) k_jacobi<MODE, SOLVE_DIV><<<g, DFSPH_JACOBI_BLOCK, 0, st>>>(f, c->sph, sp, c->ctrl, c->partial, list, list_n, skip, base, finalize, gw, pr);
        else k_jacobi<MODE, SOLVE_PRESS><<<g, DFSPH_JACOBI_BLOCK, 0, st>>>(f, c->sph, sp, c->ctrl, c->partial, list, list_n, skip, base, finalize, gw, pr);
        c->launches++;
    };
    // multi-GPU: ghost kappa that arrived on the communication stream is moved into pos[n ..] on the main stream
    auto land_ghost_kappa = [&]() -> int {
        CUDA_TRY(c, cudaStreamWaitEvent(st, c->ev_k2, 0));
        if (c->ng > 0) { k_copy_real4<<<div_up(c->ng, DFSPH_BLOCK), DFSPH_BLOCK, 0, st>>>(c->ghost_stage, gpos + c->n, c->ng); c->launches++; }
        return 0;
    };

    // Single GPU, not profiling: the loop runs as a CUDA graph -- k_solve_begin, then a WHILE node whose body is pass A,
    // pass B (which takes the reference's loop decision on the device) and k_loop_cond.  No host synchronisation.
    auto graph_loop = [&](int solve) -> int {
        dfsph_b200_ctx::SolveGraph& sg = c->sgraph[solve][c->cur * 2 + c->cur_pos];
        if (!sg.exec || sg.n != n) {
            if (sg.exec) { cudaGraphExecDestroy(sg.exec); sg.exec = nullptr; }
            if (sg.graph) { cudaGraphDestroy(sg.graph); sg.graph = nullptr; }
            if (!c->capture_stream) CUDA_TRY(c, cudaStreamCreateWithFlags(&c->capture_stream, cudaStreamNonBlocking));
            CUDA_TRY(c, cudaGraphCreate(&sg.graph, 0));
            cudaGraphConditionalHandle handle;
            CUDA_TRY(c, cudaGraphConditionalHandleCreate(&handle, sg.graph, 1, cudaGraphCondAssignDefault));
            cudaGraphNode_t n_begin, n_while;
            {
                cudaKernelNodeParams kp; memset(&kp, 0, sizeof(kp));
                void* args[2] = {(void*)&c->ctrl, (void*)&solve};
                kp.func = (void*)k_solve_begin; kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.kernelParams = args;
                CUDA_TRY(c, cudaGraphAddKernelNode(&n_begin, sg.graph, nullptr, 0, &kp));
            }
            cudaGraphNodeParams cp = {};
            cp.type = cudaGraphNodeTypeConditional;
            cp.conditional.handle = handle; cp.conditional.type = cudaGraphCondTypeWhile; cp.conditional.size = 1;
            CUDA_TRY(c, cudaGraphAddNode(&n_while, sg.graph, &n_begin, 1, &cp));
            cudaGraph_t body = cp.conditional.phGraph_out[0];
            cudaStream_t cs = c->capture_stream;
            CUDA_TRY(c, cudaStreamBeginCaptureToGraph(cs, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
            const GhostWait no_wait{nullptr, nullptr, 0u};
            k_accel<MODE><<<grid, DFSPH_BLOCK, 0, cs>>>(f, c->sph, c->ctrl, nullptr, 0u, nullptr, no_wait);
            const unsigned gj = std::max(div_up(n, DFSPH_JACOBI_BLOCK), 1u);
            if (solve == SOLVE_DIV) k_jacobi<MODE, SOLVE_DIV><<<gj, DFSPH_JACOBI_BLOCK, 0, cs>>>(f, c->sph, sp, c->ctrl, c->partial, nullptr, 0u, nullptr, 0u, 1, no_wait, no_red);
            else k_jacobi<MODE, SOLVE_PRESS><<<gj, DFSPH_JACOBI_BLOCK, 0, cs>>>(f, c->sph, sp, c->ctrl, c->partial, nullptr, 0u, nullptr, 0u, 1, no_wait, no_red);
            k_loop_cond<<<1, 1, 0, cs>>>(handle, c->ctrl);
            cudaGraph_t captured = nullptr;
            CUDA_TRY(c, cudaStreamEndCapture(cs, &captured));
            CUDA_TRY(c, cudaGraphInstantiate(&sg.exec, sg.graph, 0));
            sg.n = n;
        }
        CUDA_TRY(c, cudaGraphLaunch(sg.exec, st));
        c->launches += 1;   // + 3 kernels per iteration, added from the iteration counters when the step's statistics are read
        return 0;
    };

    // Slabs with peer memory, not profiling: the same WHILE graph with the ghost refreshes as nodes of the loop body --
    //   wait kappa | pass A | push a | wait a | pass B (+ fused all-reduce + loop decision) | push kappa | k_loop_cond.
    // Everything that changes from step to step (owned count, export counts, slot offsets in the peers' arrays, sequence
    // numbers) is read from device memory, so one graph per buffer parity serves the whole run; the grids are sized for the
    // context's capacity.  All ranks take the identical loop decision (the all-reduce), so they leave the loop together.
    auto graph_loop_slab = [&](int solve) -> int {
        dfsph_b200_ctx::SolveGraph& sg = c->sgraph_slab[solve][c->cur * 2 + c->cur_pos];
        { int rg = exchange_ghosts(c, gpos); if (rg) return rg; }   // warm-start kappa of the ghosts (push + wait)
        if (!sg.exec || sg.n != c->cap) {
            if (sg.exec) { cudaGraphExecDestroy(sg.exec); sg.exec = nullptr; }
            if (sg.graph) { cudaGraphDestroy(sg.graph); sg.graph = nullptr; }
            if (!c->capture_stream) CUDA_TRY(c, cudaStreamCreateWithFlags(&c->capture_stream, cudaStreamNonBlocking));
            CUDA_TRY(c, cudaGraphCreate(&sg.graph, 0));
            cudaGraphConditionalHandle handle;
            CUDA_TRY(c, cudaGraphConditionalHandleCreate(&handle, sg.graph, 1, cudaGraphCondAssignDefault));
            cudaGraphNode_t n_begin, n_while;
            {
                cudaKernelNodeParams kp; memset(&kp, 0, sizeof(kp));
                void* args[2] = {(void*)&c->ctrl, (void*)&solve};
                kp.func = (void*)k_solve_begin; kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.kernelParams = args;
                CUDA_TRY(c, cudaGraphAddKernelNode(&n_begin, sg.graph, nullptr, 0, &kp));
            }
            cudaGraphNodeParams cp = {};
            cp.type = cudaGraphNodeTypeConditional;
            cp.conditional.handle = handle; cp.conditional.type = cudaGraphCondTypeWhile; cp.conditional.size = 1;
            CUDA_TRY(c, cudaGraphAddNode(&n_while, sg.graph, &n_begin, 1, &cp));
            cudaGraph_t body = cp.conditional.phGraph_out[0];
            cudaStream_t cs = c->capture_stream;
            FluidArrays fg = f;
            fg.n_dev = &c->push_desc->n_local;
            const unsigned cap = c->cap;
            const PushArgs pa = push_args(c, GH_ACC), pk = push_args(c, GH_POS);
            const GhostWait no_wait{nullptr, nullptr, 0u};
            CUDA_TRY(c, cudaStreamBeginCaptureToGraph(cs, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
            k_wait_flags<<<1, 1, 0, cs>>>(pk.wait_l, pk.wait_r, pk.seq_word);
            k_accel<MODE><<<std::max(div_up(cap, DFSPH_BLOCK), 1u), DFSPH_BLOCK, 0, cs>>>(fg, c->sph, c->ctrl, nullptr, 0u, nullptr, no_wait);
            launch_push(c, pa, cs, DFSPH_PUSH_BLOCKS);
            k_wait_flags<<<1, 1, 0, cs>>>(pa.wait_l, pa.wait_r, pa.seq_word);
            const unsigned gj = std::max(div_up(cap, DFSPH_JACOBI_BLOCK), 1u);
            if (solve == SOLVE_DIV) k_jacobi<MODE, SOLVE_DIV><<<gj, DFSPH_JACOBI_BLOCK, 0, cs>>>(fg, c->sph, sp, c->ctrl, c->partial, nullptr, 0u, nullptr, 0u, 1, no_wait, c->pr);
            else k_jacobi<MODE, SOLVE_PRESS><<<gj, DFSPH_JACOBI_BLOCK, 0, cs>>>(fg, c->sph, sp, c->ctrl, c->partial, nullptr, 0u, nullptr, 0u, 1, no_wait, c->pr);
            launch_push(c, pk, cs, DFSPH_PUSH_BLOCKS);
            k_loop_cond<<<1, 1, 0, cs>>>(handle, c->ctrl);
            cudaGraph_t captured = nullptr;
            CUDA_TRY(c, cudaStreamEndCapture(cs, &captured));
            CUDA_TRY(c, cudaGraphInstantiate(&sg.exec, sg.graph, 0));
            sg.n = cap;
        }
        CUDA_TRY(c, cudaGraphLaunch(sg.exec, st));
        {   // the last kappa push of the loop: the finaliser reads the ghosts' kappa
            const PushArgs w = push_args(c, GH_POS);
            k_wait_flags<<<1, 1, 0, st>>>(w.wait_l, w.wait_r, w.seq_word);
        }
        c->launches += 2;   // + 7 kernels per iteration, added from the iteration counters when the step's statistics are read
        return 0;
    };

    auto solve_loop = [&](int solve, unsigned max_it, unsigned& pred) -> int {
        if (c->use_graph && !multi && !c->profiling) return graph_loop(solve);
        if (c->use_graph && multi && c->p2p && !c->profiling) return graph_loop_slab(solve);   // (the same decision on every rank)
        k_solve_begin<<<1, 1, 0, st>>>(c->ctrl, solve);
        c->launches++;
        unsigned launched = 0;
        unsigned batch = std::min(std::max(pred + 1u, 2u), max_it);
        const size_t prof_start = c->prof_recs.size();
        if (multi) { int rg = exchange_ghosts(c, gpos); if (rg) return rg; }   // warm-start kappa of the ghosts
        bool kappa_in_flight = false, p2p_kappa_pending = false;
        while (true) {
            for (unsigned b = 0; b < batch; ++b) {
                const int seq = (int)(launched + b);
                if (!multi) {
                    launch_accel(nullptr, 0, nullptr, seq);
                    launch_jacobi(solve, nullptr, 0, nullptr, 0, 1, seq);
                } else if (c->p2p) {
                    // peer-memory path, no NCCL call per iteration: a small kernel pushes the export values over NVLink and
                    // a one-thread kernel waits for the neighbours' flags (a per-block acquire at the start of the big
                    // kernels was measured slower: system-scope acquires flush the SM's L1); pass B ends with the fused
                    // all-reduce + loop control
                    if (seq > 0) { const PushArgs w = push_args(c, GH_POS); k_wait_flags<<<1, 1, 0, st>>>(w.wait_l, w.wait_r, w.seq_word); c->launches++; }
                    launch_accel(nullptr, 0, nullptr, seq);
                    { int rg = p2p_refresh(c, GH_ACC, true); if (rg) return rg; }
                    launch_jacobi(solve, nullptr, 0, nullptr, 0, 1, seq, GhostWait{nullptr, nullptr, 0u}, &c->pr);
                    { int rg = p2p_refresh(c, GH_POS, false); if (rg) return rg; }
                    p2p_kappa_pending = true;
                } else if (!c->overlap) {
                    // serial refresh: kappa -> pass A -> acceleration -> pass B -> all-reduce
                    if (seq > 0 && !(c->dbg_skip & 1)) { int rg = exchange_ghosts(c, gpos); if (rg) return rg; }
                    launch_accel(nullptr, 0, nullptr, seq);
                    if (!(c->dbg_skip & 1)) { int rg = exchange_ghosts(c, c->acc); if (rg) return rg; }
                    launch_jacobi(solve, nullptr, 0, nullptr, 0, 1, seq);
                    if (!(c->dbg_skip & 2)) NCCL_TRY(c, c->nccl.AllReduce(&c->ctrl->err_sum, &c->ctrl->err_sum, 1, ncclDouble, ncclSum, c->comm, st));
                    if (solve == SOLVE_DIV) k_solve_control<SOLVE_DIV><<<1, 1, 0, st>>>(c->ctrl, sp, c->sph);
                    else k_solve_control<SOLVE_PRESS><<<1, 1, 0, st>>>(c->ctrl, sp, c->sph);
                    c->launches++;
                } else {
                    // boundary-first: the export particles' results travel while the interior is being computed
                    if (kappa_in_flight) { int rl = land_ghost_kappa(); if (rl) return rl; }
                    const unsigned nbB = std::max(div_up(c->n_exp_all, DFSPH_JACOBI_BLOCK), 1u);
                    if (c->n_exp_all) launch_accel(c->exp_all, c->n_exp_all, nullptr, -1);
                    CUDA_TRY(c, cudaEventRecord(c->ev_a, st));
                    { int rg = exchange_ghosts_async(c, c->acc, c->acc + c->n, c->ev_a, c->ev_a2); if (rg) return rg; }
                    launch_accel(nullptr, 0, c->is_export, seq);
                    CUDA_TRY(c, cudaStreamWaitEvent(st, c->ev_a2, 0));
                    if (c->n_exp_all) launch_jacobi(solve, c->exp_all, c->n_exp_all, nullptr, 0, 0, -1);
                    CUDA_TRY(c, cudaEventRecord(c->ev_k, st));
                    { int rg = exchange_ghosts_async(c, gpos, c->ghost_stage, c->ev_k, c->ev_k2); if (rg) return rg; }
                    kappa_in_flight = true;
                    launch_jacobi(solve, nullptr, 0, c->is_export, c->n_exp_all ? nbB : 0u, 1, seq);
                    NCCL_TRY(c, c->nccl.AllReduce(&c->ctrl->err_sum, &c->ctrl->err_sum, 1, ncclDouble, ncclSum, c->comm, st));
                    if (solve == SOLVE_DIV) k_solve_control<SOLVE_DIV><<<1, 1, 0, st>>>(c->ctrl, sp, c->sph);
                    else k_solve_control<SOLVE_PRESS><<<1, 1, 0, st>>>(c->ctrl, sp, c->sph);
                    c->launches++;
                }
            }
            launched += batch;
            CUDA_TRY(c, cudaMemcpyAsync(c->h_ctrl, c->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(c, cudaStreamSynchronize(st));
            if (c->h_ctrl->done || launched >= max_it) break;
            batch = std::min(2u, max_it - launched);
        }
        if (kappa_in_flight) { int rl = land_ghost_kappa(); if (rl) return rl; }   // final kappa of the ghosts for the finaliser
        else if (p2p_kappa_pending) {
            const PushArgs w = push_args(c, GH_POS);
            k_wait_flags<<<1, 1, 0, st>>>(w.wait_l, w.wait_r, w.seq_word);
            c->launches++;
        }
        else if (multi) { int rg = exchange_ghosts(c, gpos); if (rg) return rg; }
        pred = c->h_ctrl->iter;
        // launches past the converged iteration exited immediately: keep them out of the per-kernel statistics
        for (size_t r = prof_start; r < c->prof_recs.size();) {
            if (c->prof_recs[r].seq >= (int)pred) {
                c->prof_pool.push_back(c->prof_recs[r].a); c->prof_pool.push_back(c->prof_recs[r].b);
                c->prof_recs.erase(c->prof_recs.begin() + r);
            } else ++r;
        }
        return 0;
    };

    const bool visc = c->par.viscosity_method == 1;
    trace_mark(c, "solver: init sweep");
    if (div) {
        // the reference's iteration is a no-op for an empty model: avg stays 0, one iteration is counted
        int rc = solve_loop(SOLVE_DIV, c->par.max_iterations_v, c->pred_iter_v);
        if (rc) return rc;
        trace_mark(c, "solver: divergence loop");
        ProfScope ps(c, DFSPH_B200_PROF_DIV_FINAL);
        if (visc) k_div_final<MODE, true, false><<<grid, DFSPH_BLOCK, 0, st>>>(f, c->sph, sp, c->ctrl);
        else k_div_final<MODE, true, true><<<grid, DFSPH_BLOCK, 0, st>>>(f, c->sph, sp, c->ctrl);
    } else if (!visc) {
        ProfScope ps(c, DFSPH_B200_PROF_DIV_FINAL);
        k_div_final<MODE, false, true><<<grid, DFSPH_BLOCK, 0, st>>>(f, c->sph, sp, c->ctrl);
    }
    if (visc) {
        // Simulation::computeNonPressureForces -> Viscosity_Standard::step (next-row f1), fused with clearAccelerations,
        // the CFL scan and the kick; needs the post-divergence-solve velocities of the neighbours
        if (multi) { int rg = exchange_ghosts(c, c->vel[c->cur]); if (rg) return rg; }   // ghost (v, rho)
        Real4* vtmp = c->vel[1 - c->cur];
        { ProfScope ps(c, DFSPH_B200_PROF_DIV_FINAL);
          k_viscosity_kick<MODE><<<grid, DFSPH_BLOCK, 0, st>>>(f, c->sph, sp, c->ctrl, c->bpos, vtmp); }
        if (n > 0) k_copy_real4<<<div_up(n, DFSPH_BLOCK), DFSPH_BLOCK, 0, st>>>(vtmp, c->vel[c->cur], n);
        c->launches += 2;
    }
    if (multi) {
        NCCL_TRY(c, c->nccl.AllReduce(&c->ctrl->maxvel_bits, &c->ctrl->maxvel_bits, 1, ncclUint64, ncclMax, c->comm, st));   // CFL maximum
        int rg = exchange_ghosts(c, c->vel[c->cur]); if (rg) return rg;                                                       // kicked velocities
    }
    trace_mark(c, "solver: div finaliser + kick + CFL/ghost v");
    k_update_time_step<<<1, 1, 0, st>>>(c->ctrl, sp);
    { ProfScope ps(c, DFSPH_B200_PROF_PRESS_INIT); k_press_init<MODE><<<grid, DFSPH_BLOCK, 0, st>>>(f, c->sph, c->ctrl); }
    c->launches += 3;
    trace_mark(c, "solver: pressure init");
    {
        int rc = solve_loop(SOLVE_PRESS, c->par.max_iterations, c->pred_iter);
        if (rc) return rc;
    }
    trace_mark(c, "solver: pressure loop");
    Real4* pos_out = c->pos[1 - c->cur_pos];
    { ProfScope ps(c, DFSPH_B200_PROF_PRESS_FINAL); k_press_final<MODE><<<grid, DFSPH_BLOCK, 0, st>>>(f, c->sph, c->ctrl, pos_out,
          c->final_x_stage, c->final_v_stage, multi ? nullptr : c->id[c->cur]); }
    c->final_stage_written = c->final_x_stage != nullptr && n > 0;
    k_step_end<<<1, 1, 0, st>>>(c->ctrl);
    c->launches += 2;
    c->cur_pos = 1 - c->cur_pos;
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

static int do_step(dfsph_b200_ctx* c, dfsph_b200_step_stats* stats)
{
    if (c->capacity_error) CTX_FAIL(c, DFSPH_B200_ERR_CAPACITY, "a neighbour list overflowed in an earlier step (need %u fluid / %u boundary slots, have %u / %u): "
                                    "the state is frozen at that step's start; raise max_*_neighbors and set the fluid again",
                                    c->h_ctrl->overflow, c->h_ctrl->overflow_b, c->Kf, c->Kb);
    int rc = prepare(c);
    if (rc) return rc;
    if (c->nb > 0 && !c->have_bvol) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "boundary volumes missing: pass V to add_boundary or call compute_boundary_volume");
    c->launches = 0;
    cudaStream_t st = c->stream;
    // divergence solver disabled -> iterationsV = 0 (TimeStepDFSPH.cpp:170)
    CUDA_TRY(c, cudaEventRecord(c->ev[0], st));
    trace_mark(c, "(start)");
    k_step_begin<<<1, 1, 0, st>>>(c->ctrl);
    c->launches++;
    if (!c->tables_valid) { rc = run_search(c); if (rc) return rc; }
    if (c->late_vel_stage) {   // step_host: the velocity upload overlapped the search; scatter it into the sorted order now
        CUDA_TRY(c, cudaStreamWaitEvent(st, c->ev_copy_in, 0));
        k_unpack3<<<std::max(div_up(c->n, 256), 1u), 256, 0, st>>>(c->late_vel_stage, c->vel[c->cur], c->n, c->id[c->cur], 0);
        c->launches++;
        c->late_vel_stage = nullptr;
    }
    CUDA_TRY(c, cudaEventRecord(c->ev[1], st));
#if DFSPH_REAL_IS_DOUBLE
    if (c->solver_mode == KM_CUBIC) rc = run_solver<KM_CUBIC>(c);
    else if (c->solver_mode == KM_GENERIC) rc = run_solver<KM_GENERIC>(c);
    else rc = run_solver<KM_LUT>(c);
#else
    rc = run_solver<KM_CUBIC_AVX>(c);
#endif
    if (rc) return rc;
    c->tables_valid = false;   // positions advanced; table describes the pre-advection positions (as in the reference)
    trace_mark(c, "solver: finaliser + advection");
    CUDA_TRY(c, cudaEventRecord(c->ev[2], st));
    if (stats) {
        CUDA_TRY(c, cudaMemcpyAsync(c->h_ctrl, c->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(c, cudaStreamSynchronize(st));
        prof_collect(c);
        trace_collect(c);
        const Ctrl& hc = *c->h_ctrl;
        memset(stats, 0, sizeof(*stats));
        stats->iterations = hc.iterations;
        stats->iterations_v = c->par.enable_divergence_solver ? hc.iterations_v : 0u;
        stats->avg_density_error = hc.avg_err;
        stats->avg_density_error_v = hc.avg_err_v;
        stats->time_step_size = (double)hc.h;
        stats->time = hc.time;
        stats->num_particles = c->n;
        stats->max_neighbors = hc.max_nbr;
        stats->gpu_launches = c->launches;
        if (c->use_graph && !c->multi && !c->profiling)   // kernels launched by the loop graphs: pass A, pass B, k_loop_cond per iteration + k_solve_begin
            stats->gpu_launches += 3u * (hc.iterations + stats->iterations_v) + (c->par.enable_divergence_solver ? 2u : 1u);
        else if (c->use_graph && c->multi && c->p2p && !c->profiling)   // slab loop graphs: + two waits and two pushes per iteration
            stats->gpu_launches += 7u * (hc.iterations + stats->iterations_v) + (c->par.enable_divergence_solver ? 2u : 1u);
        cudaEventElapsedTime(&stats->ms_search, c->ev[0], c->ev[1]);
        cudaEventElapsedTime(&stats->ms_solver, c->ev[1], c->ev[2]);
        c->par.time_step_size = hc.h;   // what get_params reports: the device's (CFL-adapted) step size
        if (hc.fatal) {
            c->capacity_error = true;
            CTX_FAIL(c, DFSPH_B200_ERR_CAPACITY, "neighbour table capacity exceeded: need %u fluid / %u boundary slots (have %u / %u); the step was not "
                     "carried out and the state is frozen; raise max_*_neighbors and set the fluid again", hc.overflow, hc.overflow_b, c->Kf, c->Kb);
        }
    }
    return DFSPH_B200_OK;
}

extern "C" {

int dfsph_b200_step(dfsph_b200_ctx* c, dfsph_b200_step_stats* stats)
{
    CHECK_CTX(c);
    return do_step(c, stats);
}

int dfsph_b200_search_and_density(dfsph_b200_ctx* c)
{
    CHECK_CTX(c);
    int rc = prepare(c);
    if (rc) return rc;
    rc = run_search(c);
    if (rc) return rc;
    const unsigned grid = std::max(div_up(c->n, DFSPH_BLOCK), 1u);
    FluidArrays f = fluid_arrays(c);
    // density only: run the fused sweep without the divergence part (factor is a by-product)
#if DFSPH_REAL_IS_DOUBLE
    if (c->solver_mode == KM_CUBIC) k_init_sweep<KM_CUBIC, false><<<grid, DFSPH_BLOCK, 0, c->stream>>>(f, c->sph, c->bpos, c->ctrl);
    else if (c->solver_mode == KM_GENERIC) k_init_sweep<KM_GENERIC, false><<<grid, DFSPH_BLOCK, 0, c->stream>>>(f, c->sph, c->bpos, c->ctrl);
    else k_init_sweep<KM_LUT, false><<<grid, DFSPH_BLOCK, 0, c->stream>>>(f, c->sph, c->bpos, c->ctrl);
#else
    k_init_sweep<KM_CUBIC_AVX, false><<<grid, DFSPH_BLOCK, 0, c->stream>>>(f, c->sph, c->bpos, c->ctrl);
#endif
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return DFSPH_B200_OK;
}

static size_t field_elem_bytes(dfsph_b200_field f)
{
    switch (f) {
        case DFSPH_B200_FIELD_POSITION: case DFSPH_B200_FIELD_VELOCITY: case DFSPH_B200_FIELD_PRESSURE_ACCEL: return 3 * sizeof(Real);
        case DFSPH_B200_FIELD_ID: case DFSPH_B200_FIELD_STATE: case DFSPH_B200_FIELD_NUM_NEIGHBORS: return sizeof(unsigned);
        default: return sizeof(Real);
    }
}

int dfsph_b200_download(dfsph_b200_ctx* c, dfsph_b200_field field, void* dst, size_t bytes, int by_id)
{
    CHECK_CTX(c);
    cudaSetDevice(c->cfg.device);
    if (!dst) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "dst is NULL");
    cudaStream_t st = c->stream;
    if (field == DFSPH_B200_FIELD_BOUNDARY_VOLUME) {
        if (bytes != c->h_bpos.size() * sizeof(Real)) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "size mismatch for boundary volume");
        Real* o = (Real*)dst;
        for (size_t i = 0; i < c->h_bpos.size(); ++i) o[i] = c->h_bpos[i].w;
        return DFSPH_B200_OK;
    }
    const unsigned n = c->n;
    const size_t eb = field_elem_bytes(field);
    if (bytes != (size_t)n * eb) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "size mismatch: field needs %zu bytes, got %zu", (size_t)n * eb, bytes);
    if (n == 0) return DFSPH_B200_OK;
    { int rc = ensure_stage(c, (size_t)n * eb); if (rc) return rc; }
    if (by_id && c->multi) CTX_FAIL(c, DFSPH_B200_ERR_UNSUPPORTED, "by_id transfers are not available in multi-GPU runs (ids are global): use by_id=0 and FIELD_ID");
    const unsigned* idmap = by_id ? c->id[c->cur] : nullptr;
    const unsigned g = div_up(n, 256);
    switch (field) {
        case DFSPH_B200_FIELD_POSITION: k_pack3<<<g, 256, 0, st>>>(c->pos[c->cur_pos], (Real*)c->stage, n, idmap); break;
        case DFSPH_B200_FIELD_VELOCITY: k_pack3<<<g, 256, 0, st>>>(c->vel[c->cur], (Real*)c->stage, n, idmap); break;
        case DFSPH_B200_FIELD_PRESSURE_ACCEL: k_pack3<<<g, 256, 0, st>>>(c->acc, (Real*)c->stage, n, idmap); break;
        case DFSPH_B200_FIELD_DENSITY: k_pack1<Real><<<g, 256, 0, st>>>(c->density, (Real*)c->stage, n, idmap); break;
        case DFSPH_B200_FIELD_FACTOR: k_pack1<Real><<<g, 256, 0, st>>>(c->factor, (Real*)c->stage, n, idmap); break;
        case DFSPH_B200_FIELD_DENSITY_ADV: k_pack1<Real><<<g, 256, 0, st>>>(c->density_adv, (Real*)c->stage, n, idmap); break;
        case DFSPH_B200_FIELD_KAPPA: k_pack1<Real><<<g, 256, 0, st>>>(c->kappa[c->cur], (Real*)c->stage, n, idmap); break;
        case DFSPH_B200_FIELD_KAPPA_V: k_pack1<Real><<<g, 256, 0, st>>>(c->kappa_v[c->cur], (Real*)c->stage, n, idmap); break;
        case DFSPH_B200_FIELD_ID: k_pack1<unsigned><<<g, 256, 0, st>>>(c->id[c->cur], (unsigned*)c->stage, n, idmap); break;
        case DFSPH_B200_FIELD_STATE: k_pack1<unsigned><<<g, 256, 0, st>>>(c->state[c->cur], (unsigned*)c->stage, n, idmap); break;
        case DFSPH_B200_FIELD_NUM_NEIGHBORS: k_pack1<unsigned><<<g, 256, 0, st>>>(c->nnbr, (unsigned*)c->stage, n, idmap); break;
        default: CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "unknown field %d", (int)field);
    }
    CUDA_TRY(c, cudaMemcpyAsync(dst, c->stage, bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    return DFSPH_B200_OK;
}

int dfsph_b200_upload(dfsph_b200_ctx* c, dfsph_b200_field field, const void* src, size_t bytes, int by_id)
{
    CHECK_CTX(c);
    cudaSetDevice(c->cfg.device);
    if (!src) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "src is NULL");
    const unsigned n = c->n;
    const size_t eb = field_elem_bytes(field);
    if (bytes != (size_t)n * eb) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "size mismatch: field needs %zu bytes, got %zu", (size_t)n * eb, bytes);
    if (n == 0) return DFSPH_B200_OK;
    { int rc = ensure_stage(c, (size_t)n * eb); if (rc) return rc; }
    cudaStream_t st = c->stream;
    CUDA_TRY(c, cudaMemcpyAsync(c->stage, src, bytes, cudaMemcpyHostToDevice, st));
    const unsigned* idmap = by_id ? c->id[c->cur] : nullptr;
    const unsigned g = div_up(n, 256);
    switch (field) {
        case DFSPH_B200_FIELD_POSITION: k_unpack3<<<g, 256, 0, st>>>((const Real*)c->stage, c->pos[c->cur_pos], n, idmap, 0); c->tables_valid = false; break;
        case DFSPH_B200_FIELD_VELOCITY: k_unpack3<<<g, 256, 0, st>>>((const Real*)c->stage, c->vel[c->cur], n, idmap, 0); break;
        case DFSPH_B200_FIELD_KAPPA: k_unpack1<Real><<<g, 256, 0, st>>>((const Real*)c->stage, c->kappa[c->cur], n, idmap); break;
        case DFSPH_B200_FIELD_KAPPA_V: k_unpack1<Real><<<g, 256, 0, st>>>((const Real*)c->stage, c->kappa_v[c->cur], n, idmap); break;
        case DFSPH_B200_FIELD_STATE: k_unpack1<unsigned><<<g, 256, 0, st>>>((const unsigned*)c->stage, c->state[c->cur], n, idmap); break;
        default: CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "field %d is not uploadable (solver output)", (int)field);
    }
    CUDA_TRY(c, cudaStreamSynchronize(st));
    return DFSPH_B200_OK;
}

int dfsph_b200_step_host(dfsph_b200_ctx* c, void* x_inout, void* v_inout, void* density_out, dfsph_b200_step_stats* stats)
{
    CHECK_CTX(c);
    cudaSetDevice(c->cfg.device);
    if (!x_inout || !v_inout) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "x/v buffers are NULL");
    // single GPU: rows are addressed by particle id (= host array index).  Multi-GPU: ids are global, so rows are in this
    // rank's device order; the buffers hold dfsph_b200_capacity() rows, the first num_particles() of them are read,
    // and after the step (migration may have changed the count) the first stats->num_particles rows are written.
    const bool multi = c->multi;
    const unsigned n = c->n;
    const unsigned rows = multi ? std::max(c->cap, n) : n;
    const size_t b3 = (size_t)n * 3 * sizeof(Real);
    { int rc = ensure_stage(c, (size_t)rows * 7 * sizeof(Real)); if (rc) return rc; }
    cudaStream_t st = c->stream;
    if (!c->copy_stream) {
        CUDA_TRY(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_copy_in, cudaEventDisableTiming));
        CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_density, cudaEventDisableTiming));
    }
    cudaStream_t cs = c->copy_stream;
    Real* sx = (Real*)c->stage;
    Real* sv = sx + (size_t)rows * 3;
    Real* sd = sv + (size_t)rows * 3;
    const unsigned g = std::max(div_up(n, 256), 1u);
    if (n > 0) {
        CUDA_TRY(c, cudaMemcpyAsync(sx, x_inout, b3, cudaMemcpyHostToDevice, st));
        if (!multi) {
            // x first (the search needs it); v follows on the copy stream and is scattered after the reorder (do_step)
            CUDA_TRY(c, cudaEventRecord(c->ev_copy_in, st));
            k_unpack3<<<g, 256, 0, st>>>(sx, c->pos[c->cur_pos], n, c->id[c->cur], 0);
            CUDA_TRY(c, cudaStreamWaitEvent(cs, c->ev_copy_in, 0));   // keep the two uploads back to back on the H2D engine
            CUDA_TRY(c, cudaMemcpyAsync(sv, v_inout, b3, cudaMemcpyHostToDevice, cs));
            CUDA_TRY(c, cudaEventRecord(c->ev_copy_in, cs));
            c->late_vel_stage = sv;
        } else {
            // the migration of the search packs velocities, so both fields land before the step
            k_unpack3<<<g, 256, 0, st>>>(sx, c->pos[c->cur_pos], n, nullptr, 0);
            CUDA_TRY(c, cudaMemcpyAsync(sv, v_inout, b3, cudaMemcpyHostToDevice, st));
            k_unpack3<<<g, 256, 0, st>>>(sv, c->vel[c->cur], n, nullptr, 0);
        }
        c->tables_valid = false;
    }
    if (density_out) { c->early_density_stage = sd; c->early_density_host = density_out; }
    // the staged inputs are consumed before the finaliser runs (x by the sort, v after the reorder), so the finaliser can
    // leave the packed results in the same staging rows
    c->final_x_stage = sx; c->final_v_stage = sv; c->final_stage_written = false;
    int rc = do_step(c, stats);
    const bool packed = c->final_stage_written;
    c->late_vel_stage = nullptr; c->early_density_host = nullptr;
    c->final_x_stage = c->final_v_stage = nullptr; c->final_stage_written = false;
    if (rc) { cudaStreamSynchronize(cs); return rc; }
    const unsigned n1 = c->n;   // (multi-GPU: after migration)
    if (n1 > 0) {
        const unsigned g1 = div_up(n1, 256);
        const size_t b31 = (size_t)n1 * 3 * sizeof(Real);
        const unsigned* idmap = multi ? nullptr : c->id[c->cur];
        if (!packed) {
            k_pack3<<<g1, 256, 0, st>>>(c->pos[c->cur_pos], sx, n1, idmap);
            k_pack3<<<g1, 256, 0, st>>>(c->vel[c->cur], sv, n1, idmap);
        }
        CUDA_TRY(c, cudaMemcpyAsync(x_inout, sx, b31, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(c, cudaMemcpyAsync(v_inout, sv, b31, cudaMemcpyDeviceToHost, st));
    }
    if (stats) stats->gpu_launches += (n > 0 ? (multi ? 2u : 1u) : 0u) + (n1 > 0 && !packed ? 2u : 0u);   // unpacks (+ packs) issued here
    CUDA_TRY(c, cudaStreamSynchronize(cs));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    return DFSPH_B200_OK;
}

uint64_t dfsph_b200_capacity(const dfsph_b200_ctx* c) { return c ? c->cap : 0; }

}  // extern "C"

// ---- neighbour export (tests / non-ported host code) ---------------------------------------------------------------
__global__ void k_export_neighbors(unsigned n, const unsigned* __restrict__ tab, unsigned K, const unsigned* __restrict__ cnt,
                                   const unsigned long long* __restrict__ offsets, const unsigned* __restrict__ remap, unsigned* __restrict__ out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned m = cnt[i];
    const unsigned* t = tab + (size_t)(i >> 5) * K * DFSPH_TILE + (i & 31u);
    unsigned* o = out + offsets[i];
    for (unsigned k = 0; k < m; ++k) {   // insertion sort -> ascending lists
        unsigned v = t[(size_t)k * DFSPH_TILE];
        if (remap) v = remap[v];
        unsigned b = k;
        while (b > 0 && o[b - 1] > v) { o[b] = o[b - 1]; --b; }
        o[b] = v;
    }
}

extern "C" {

int dfsph_b200_neighbors(dfsph_b200_ctx* c, int other, uint32_t* counts, uint64_t* offsets, uint32_t* idx, uint64_t cap)
{
    CHECK_CTX(c);
    if (other != 0 && other != 1) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "other must be 0 (fluid) or 1 (boundary)");
    if (!counts) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "counts is NULL");
    int rc = prepare(c);
    if (rc) return rc;
    if (!c->tables_valid) { rc = run_search(c); if (rc) return rc; }
    const unsigned n = c->n;
    if (n == 0) { if (offsets) offsets[0] = 0; return DFSPH_B200_OK; }
    const unsigned* cnt = other == 0 ? c->cnt_f : c->cnt_b;
    CUDA_TRY(c, cudaMemcpyAsync(counts, cnt, (size_t)n * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->h_ctrl, c->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->h_ctrl->fatal) {   // the lists are truncated at the table capacity: say so instead of handing them out
        c->capacity_error = true;
        CTX_FAIL(c, DFSPH_B200_ERR_CAPACITY, "neighbour table capacity exceeded: need %u fluid / %u boundary slots (have %u / %u); "
                 "raise max_*_neighbors and set the fluid again", c->h_ctrl->overflow, c->h_ctrl->overflow_b, c->Kf, c->Kb);
    }
    if (!idx) return DFSPH_B200_OK;
    if (!offsets) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "offsets is NULL");
    std::vector<unsigned long long> off(n + 1);
    off[0] = 0;
    for (unsigned i = 0; i < n; ++i) off[i + 1] = off[i] + counts[i];
    for (unsigned i = 0; i <= n; ++i) offsets[i] = off[i];
    if (off[n] > cap) CTX_FAIL(c, DFSPH_B200_ERR_CAPACITY, "idx capacity %llu < %llu entries", (unsigned long long)cap, off[n]);
    if (off[n] == 0) return DFSPH_B200_OK;
    unsigned long long* d_off = nullptr;
    unsigned* d_out = nullptr;
    CUDA_TRY(c, cudaMalloc((void**)&d_off, (size_t)(n + 1) * sizeof(unsigned long long)));
    cudaError_t e = cudaMalloc((void**)&d_out, (size_t)off[n] * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_off, off.data(), (size_t)(n + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        k_export_neighbors<<<div_up(n, 256), 256, 0, c->stream>>>(n, other == 0 ? c->tab_f : c->tab_b, other == 0 ? c->Kf : c->Kb, cnt, d_off,
                                                                 other == 0 ? nullptr : c->borig, d_out);
        e = cudaMemcpyAsync(idx, d_out, (size_t)off[n] * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_off);
    if (d_out) cudaFree(d_out);
    if (e != cudaSuccess) { c->sticky = 1; CTX_FAIL(c, DFSPH_B200_ERR_CUDA, "neighbors: %s", cudaGetErrorString(e)); }
    return DFSPH_B200_OK;
}

}  // extern "C"

// ---- kernel function evaluation (KernelTests.cpp on the device) -------------------------------------------------------
template <int MODE>
__global__ void k_eval_kernel(SphConst c, unsigned n, const Real* __restrict__ r, Real* __restrict__ W, Real* __restrict__ gW)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Real rx = r[3 * (size_t)i], ry = r[3 * (size_t)i + 1], rz = r[3 * (size_t)i + 2];
    const Real r2 = rx * rx + ry * ry + rz * rz;
    if (W) W[i] = sph_W<MODE>(c, r2);
    if (gW) {
        const Real g = sph_gradW_scale<MODE>(c, r2);
        gW[3 * (size_t)i] = g * rx; gW[3 * (size_t)i + 1] = g * ry; gW[3 * (size_t)i + 2] = g * rz;
    }
}

extern "C" {

// kernel: 0 cubic (scalar arithmetic), 1 Wendland quintic C2, 2 Poly6, 3 Spiky, 4 precomputed cubic,
// -1 = the solver's own kernel (CubicKernel_AVX in f32; the configured kernel / gradKernel pair in f64)
int dfsph_b200_eval_kernel(dfsph_b200_ctx* c, int kernel, uint64_t n64, const void* r, void* W, void* gradW)
{
    CHECK_CTX(c);
    cudaSetDevice(c->cfg.device);
    if (!r || (!W && !gradW)) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "null argument");
    if (c->lutW == nullptr) {
        const double v = c->volume; if (v <= 0.0) c->volume = 1.0;
        int rc = setup_constants(c); c->volume = v; if (rc) return rc;
    }
    int mode;
    if (kernel == DFSPH_B200_KERNEL_CUBIC) mode = KM_CUBIC;
    else if (kernel == DFSPH_B200_KERNEL_PRECOMPUTED_CUBIC) mode = KM_LUT;
    else if (kernel == -1) mode = c->solver_mode;
    else if (kernel >= 1 && kernel <= 3) mode = KM_GENERIC;
    else CTX_FAIL(c, DFSPH_B200_ERR_UNSUPPORTED, "unsupported kernel id %d", kernel);
    SphConst sc = c->sph;
    if (kernel >= 1 && kernel <= 3) { sc.w_kind = kernel; sc.g_kind = kernel; }
    const unsigned n = (unsigned)n64;
    if (n == 0) return DFSPH_B200_OK;
    Real *dr = nullptr, *dW = nullptr, *dG = nullptr;
    CUDA_TRY(c, cudaMalloc((void**)&dr, (size_t)n * 3 * sizeof(Real)));
    cudaError_t e = cudaMalloc((void**)&dW, (size_t)n * sizeof(Real));
    if (e == cudaSuccess) e = cudaMalloc((void**)&dG, (size_t)n * 3 * sizeof(Real));
    if (e == cudaSuccess) e = cudaMemcpy(dr, r, (size_t)n * 3 * sizeof(Real), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        const unsigned g = div_up(n, 256);
        if (mode == KM_CUBIC_AVX) k_eval_kernel<KM_CUBIC_AVX><<<g, 256, 0, c->stream>>>(sc, n, dr, dW, dG);
        else if (mode == KM_CUBIC) k_eval_kernel<KM_CUBIC><<<g, 256, 0, c->stream>>>(sc, n, dr, dW, dG);
        else if (mode == KM_GENERIC) k_eval_kernel<KM_GENERIC><<<g, 256, 0, c->stream>>>(sc, n, dr, dW, dG);
        else k_eval_kernel<KM_LUT><<<g, 256, 0, c->stream>>>(sc, n, dr, dW, dG);
        e = cudaStreamSynchronize(c->stream);
    }
    if (e == cudaSuccess && W) e = cudaMemcpy(W, dW, (size_t)n * sizeof(Real), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && gradW) e = cudaMemcpy(gradW, dG, (size_t)n * 3 * sizeof(Real), cudaMemcpyDeviceToHost);
    cudaFree(dr); if (dW) cudaFree(dW); if (dG) cudaFree(dG);
    if (e != cudaSuccess) { c->sticky = 1; CTX_FAIL(c, DFSPH_B200_ERR_CUDA, "eval_kernel: %s", cudaGetErrorString(e)); }
    return DFSPH_B200_OK;
}

// Host placement for the host-buffer path on multi-socket boxes: run the calling thread on the CPUs of the GPU's NUMA
// node and prefer that node's memory, so that pinned buffers allocated afterwards are local to the GPU's PCIe root
// (8 ranks staging through one socket's memory were the reason the end-to-end rate stopped scaling).  Plain sysfs +
// syscalls, no libnuma.  Returns the node (>= 0), or -1 when the topology is unknown / the binding is not permitted (the
// process then simply stays where it is).
int dfsph_b200_bind_host_numa(int device)
{
    char bus[64] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char* q = bus; *q; ++q) *q = (char)tolower((unsigned char)*q);
    char path[256];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    int node = -1;
    if (FILE* f = fopen(path, "r")) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
    if (node < 0) return -1;
    // CPUs of the node, intersected with what this process may use
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    cpu_set_t allowed, want;
    CPU_ZERO(&want);
    if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) CPU_ZERO(&allowed);
    int picked = 0;
    if (FILE* f = fopen(path, "r")) {
        int a, b;
        while (fscanf(f, "%d", &a) == 1) {
            b = a;
            int ch = fgetc(f);
            if (ch == '-') { if (fscanf(f, "%d", &b) != 1) b = a; ch = fgetc(f); }
            for (int k = a; k <= b && k < CPU_SETSIZE; ++k) if (CPU_ISSET(k, &allowed)) { CPU_SET(k, &want); ++picked; }
            if (ch != ',') break;
        }
        fclose(f);
    }
    if (picked > 0) sched_setaffinity(0, sizeof(want), &want);
    // memory policy: prefer the node for every later allocation of this thread (MPOL_PREFERRED = 1)
    if (node < 1024) {
        unsigned long mask[16] = {0};
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, (unsigned long)(sizeof(mask) * 8));
    }
    return node;
}

// pinned host buffers for the host-buffer (e2e) path
void* dfsph_b200_alloc_pinned(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}
void dfsph_b200_free_pinned(void* p) { if (p) cudaFreeHost(p); }
int dfsph_b200_host_register(void* p, size_t bytes)
{
    if (!p || bytes == 0) return DFSPH_B200_ERR_INVALID;
    if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return DFSPH_B200_ERR_CUDA; }
    return DFSPH_B200_OK;
}
int dfsph_b200_host_unregister(void* p)
{
    if (!p) return DFSPH_B200_ERR_INVALID;
    if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return DFSPH_B200_ERR_CUDA; }
    return DFSPH_B200_OK;
}

// ---- multi-GPU ---------------------------------------------------------------------------------------------------
// id256: two NCCL unique ids (one communicator for reductions / counts / migration, one for the halo refresh that
// runs concurrently on the communication stream)
int dfsph_b200_comm_get_unique_id(void* id256)
{
    static NcclApi api;
    std::string err;
    if (!id256 || !api.load(err)) { g_create_error = err.empty() ? "null argument" : err; return DFSPH_B200_ERR_COMM; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    for (int k = 0; k < 2; ++k) {
        ncclUniqueId id;
        if (api.GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return DFSPH_B200_ERR_COMM; }
        memcpy((char*)id256 + 128 * k, &id, 128);
    }
    return DFSPH_B200_OK;
}

int dfsph_b200_comm_init(dfsph_b200_ctx* c, const void* id128 /* the 256 bytes of comm_get_unique_id */, int rank, int world, int axis, double slab_lo, double slab_hi)
{
    CHECK_CTX(c);
    cudaSetDevice(c->cfg.device);
    if (!id128 || world < 1 || rank < 0 || rank >= world) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "bad rank/world");
    if (c->cap != 0) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "dfsph_b200_comm_init must be called before set_fluid");
    bool given = false;
    for (int k = 0; k < 3; ++k) if (c->cfg.domain_max[k] > c->cfg.domain_min[k]) given = true;
    if (!given) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "multi-GPU runs need an explicit cell-grid domain (config.domain_min/max) shared by all ranks");
    if (!(slab_hi > slab_lo)) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "empty slab");
    if (axis < 0 || axis > 2) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "slab axis must be 0, 1 or 2");
    // ghosts only ever come from the two adjacent ranks: an interior slab must be at least one cell (support radius) wide
    if (slab_lo > -1e299 && slab_hi < 1e299 && (slab_hi - slab_lo) < 4.0 * c->cfg.particle_radius * (1.0 + 1.0e-5))
        CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "slab [%g, %g) is narrower than one support radius", slab_lo, slab_hi);
    c->slab_axis = axis;
    if (world == 1) return DFSPH_B200_OK;
    std::string err;
    if (!c->nccl.load(err)) CTX_FAIL(c, DFSPH_B200_ERR_COMM, "%s", err.c_str());
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NCCL_TRY(c, c->nccl.CommInitRank(&c->comm, world, id, rank));
    ncclUniqueId id2;
    memcpy(&id2, (const char*)id128 + 128, 128);
    NCCL_TRY(c, c->nccl.CommInitRank(&c->comm2, world, id2, rank));
    CUDA_TRY(c, cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming));
    CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_a2, cudaEventDisableTiming));
    CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_k, cudaEventDisableTiming));
    CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_k2, cudaEventDisableTiming));
    c->multi = true;
    if (const char* e = getenv("DFSPH_B200_NO_OVERLAP")) c->overlap = !(e[0] == '1');
    if (const char* e = getenv("DFSPH_B200_DEBUG_SKIP")) c->dbg_skip = atoi(e);
    c->rank = rank; c->world = world;
    c->has_left = rank > 0; c->has_right = rank < world - 1;
    c->slab_lo = slab_lo; c->slab_hi = slab_hi;
    const unsigned one = 1u;
    CUDA_TRY(c, cudaMemcpy(&c->ctrl->multi, &one, sizeof(one), cudaMemcpyHostToDevice));
    return DFSPH_B200_OK;
}

// P2P blob layout: 8 cudaIpcMemHandle_t (pos[0], pos[1], vel[0], vel[1], acc, flags, reduce values, reduce sequence words) = 512 bytes
// like CUDA_TRY but not sticky: a failed peer mapping only means "stay on the NCCL path"
#define CUDA_SOFT(ctx, expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cudaGetLastError(); \
    CTX_FAIL(ctx, DFSPH_B200_ERR_COMM, "%s failed: %s", #expr, cudaGetErrorString(_e)); } } while (0)
int dfsph_b200_p2p_export(dfsph_b200_ctx* c, void* blob512)
{
    CHECK_CTX(c);
    cudaSetDevice(c->cfg.device);
    if (!blob512) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "null blob");
    if (!c->multi || c->cap == 0) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "call comm_init and set_fluid first");
    // words 0..5: arrival flags (kind x side), 6: ticket of the push kernel, 8..10: sequence number per kind of refresh
    if (!c->flags) { if (dev_alloc(c, &c->flags, 16)) return DFSPH_B200_ERR_CUDA; CUDA_SOFT(c, cudaMemset(c->flags, 0, 16 * sizeof(unsigned))); }
    if (!c->push_desc) { if (dev_alloc(c, &c->push_desc, 1)) return DFSPH_B200_ERR_CUDA; CUDA_SOFT(c, cudaMemset(c->push_desc, 0, sizeof(PushDesc))); }
    memset(blob512, 0, 512);
    cudaIpcMemHandle_t* h = (cudaIpcMemHandle_t*)blob512;
    CUDA_SOFT(c, cudaIpcGetMemHandle(&h[0], c->pos[0]));
    CUDA_SOFT(c, cudaIpcGetMemHandle(&h[1], c->pos[1]));
    CUDA_SOFT(c, cudaIpcGetMemHandle(&h[2], c->vel[0]));
    CUDA_SOFT(c, cudaIpcGetMemHandle(&h[3], c->vel[1]));
    CUDA_SOFT(c, cudaIpcGetMemHandle(&h[4], c->acc));
    CUDA_SOFT(c, cudaIpcGetMemHandle(&h[5], c->flags));
    if (!c->red_val) {
        if (dev_alloc(c, &c->red_val, 2 * DFSPH_MAX_RANKS)) return DFSPH_B200_ERR_CUDA;
        if (dev_alloc(c, &c->red_seq, DFSPH_MAX_RANKS)) return DFSPH_B200_ERR_CUDA;
        CUDA_SOFT(c, cudaMemset(c->red_val, 0, 2 * DFSPH_MAX_RANKS * sizeof(double)));
        CUDA_SOFT(c, cudaMemset(c->red_seq, 0, DFSPH_MAX_RANKS * sizeof(unsigned)));
    }
    CUDA_SOFT(c, cudaIpcGetMemHandle(&h[6], c->red_val));
    CUDA_SOFT(c, cudaIpcGetMemHandle(&h[7], c->red_seq));
    return DFSPH_B200_OK;
}

int dfsph_b200_p2p_import(dfsph_b200_ctx* c, const void* blobs_all)
{
    CHECK_CTX(c);
    cudaSetDevice(c->cfg.device);
    if (!c->multi || !c->flags) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "call p2p_export first");
    if (!blobs_all) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "null blobs");
    invalidate_graphs(c);   // peer pointers are kernel arguments of the slab loop graphs
    if (c->world > DFSPH_MAX_RANKS) CTX_FAIL(c, DFSPH_B200_ERR_UNSUPPORTED, "peer-memory path supports up to %d ranks", DFSPH_MAX_RANKS);
    const char* all = (const char*)blobs_all;
    const void* blobs[2] = { c->has_left ? all + 512 * (c->rank - 1) : nullptr, c->has_right ? all + 512 * (c->rank + 1) : nullptr };
    // all-to-all tables of the fused density-error all-reduce
    memset(&c->pr, 0, sizeof(c->pr));
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) { c->pr.val[r] = c->red_val; c->pr.seq[r] = c->red_seq; continue; }
        const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)(all + 512 * r);
        void *pv = nullptr, *ps = nullptr;
        CUDA_SOFT(c, cudaIpcOpenMemHandle(&pv, h[6], cudaIpcMemLazyEnablePeerAccess));
        CUDA_SOFT(c, cudaIpcOpenMemHandle(&ps, h[7], cudaIpcMemLazyEnablePeerAccess));
        c->red_opened.push_back(pv); c->red_opened.push_back(ps);
        c->pr.val[r] = (double*)pv; c->pr.seq[r] = (unsigned*)ps;
    }
    c->pr.rank = c->rank; c->pr.world = c->world;
    for (int s = 0; s < 2; ++s) {
        const bool need = s == 0 ? c->has_left : c->has_right;
        if (!need) continue;
        if (!blobs[s]) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "missing neighbour blob");
        const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)blobs[s];
        void* p[6];
        for (int k = 0; k < 6; ++k) CUDA_SOFT(c, cudaIpcOpenMemHandle(&p[k], h[k], cudaIpcMemLazyEnablePeerAccess));
        c->peer[s].pos[0] = (Real4*)p[0]; c->peer[s].pos[1] = (Real4*)p[1];
        c->peer[s].vel[0] = (Real4*)p[2]; c->peer[s].vel[1] = (Real4*)p[3];
        c->peer[s].acc = (Real4*)p[4]; c->peer[s].flags = (unsigned*)p[5];
        c->peer[s].open = true;
    }
    c->p2p = true;
    return DFSPH_B200_OK;
}

int dfsph_b200_p2p_disable(dfsph_b200_ctx* c)
{
    CHECK_CTX(c);
    c->p2p = false;     // back to NCCL send/recv + ncclAllReduce (all ranks must switch together, between steps)
    return DFSPH_B200_OK;
}

int dfsph_b200_set_profiling(dfsph_b200_ctx* c, int on)
{
    CHECK_CTX(c);
    c->profiling = on != 0;
    if (on) { for (int k = 0; k < DFSPH_B200_PROF_CLASSES; ++k) { c->prof_ms[k] = 0.0; c->prof_count[k] = 0; } }
    return DFSPH_B200_OK;
}

int dfsph_b200_get_profile(dfsph_b200_ctx* c, double* ms, uint64_t* count)
{
    CHECK_CTX(c);
    cudaSetDevice(c->cfg.device);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    prof_collect(c);
    for (int k = 0; k < DFSPH_B200_PROF_CLASSES; ++k) { if (ms) ms[k] = c->prof_ms[k]; if (count) count[k] = c->prof_count[k]; }
    return DFSPH_B200_OK;
}

int dfsph_b200_timer_start(dfsph_b200_ctx* c)
{
    CHECK_CTX(c);
    cudaSetDevice(c->cfg.device);
    if (!c->timer_a) { CUDA_TRY(c, cudaEventCreate(&c->timer_a)); CUDA_TRY(c, cudaEventCreate(&c->timer_b)); }
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaEventRecord(c->timer_a, c->stream));
    return DFSPH_B200_OK;
}

int dfsph_b200_timer_stop(dfsph_b200_ctx* c, float* ms)
{
    CHECK_CTX(c);
    cudaSetDevice(c->cfg.device);
    if (!c->timer_a || !ms) CTX_FAIL(c, DFSPH_B200_ERR_INVALID, "timer_start has not been called");
    CUDA_TRY(c, cudaEventRecord(c->timer_b, c->stream));
    CUDA_TRY(c, cudaEventSynchronize(c->timer_b));
    CUDA_TRY(c, cudaEventElapsedTime(ms, c->timer_a, c->timer_b));
    return DFSPH_B200_OK;
}

int dfsph_b200_synchronize(dfsph_b200_ctx* c)
{
    CHECK_CTX(c);
    cudaSetDevice(c->cfg.device);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->ctrl && !c->capacity_error) {
        // steps issued without statistics do not synchronise: a neighbour-list overflow is reported here at the latest
        CUDA_TRY(c, cudaMemcpy(c->h_ctrl, c->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost));
        if (c->h_ctrl->fatal) c->capacity_error = true;
    }
    if (c->capacity_error) CTX_FAIL(c, DFSPH_B200_ERR_CAPACITY, "neighbour table capacity exceeded: need %u fluid / %u boundary slots (have %u / %u); "
                                    "the state is frozen at the failed step's start", c->h_ctrl->overflow, c->h_ctrl->overflow_b, c->Kf, c->Kb);
    return DFSPH_B200_OK;
}

}  // extern "C"
