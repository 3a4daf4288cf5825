// DFSPH solver kernels.  Every kernel is one thread per fluid particle ("writer owns particle
// i", race-free by construction) sweeping that particle's neighbour table.  Per step:
//   k_init_sweep      K1+K2+K3: computeDensities (TimeStep.cpp:54-169) + computeDFSPHFactor (TimeStepDFSPH.cpp:735-825 /
//                     1106-1186) + computeDensityChange and the divergence warm start (:894-950 / 1247-1295, :410-461)
//                     in ONE sweep -- all three depend only on x, v of the step start.  Also the only kernel that
//                     touches boundary particles: it leaves G_i = sum_b V_b gradW_ib in `bgrad`, which is all later
//                     kernels need of a static Akinci boundary.
//   k_accel           pass A of an iteration: computePressureAccel (:954-1039 / 1299-1367)
//   k_jacobi<DIV|PRESS>  pass B: compute_aij_pj + Jacobi update + density-error reduction + loop control
//                     (:1042-1100 / 1370-1420, :576-612, :653-699, :324-341, :477-495), decided on the device.
//   k_div_final       divergence finaliser (:500-539) + clearAccelerations (TimeStep.cpp:35-50) + CFL scan
//                     (Simulation.cpp:415-493) + v += h a (TimeStepDFSPH.cpp:192-208)
//   k_press_init      computeDensityAdv + pressure warm start (:830-889 / 1191-1242, :272-307)
//   k_press_final     pressure finaliser (:351-382) + x += h v (:220-237)
// Float build = semantics of the reference's AVX variant, double build = scalar variant.
#pragma once
#include "common.cuh"
#include "sph_kernels.cuh"

#define DFSPH_EPS ((Real)1.0e-5)   /* TimeStepDFSPH::m_eps, TimeStepDFSPH.h:28 */

// Multi-GPU with peer memory: flags the kernel has to see before it may read ghost values, and the all-to-all table of
// the fused density-error all-reduce (every rank stores its partial sum into every rank's table over NVLink).
#define DFSPH_MAX_RANKS 16
struct GhostWait { const unsigned* left; const unsigned* right; unsigned seq; };
struct PeerReduce {
    double* val[DFSPH_MAX_RANKS];      // val[r]: table of rank r (device pointer, peer-mapped), layout [2 parities][DFSPH_MAX_RANKS]
    unsigned* seq[DFSPH_MAX_RANKS];    // seq[r]: sequence words of rank r, layout [DFSPH_MAX_RANKS]
    int rank, world;                   // world == 0: not in use
};

__device__ __forceinline__ void ghost_wait(const GhostWait& w)
{
    if (w.left == nullptr && w.right == nullptr) return;
    if (threadIdx.x == 0) {
        unsigned v;
        if (w.left) do { asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(w.left) : "memory"); } while ((int)(v - w.seq) < 0);
        if (w.right) do { asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(w.right) : "memory"); } while ((int)(v - w.seq) < 0);
    }
    __syncthreads();
}

struct FluidArrays {
    Real4* pos;          // (x, y, z, kappa of the running solve); element [n] is the far-away sentinel particle
    Real4* vel;          // (vx, vy, vz, -); [n] = 0
    Real4* acc;          // pressure acceleration (ax, ay, az, -); [n] = 0
    Real4* bgrad;        // G_i = sum_b V_b gradW(x_i - x_b)
    Real* density;
    Real* factor;
    Real* density_adv;
    Real* kappa;         // persistent warm-start value p/rho^2   (time-step independent form)
    Real* kappa_v;       // persistent warm-start value p_v/rho^2
    unsigned* state;
    unsigned* nnbr;      // fluid + boundary neighbour count (particle-deficiency test)
    const unsigned* tab_f;
    const unsigned* cnt_f;    // per particle
    const unsigned* tcnt_f;   // per warp tile: max count of the tile rounded up to DFSPH_PAD; slots beyond a particle's
    const unsigned* tab_b;    //   own count hold the sentinel index, so sweeps run warp-uniform, branch-free loops
    const unsigned* cnt_b;
    const unsigned* tcnt_b;
    unsigned Kf, Kb;
    unsigned n;
    const unsigned* n_dev;    // slab loop graphs: the owned count lives in device memory (migration changes it every step); else null
    cudaTextureObject_t acc_tex;   // linear textures over acc / pos / vel (0 = none: plain loads).  Scattered gathers through the
    cudaTextureObject_t pos_tex;   // texture unit cost fewer L1 data-pipe wavefronts than LDG.128 (profiles/r2_tex_microbench.md)
    cudaTextureObject_t vel_tex;
};

// Which L1 front end each scattered gather uses (1 = texture unit, 0 = LSU).  Compile-time so that tuning builds can be
// compared (tools/variant_bench.py).
#ifndef DFSPH_TEX_POS_A
#define DFSPH_TEX_POS_A 0   /* (x_j, kappa_j) in pressure_accel: pass A and the two finalisers */
#endif
#ifndef DFSPH_TEX_POS_B
#define DFSPH_TEX_POS_B 0   /* x_j in pass B */
#endif
#ifndef DFSPH_TEX_ACC_B
#define DFSPH_TEX_ACC_B 1   /* a_j in pass B */
#endif
#ifndef DFSPH_TEX_POS_V
#define DFSPH_TEX_POS_V 0   /* x_j in the two-array (x, v) sweeps: init sweep, pressure init, viscosity */
#endif
#ifndef DFSPH_TEX_VEL
#define DFSPH_TEX_VEL 1     /* v_j in those sweeps */
#endif

// PLAIN: the array has a benign in-kernel writer (pos.w), so the LSU fallback must not use the non-coherent path
template <bool USE_TEX, bool PLAIN>
__device__ __forceinline__ Real4 gather4(const Real4* __restrict__ base, cudaTextureObject_t tex, unsigned j)
{
    if (USE_TEX && tex) {
#if DFSPH_REAL_IS_DOUBLE
        const int4 t0 = tex1Dfetch<int4>(tex, (int)(2u * j)), t1 = tex1Dfetch<int4>(tex, (int)(2u * j + 1u));   // 32 B record = two texels
        return make_real4(__hiloint2double(t0.y, t0.x), __hiloint2double(t0.w, t0.z), __hiloint2double(t1.y, t1.x), __hiloint2double(t1.w, t1.z));
#else
        const float4 t = tex1Dfetch<float4>(tex, (int)j);
        return make_real4(t.x, t.y, t.z, t.w);
#endif
    }
    return PLAIN ? ld_plain(base + j) : ld_gather(base + j);
}

// Table rows of particle i: a CTA-uniform 64-bit base (first tile of the CTA) plus a 32-bit per-thread element offset, so that
// the sweep loop addresses the index rows as [uniform register + 32-bit register] and advances one integer per batch (with a
// per-thread 64-bit pointer the compiler re-derived the address from the kernel parameters in every batch: ~10 of the 120
// instructions of a 4-pair batch of pass A).  CTAs start at multiples of 32 particles (block sizes are multiples of 32).
#ifndef DFSPH_TAB_UNIFORM
#define DFSPH_TAB_UNIFORM 1
#endif
struct TabRows { const unsigned* base; unsigned off; };
__device__ __forceinline__ TabRows tab_rows(const unsigned* tab, unsigned K, unsigned i)
{
#if !DFSPH_TAB_UNIFORM
    return TabRows{tab + (size_t)(i >> 5) * K * DFSPH_TILE + (i & 31u), 0u};
#endif
    const unsigned tile0 = (blockIdx.x * blockDim.x) >> 5;          // uniform over the CTA
    TabRows r;
    r.base = tab + (size_t)tile0 * K * DFSPH_TILE;
    r.off = ((i >> 5) - tile0) * K * DFSPH_TILE + (i & 31u);
    return r;
}
// kernels that run over an index list (multi-GPU export lists) have no CTA-uniform tile
__device__ __forceinline__ TabRows tab_rows_any(const unsigned* tab, unsigned K, unsigned i)
{
    TabRows r;
    r.base = tab + (size_t)(i >> 5) * K * DFSPH_TILE;
    r.off = i & 31u;
    return r;
}

// Index rows are streamed once per sweep: which cache policy the row loads use is a tuning switch (0 = read-only path,
// 1 = evict-first `ld.global.cs`, 2 = `L1::no_allocate`), so that the streamed rows need not displace the gathered records.
// Measurement build: pass A / pass B keep their index stream, gathers and stores but do no pair arithmetic (wrong physics;
// tools/variant_bench.py compares it with the product build to show how far the sweeps are from the memory system's floor).
#ifndef DFSPH_GATHER_ONLY
#define DFSPH_GATHER_ONLY 0
#endif
#ifndef DFSPH_IDX_LOAD
#define DFSPH_IDX_LOAD 2
#endif
__device__ __forceinline__ unsigned ld_index(const unsigned* p)
{
#if DFSPH_IDX_LOAD == 1
    return __ldcs(p);
#elif DFSPH_IDX_LOAD == 2
    unsigned v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

// Warp-uniform neighbour sweep: U gathers in flight per thread, table indices of the next batch prefetched while the
// current batch is processed.  F provides  Data load(unsigned j)  and  void apply(const Data&).
template <int U, class F>
__device__ __forceinline__ void neighbor_sweep(const TabRows t, unsigned count, F& f)
{
    if (count == 0) return;
    unsigned jn[U];
    unsigned o = t.off;
#pragma unroll
    for (int u = 0; u < U; ++u) jn[u] = ld_index(t.base + o + u * DFSPH_TILE);
    for (unsigned k = 0; k < count; k += U) {
        typename F::Data d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) d[u] = f.load(jn[u], u);
        o += U * DFSPH_TILE;
        if (k + U < count) {
#pragma unroll
            for (int u = 0; u < U; ++u) jn[u] = ld_index(t.base + o + u * DFSPH_TILE);
        }
        // pair geometry + kernel evaluation for the whole batch first (in lookup-table mode this issues the table
        // reads of all U pairs before any of them is consumed), then the accumulation
#pragma unroll
        for (int u = 0; u < U; ++u) f.prep(d[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) f.apply(d[u]);
    }
}

// ---- block reduction helpers -----------------------------------------------------------------------------------
#ifndef DFSPH_JACOBI_BLOCK
#define DFSPH_JACOBI_BLOCK 512   /* pass B: measured 6 % faster than 256 at 10 M particles (1024: 25 % slower) */
#endif

__device__ __forceinline__ double block_sum_double(double v)
{
    __shared__ double ws[32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) ws[w] = v;
    __syncthreads();
    double t = 0.0;
    if (w == 0) {
        t = lane < (int)(blockDim.x >> 5) ? ws[lane] : 0.0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
    }
    return t;   // valid in thread 0
}

// true in every thread of the block that arrives last at the ticket (deterministic final reduction happens there)
__device__ __forceinline__ bool last_block(Ctrl* ctrl)
{
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&ctrl->ticket, 1u);
        is_last = (t == gridDim.x - 1);
        if (is_last) ctrl->ticket = 0;
    }
    __syncthreads();
    return is_last;
}

// ---- K1+K2+K3 ------------------------------------------------------------------------------------------------------
#ifndef DFSPH_U2
#define DFSPH_U2 4   /* gathers in flight per thread, two-array sweeps */
#endif
#ifndef DFSPH_U1
#define DFSPH_U1 4   /* one-array sweeps */
#endif

template <int MODE>
struct InitFluidF {
    struct Data { Real4 x, v; Real rx, ry, rz, g, W; };
    const Real4* pos; const Real4* vel; const SphConst& c;
    cudaTextureObject_t pos_tex, vel_tex;
    Real4 xi, vi;
    Real dens, gx, gy, gz, sum_grad2, dadv;
    __device__ __forceinline__ InitFluidF(const FluidArrays& f, const SphConst& c_, Real4 xi_, Real4 vi_)
        : pos(f.pos), vel(f.vel), c(c_), pos_tex(f.pos_tex), vel_tex(f.vel_tex), xi(xi_), vi(vi_), dens(0), gx(0), gy(0), gz(0), sum_grad2(0), dadv(0) {}
    __device__ __forceinline__ Data load(unsigned j, int slot) const
    {
        Data d;
        d.x = gather4<DFSPH_TEX_POS_V != 0, true>(pos, pos_tex, j);
        d.v = gather4<DFSPH_TEX_VEL != 0, false>(vel, vel_tex, j);
        return d;
    }
    __device__ __forceinline__ void prep(Data& d) const
    {
        d.rx = xi.x - d.x.x; d.ry = xi.y - d.x.y; d.rz = xi.z - d.x.z;
        sph_W_gradW<MODE>(c, d.rx * d.rx + d.ry * d.ry + d.rz * d.rz, d.W, d.g);
    }
    __device__ __forceinline__ void apply(const Data& d)
    {
        const Real rx = d.rx, ry = d.ry, rz = d.rz, W = d.W, g = d.g;
        const Real V = c.V;
        dens += V * W;
#if DFSPH_REAL_IS_DOUBLE
        // scalar variant: grad_p_j = -V gradW; sum += |grad_p_j|^2; grad_p_i -= grad_p_j; dadv sums without V
        const Real px = V * (g * rx), py = V * (g * ry), pz = V * (g * rz);
        sum_grad2 += px * px + py * py + pz * pz;
        gx += px; gy += py; gz += pz;
        dadv += (vi.x - d.v.x) * (g * rx) + (vi.y - d.v.y) * (g * ry) + (vi.z - d.v.z) * (g * rz);
#else
        // AVX variant: V_gradW = gradW * V_j
        const Real gv = g * V;
        const Real px = rx * gv, py = ry * gv, pz = rz * gv;
        sum_grad2 += px * px + py * py + pz * pz;
        gx += px; gy += py; gz += pz;
        dadv += (vi.x - d.v.x) * px + (vi.y - d.v.y) * py + (vi.z - d.v.z) * pz;
#endif
    }
};

template <int MODE>
struct InitBoundaryF {
    struct Data { Real4 x; Real rx, ry, rz, g, W; };
    const Real4* bpos; const SphConst& c;
    Real4 xi, vi;
    Real dens, bx, by, bz, dadv;
    __device__ __forceinline__ InitBoundaryF(const Real4* bpos_, const SphConst& c_, Real4 xi_, Real4 vi_)
        : bpos(bpos_), c(c_), xi(xi_), vi(vi_), dens(0), bx(0), by(0), bz(0), dadv(0) {}
    __device__ __forceinline__ Data load(unsigned j, int slot) const { Data d; d.x = ld_gather(bpos + j); return d; }
    __device__ __forceinline__ void prep(Data& d) const
    {
        d.rx = xi.x - d.x.x; d.ry = xi.y - d.x.y; d.rz = xi.z - d.x.z;
        sph_W_gradW<MODE>(c, d.rx * d.rx + d.ry * d.ry + d.rz * d.rz, d.W, d.g);
    }
    __device__ __forceinline__ void apply(const Data& d)   // d.x.w = V_b (0 for the sentinel)
    {
        const Real rx = d.rx, ry = d.ry, rz = d.rz, W = d.W, g = d.g;
        dens += d.x.w * W;
        const Real gv = d.x.w * g;
        const Real px = gv * rx, py = gv * ry, pz = gv * rz;
        bx += px; by += py; bz += pz;
        dadv += vi.x * px + vi.y * py + vi.z * pz;   // static boundary: v_b = 0
    }
};

template <int MODE, bool DIV_SOLVER>
__global__ void __launch_bounds__(DFSPH_BLOCK) k_init_sweep(FluidArrays f, SphConst c, const Real4* __restrict__ bpos, Ctrl* ctrl)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f.n || ctrl->fatal) return;
    const Real4 xi = ld_plain(f.pos + i);   // plain loads: this kernel writes pos.w
    const Real4 vi = ld_gather(f.vel + i);
    const Real V = c.V;

    InitFluidF<MODE> ff(f, c, xi, vi);
    neighbor_sweep<DFSPH_U2>(tab_rows(f.tab_f, f.Kf, i), f.tcnt_f[i >> 5], ff);
    Real dadv = ff.dadv;
#if DFSPH_REAL_IS_DOUBLE
    dadv *= V;   // "assumes that all fluid particles have the same volume" (TimeStepDFSPH.cpp:1266-1267)
#endif

    // Akinci2012 boundary neighbours (static bodies: v_b = 0)
    InitBoundaryF<MODE> fb(bpos, c, xi, vi);
    neighbor_sweep<DFSPH_U1>(tab_rows(f.tab_b, f.Kb, i), f.tcnt_b[i >> 5], fb);
    dadv += fb.dadv;

    // density (TimeStep.cpp:70,110 / 133,166)
    const Real density = (V * c.W_zero + (ff.dens + fb.dens)) * c.density0;
    f.density[i] = density;
    f.vel[i].w = density;   // (v, rho) in one record: the viscosity sweep gathers both with one load
    st_real4(f.bgrad + i, make_real4(fb.bx, fb.by, fb.bz, (Real)0.0));

    // factor (TimeStepDFSPH.cpp:812-821 / 1170-1182)
    const Real gx = ff.gx + fb.bx, gy = ff.gy + fb.by, gz = ff.gz + fb.bz;
    const Real sum_grad2 = ff.sum_grad2 + (gx * gx + gy * gy + gz * gz);
    Real factor = sum_grad2 > DFSPH_EPS ? (Real)1.0 / sum_grad2 : (Real)0.0;

    const unsigned nn = f.cnt_f[i] + f.cnt_b[i];
    f.nnbr[i] = nn;

    if (DIV_SOLVER) {
        // divergence-solve init loop (TimeStepDFSPH.cpp:410-461)
        const Real h = ctrl->h;
        const Real invH = (Real)1.0 / h;
        f.density_adv[i] = dadv;
        Real dp = real_max(dadv, (Real)0.0);
        if (nn < 20u) dp = (Real)0.0;
        factor *= invH;
        const Real kv_old = f.kappa_v[i];
        const Real kv = dp > (Real)0.0 ? (Real)0.5 * real_min(kv_old, (Real)0.5) * invH : (Real)0.0;
        f.pos[i].w = kv;
    }
    f.factor[i] = factor;
}

// ---- pressure acceleration of particle i from the kappa values in pos.w -------------------------------------------
template <int MODE>
struct AccelF {
    struct Data { Real4 x; Real rx, ry, rz, g; };
    const Real4* pos; const SphConst& c;
    cudaTextureObject_t pos_tex;
    Real4 xi;
    Real ax, ay, az;
    __device__ __forceinline__ AccelF(const FluidArrays& f, const SphConst& c_, Real4 xi_) : pos(f.pos), c(c_), pos_tex(f.pos_tex), xi(xi_), ax(0), ay(0), az(0) {}
    __device__ __forceinline__ Data load(unsigned j, int slot) const
    {
        Data d;
        d.x = gather4<DFSPH_TEX_POS_A != 0, false>(pos, pos_tex, j);
        return d;
    }
    __device__ __forceinline__ void prep(Data& d) const
    {
#if DFSPH_GATHER_ONLY
        return;
#endif
        d.rx = xi.x - d.x.x; d.ry = xi.y - d.x.y; d.rz = xi.z - d.x.z;
#if DFSPH_REAL_IS_DOUBLE
        d.g = sph_gradW_scale<MODE>(c, d.rx * d.rx + d.ry * d.ry + d.rz * d.rz);
#else
        d.g = sph_V_gradW_scale<MODE>(c, d.rx * d.rx + d.ry * d.ry + d.rz * d.rz);   // V gradW scale
#endif
    }
    __device__ __forceinline__ void apply(const Data& d)
    {
#if DFSPH_GATHER_ONLY
        const Real tiny = (Real)1.0e-30;   // measurement build: the memory system's share of the sweep (the pressure force vanishes)
        ax += tiny * d.x.x; ay += tiny * d.x.y; az += tiny * (d.x.z + d.x.w);
        return;
#endif
        const Real rx = d.rx, ry = d.ry, rz = d.rz, g = d.g;
        const Real pSum = xi.w + d.x.w;   // density0 ratio is 1 (single phase)
#if DFSPH_REAL_IS_DOUBLE
        if (real_abs(pSum) > DFSPH_EPS) {   // scalar variant skips tiny sums (TimeStepDFSPH.cpp:1323-1327)
            const Real s = -c.V * g * pSum;
            ax += s * rx; ay += s * ry; az += s * rz;
        }
#else
        const Real s = g * pSum;            // delta_ai -= V_gradW * pSum (TimeStepDFSPH.cpp:987); g = V gradW scale
        ax -= s * rx; ay -= s * ry; az -= s * rz;
#endif
    }
};

template <int MODE>
__device__ __forceinline__ void pressure_accel(const FluidArrays& f, const SphConst& c, unsigned i, const Real4 xi, Real& ax, Real& ay, Real& az, bool listed = false)
{
    AccelF<MODE> fa(f, c, xi);
    neighbor_sweep<DFSPH_U1>(listed ? tab_rows_any(f.tab_f, f.Kf, i) : tab_rows(f.tab_f, f.Kf, i), f.tcnt_f[i >> 5], fa);
    ax = fa.ax; ay = fa.ay; az = fa.az;
    const Real ki = xi.w;
    if (real_abs(ki) > DFSPH_EPS) {          // boundary term (:993-1010 / 1333-1345): a_i -= kappa_i * G_i
        const Real4 b = ld_gather(f.bgrad + i);
        ax -= ki * b.x; ay -= ki * b.y; az -= ki * b.z;
    }
}

// Non-active particles (emitter-animated / fixed) keep a zero pressure acceleration; their lanes still walk the
// warp-uniform loop (results discarded) so that the sweep stays convergent.
// Multi-GPU overlap: the particles a neighbour rank needs (export list) are processed first by a launch over `list`,
// their values travel on the communication stream while a second launch over all particles skips them (`skip`).
// Single GPU: list == nullptr, skip == nullptr.
#ifndef DFSPH_ACCEL_MIN_BLOCKS
#if DFSPH_REAL_IS_DOUBLE
#define DFSPH_ACCEL_MIN_BLOCKS (MODE == KM_LUT ? 4 : 1)   /* double, table kernel: 64-register cap (unbounded the compiler takes 92 and halves the occupancy); the analytic kernels need ~80-116 registers */
#else
#define DFSPH_ACCEL_MIN_BLOCKS 8    /* 32 registers, full occupancy: -3 % */
#endif
#endif
template <int MODE>
__global__ void __launch_bounds__(DFSPH_BLOCK, DFSPH_ACCEL_MIN_BLOCKS) k_accel(FluidArrays f, SphConst c, const Ctrl* __restrict__ ctrl,
                                                         const unsigned* __restrict__ list, unsigned list_n, const unsigned char* __restrict__ skip,
                                                         GhostWait gw)
{
    if (ctrl->done) return;
    ghost_wait(gw);     // ghost kappa pushed by the neighbour ranks (no-op on a single GPU)
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (list) { if (i >= list_n) return; i = list[i]; }
    else { if (i >= (f.n_dev ? *f.n_dev : f.n)) return; if (skip && skip[i]) return; }
    Real ax, ay, az;
    const Real4 xi = ld_gather(f.pos + i);
    pressure_accel<MODE>(f, c, i, xi, ax, ay, az, list != nullptr);
    if (f.state[i] != 0u) { ax = ay = az = (Real)0.0; }
    st_real4(f.acc + i, make_real4(ax, ay, az, (Real)0.0));
}

// ---- pass B --------------------------------------------------------------------------------------------------------
enum { SOLVE_DIV = 0, SOLVE_PRESS = 1 };

template <int MODE>
struct JacobiF {
    struct Data { Real4 x, a; Real rx, ry, rz, g; };
    const Real4* pos; const Real4* acc; const SphConst& c;
    cudaTextureObject_t acc_tex, pos_tex;
#if DFSPH_GATHER_ONLY == 2
    const unsigned* rec32;
#endif
    Real4 xi, ai;
    Real sum;
    __device__ __forceinline__ JacobiF(const FluidArrays& f, const SphConst& c_, Real4 xi_, Real4 ai_) : pos(f.pos), acc(f.acc), c(c_), acc_tex(f.acc_tex), pos_tex(f.pos_tex),
#if DFSPH_GATHER_ONLY == 2
        rec32(f.tab_b),
#endif
        xi(xi_), ai(ai_), sum(0) {}
    __device__ __forceinline__ Data load(unsigned j, int slot) const
    {
        // The two scattered gathers of pass B: the L1 data pipe is the limiter of this kernel, and a gather through the
        // texture unit costs fewer wavefronts than LDG.128 (tools/micro/tex_bench.cu).
        Data d;
#if DFSPH_GATHER_ONLY == 2 && !DFSPH_REAL_IS_DOUBLE
        // measurement build: ONE 32-byte gather per pair from a combined-record array (stand-in: the boundary table's memory,
        // contents irrelevant) instead of x_j through the LSU + a_j through the texture unit
        double a0, a1, a2, a3;
        asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a0), "=d"(a1), "=d"(a2), "=d"(a3) : "l"(reinterpret_cast<const char*>(rec32) + 32ull * j));
        const unsigned m = 0x3fu;
        d.x = make_real4((Real)(__double2loint(a0) & m), (Real)(__double2hiint(a0) & m), (Real)(__double2loint(a1) & m), (Real)0.0);
        d.a = make_real4((Real)(__double2loint(a2) & m), (Real)(__double2hiint(a2) & m), (Real)(__double2loint(a3) & m), (Real)0.0);
        return d;
#endif
        d.x = gather4<DFSPH_TEX_POS_B != 0, true>(pos, pos_tex, j);
        d.a = gather4<DFSPH_TEX_ACC_B != 0, false>(acc, acc_tex, j);
        return d;
    }
    __device__ __forceinline__ void prep(Data& d) const
    {
#if DFSPH_GATHER_ONLY
        return;
#endif
        d.rx = xi.x - d.x.x; d.ry = xi.y - d.x.y; d.rz = xi.z - d.x.z;
#if DFSPH_REAL_IS_DOUBLE
        d.g = sph_gradW_scale<MODE>(c, d.rx * d.rx + d.ry * d.ry + d.rz * d.rz);
#else
        d.g = sph_V_gradW_scale<MODE>(c, d.rx * d.rx + d.ry * d.ry + d.rz * d.rz);   // V gradW scale
#endif
    }
    __device__ __forceinline__ void apply(const Data& d)
    {
#if DFSPH_GATHER_ONLY
        sum += (Real)1.0e-30 * ((d.x.x + d.x.y) + (d.x.z + d.a.x) + (d.a.y + d.a.z));   // measurement build
        return;
#endif
        const Real rx = d.rx, ry = d.ry, rz = d.rz, g = d.g;
#if DFSPH_REAL_IS_DOUBLE
        sum += (ai.x - d.a.x) * (g * rx) + (ai.y - d.a.y) * (g * ry) + (ai.z - d.a.z) * (g * rz);
#else
        const Real s = g;   // V gradW scale
        sum += (ai.x - d.a.x) * (rx * s) + (ai.y - d.a.y) * (ry * s) + (ai.z - d.a.z) * (rz * s);
#endif
    }
};

// loop control (TimeStepDFSPH.cpp:324-341 / 477-495) on the reduced density error; one thread
template <int SOLVE>
__device__ __forceinline__ void solve_control(Ctrl* ctrl, const SolverParams& sp, const SphConst& c, unsigned long long n)
{
    const Real density_error = (Real)ctrl->err_sum;
    const Real avg = n > 0ull ? density_error / (Real)n : (Real)0.0;   // empty model: iteration is a no-op (:550-551)
    Real eta;
    unsigned min_it, max_it;
    if (SOLVE == SOLVE_PRESS) { eta = sp.max_error * (Real)0.01 * c.density0; min_it = sp.min_iter; max_it = sp.max_iter; }
    else { eta = ((Real)1.0 / ctrl->h) * sp.max_error_v * (Real)0.01 * c.density0; min_it = 1u; max_it = sp.max_iter_v; }
    const bool chk = avg <= eta;
    const unsigned it = ctrl->iter + 1u;
    ctrl->iter = it;
    if (SOLVE == SOLVE_PRESS) { ctrl->avg_err = (double)avg; ctrl->iterations = it; }
    else { ctrl->avg_err_v = (double)avg; ctrl->iterations_v = it; }
    const bool cont = (!chk || (it < min_it)) && (it < max_it);
    ctrl->done = cont ? 0 : 1;
}

// list / skip: see k_accel.  partial_base: first slot of this launch in `partial`; finalize: this launch elects the last
// block, which sums partial[0 .. partial_base + gridDim.x) (the export-list launch runs first with finalize = 0).
#ifndef DFSPH_JACOBI_MIN_BLOCKS
#if DFSPH_REAL_IS_DOUBLE
#define DFSPH_JACOBI_MIN_BLOCKS (MODE == KM_LUT ? 2 : 1)   /* double, table kernel: 64 registers, 2 x 512 threads per SM (1 block: 104 registers, 35 % slower; 3: spills) */
#else
#define DFSPH_JACOBI_MIN_BLOCKS 3   /* 40 registers, 75 % occupancy: 0.97 -> 0.89 ms at 10 M (4 blocks spill: 1.27 ms) */
#endif
#endif
template <int MODE, int SOLVE>
__global__ void __launch_bounds__(DFSPH_JACOBI_BLOCK, DFSPH_JACOBI_MIN_BLOCKS) k_jacobi(FluidArrays f, SphConst c, SolverParams sp, Ctrl* ctrl, double* __restrict__ partial,
                                                          const unsigned* __restrict__ list, unsigned list_n, const unsigned char* __restrict__ skip,
                                                          unsigned partial_base, int finalize, GhostWait gw, PeerReduce pr)
{
    if (ctrl->done) return;
    ghost_wait(gw);     // ghost pressure accelerations pushed by the neighbour ranks
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    bool active;
    if (list) { active = i < list_n; if (active) i = list[i]; }
    else active = i < (f.n_dev ? *f.n_dev : f.n) && !(skip && skip[i]);
    double err = 0.0;
    if (active) {
        const Real h = ctrl->h;
        const Real4 xi = ld_plain(f.pos + i);
        const Real4 ai = ld_gather(f.acc + i);
        JacobiF<MODE> fj(f, c, xi, ai);
        neighbor_sweep<DFSPH_U2>(list ? tab_rows_any(f.tab_f, f.Kf, i) : tab_rows(f.tab_f, f.Kf, i), f.tcnt_f[i >> 5], fj);
        Real sum = fj.sum;
#if DFSPH_REAL_IS_DOUBLE
        sum *= c.V;
#endif
        const Real4 b = ld_gather(f.bgrad + i);
        sum += ai.x * b.x + ai.y * b.y + ai.z * b.z;

        if (SOLVE == SOLVE_DIV || f.state[i] == 0u) {
            Real aij_pj = sum;
            Real s_i;
            if (SOLVE == SOLVE_PRESS) { aij_pj *= h * h; s_i = (Real)1.0 - f.density_adv[i]; }
            else { aij_pj *= h; s_i = -f.density_adv[i]; }
            Real residuum = real_min(s_i - aij_pj, (Real)0.0);
            if (SOLVE == SOLVE_DIV && f.nnbr[i] < 20u) residuum = (Real)0.0;
            const Real knew = real_max(xi.w - (Real)0.5 * (s_i - aij_pj) * f.factor[i], (Real)0.0);
            f.pos[i].w = knew;
            err = -(double)(c.density0 * residuum);
        }
    }
    // density-error reduction: block partials, summed in fixed order by the last block (deterministic)
    const double bsum = block_sum_double(err);
    if (threadIdx.x == 0) partial[partial_base + blockIdx.x] = bsum;
    if (!finalize) return;
    if (last_block(ctrl)) {
        double t = 0.0;
        for (unsigned b = threadIdx.x; b < partial_base + gridDim.x; b += blockDim.x) t += partial[b];
        t = block_sum_double(t);
        if (pr.world > 0) {
            // fused all-reduce over peer memory: publish my partial sum in every rank's table, wait for everybody's,
            // add them up in rank order (bitwise identical on all ranks), then take the loop decision right here
            __shared__ double my_sum;
            __shared__ unsigned red_seq;
            if (threadIdx.x == 0) { my_sum = t; red_seq = ctrl->red_seq + 1u; }
            __syncthreads();
            const unsigned s = red_seq, par = s & 1u;
            if ((int)threadIdx.x < pr.world) {
                const int r = threadIdx.x;
                pr.val[r][par * DFSPH_MAX_RANKS + pr.rank] = my_sum;
                __threadfence_system();
                asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(pr.seq[r] + pr.rank), "r"(s) : "memory");
                unsigned v;
                do { asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(pr.seq[pr.rank] + r) : "memory"); } while ((int)(v - s) < 0);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence_system();
                double g = 0.0;
                const volatile double* mine = pr.val[pr.rank] + par * DFSPH_MAX_RANKS;
                for (int r = 0; r < pr.world; ++r) g += mine[r];
                ctrl->err_sum = g;
                ctrl->red_seq = s;
                solve_control<SOLVE>(ctrl, sp, c, ctrl->n_global);
            }
        } else if (threadIdx.x == 0) {
            ctrl->err_sum = t;
            if (!ctrl->multi) solve_control<SOLVE>(ctrl, sp, c, (unsigned long long)f.n);
        }
    }
}

// multi-GPU: runs after ncclAllReduce(sum) of ctrl->err_sum; every rank takes the identical decision
template <int SOLVE>
__global__ void k_solve_control(Ctrl* ctrl, SolverParams sp, SphConst c)
{
    if (ctrl->done) return;
    solve_control<SOLVE>(ctrl, sp, c, ctrl->n_global);
}

// Start of a solve: reset the loop state.  For n == 0 the reference's iteration returns immediately with avg = 0.
__global__ void k_solve_begin(Ctrl* ctrl, int solve)
{
    ctrl->iter = 0;
    ctrl->done = ctrl->fatal ? 1 : 0;   // a failed step runs no iteration
    ctrl->ticket = 0;
    if (solve == SOLVE_PRESS) { ctrl->iterations = 0; ctrl->avg_err = 0.0; }
    else { ctrl->iterations_v = 0; ctrl->avg_err_v = 0.0; }
}

// ---- divergence finaliser + non-pressure kick + CFL -----------------------------------------------------------------
// KICK = false: only the divergence finaliser; the non-pressure kick and the CFL scan happen in k_viscosity_kick.
template <int MODE, bool DIV_SOLVER, bool KICK>
__global__ void __launch_bounds__(DFSPH_BLOCK) k_div_final(FluidArrays f, SphConst c, SolverParams sp, Ctrl* ctrl)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const Real h = ctrl->h;    // still the step's initial h here
    Real velmag = (Real)0.0;
    if (ctrl->fatal) return;
    if (i < f.n) {
        Real4 v = ld_gather(f.vel + i);
        const unsigned st = f.state[i];
        if (DIV_SOLVER) {
            const Real4 xi = ld_gather(f.pos + i);
            Real ax, ay, az;
            pressure_accel<MODE>(f, c, i, xi, ax, ay, az);
            if (st != 0u) { ax = ay = az = (Real)0.0; }
            st_real4(f.acc + i, make_real4(ax, ay, az, (Real)0.0));
            v.x += h * ax; v.y += h * ay; v.z += h * az;                 // TimeStepDFSPH.cpp:516
            f.factor[i] *= h;                                            // :518
            f.kappa_v[i] = xi.w * h;                                     // :536
        }
        if (KICK) {
            // clearAccelerations: a = g (TimeStep.cpp:41-48); CFL term |v + a h|^2 (Simulation.cpp:431-439)
            const Real tx = v.x + sp.gx * h, ty = v.y + sp.gy * h, tz = v.z + sp.gz * h;
            velmag = tx * tx + ty * ty + tz * tz;
            if (st == 0u) { v.x += h * sp.gx; v.y += h * sp.gy; v.z += h * sp.gz; }   // TimeStepDFSPH.cpp:192-208
        }
        st_real4(f.vel + i, v);
    }
    if (KICK) {
        // block max -> global max (non-negative floats order like their bit patterns)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) velmag = real_max(velmag, __shfl_xor_sync(0xffffffffu, velmag, d));
        __shared__ Real wm[DFSPH_BLOCK / 32];
        if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = velmag;
        __syncthreads();
        if (threadIdx.x == 0) {
            Real m = wm[0];
            for (int w = 1; w < DFSPH_BLOCK / 32; ++w) m = real_max(m, wm[w]);
#if DFSPH_REAL_IS_DOUBLE
            atomicMax(&ctrl->maxvel_bits, (unsigned long long)__double_as_longlong(m));
#else
            atomicMax(&ctrl->maxvel_bits, (unsigned long long)__float_as_uint(m));
#endif
        }
    }
}

// ---- Viscosity_Standard (next-row f1; Viscosity/Viscosity_Standard.cpp:48-238 AVX, :242-400 scalar) + kick + CFL -----------
// a_i = g + d mu sum_j (m_j/rho_j) (v_ij . x_ij)/(|x_ij|^2 + 0.01 h^2) gradW_ij  [+ boundary term if mu_b != 0], d = 10;
// then the CFL term and v += h a exactly as in k_div_final.  Neighbours read the pre-kick velocities, so the kicked
// velocity goes to vel_out.
template <int MODE>
struct ViscosityF {
    struct Data { Real4 x, v; Real rx, ry, rz, g; };
    const Real4* pos; const Real4* vel; const SphConst& c;
    cudaTextureObject_t pos_tex, vel_tex;
    Real4 xi, vi;
    Real ax, ay, az, eps2, dvisc;
    __device__ __forceinline__ ViscosityF(const FluidArrays& f, const SphConst& c_, Real4 xi_, Real4 vi_, Real dvisc_)
        : pos(f.pos), vel(f.vel), c(c_), pos_tex(f.pos_tex), vel_tex(f.vel_tex), xi(xi_), vi(vi_), ax(0), ay(0), az(0), eps2((Real)0.01 * c_.R * c_.R), dvisc(dvisc_) {}
    __device__ __forceinline__ Data load(unsigned j, int slot) const
    {
        Data d;
        d.x = gather4<DFSPH_TEX_POS_V != 0, false>(pos, pos_tex, j);
        d.v = gather4<DFSPH_TEX_VEL != 0, false>(vel, vel_tex, j);
        return d;
    }
    __device__ __forceinline__ void prep(Data& d) const
    {
        d.rx = xi.x - d.x.x; d.ry = xi.y - d.x.y; d.rz = xi.z - d.x.z;
        d.g = sph_gradW_scale<MODE>(c, d.rx * d.rx + d.ry * d.ry + d.rz * d.rz);
    }
    __device__ __forceinline__ void apply(const Data& d)
    {
        const Real r2 = d.rx * d.rx + d.ry * d.ry + d.rz * d.rz;
        const Real vx = (vi.x - d.v.x) * d.rx + (vi.y - d.v.y) * d.ry + (vi.z - d.v.z) * d.rz;
        // d.v.w = rho_j (1 for the sentinel: its gradient is 0 anyway); dvisc = d * mu * V * rho0 = d * mu * m_j
        const Real rho_j = d.v.w > (Real)0.0 ? d.v.w : (Real)1.0;
        const Real s = d.g * ((dvisc / rho_j) * vx / (r2 + eps2));
        ax += s * d.rx; ay += s * d.ry; az += s * d.rz;
    }
};

template <int MODE>
struct ViscosityBoundaryF {
    struct Data { Real4 x; Real rx, ry, rz, g; };
    const Real4* bpos; const SphConst& c;
    Real4 xi, vi;
    Real ax, ay, az, eps2, coef;
    __device__ __forceinline__ ViscosityBoundaryF(const Real4* bpos_, const SphConst& c_, Real4 xi_, Real4 vi_, Real coef_)
        : bpos(bpos_), c(c_), xi(xi_), vi(vi_), ax(0), ay(0), az(0), eps2((Real)0.01 * c_.R * c_.R), coef(coef_) {}
    __device__ __forceinline__ Data load(unsigned j, int slot) const { Data d; d.x = ld_gather(bpos + j); return d; }
    __device__ __forceinline__ void prep(Data& d) const
    {
        d.rx = xi.x - d.x.x; d.ry = xi.y - d.x.y; d.rz = xi.z - d.x.z;
        d.g = sph_gradW_scale<MODE>(c, d.rx * d.rx + d.ry * d.ry + d.rz * d.rz);
    }
    __device__ __forceinline__ void apply(const Data& d)   // static boundary: v_b = 0; d.x.w = V_b
    {
        const Real r2 = d.rx * d.rx + d.ry * d.ry + d.rz * d.rz;
        const Real vx = vi.x * d.rx + vi.y * d.ry + vi.z * d.rz;
        const Real s = d.g * (coef * d.x.w * vx / (r2 + eps2));
        ax += s * d.rx; ay += s * d.ry; az += s * d.rz;
    }
};

template <int MODE>
__global__ void __launch_bounds__(DFSPH_BLOCK) k_viscosity_kick(FluidArrays f, SphConst c, SolverParams sp, Ctrl* ctrl,
                                                                  const Real4* __restrict__ bpos, Real4* __restrict__ vel_out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const Real h = ctrl->h;
    Real velmag = (Real)0.0;
    if (ctrl->fatal) { if (i < f.n) st_real4(vel_out + i, ld_gather(f.vel + i)); return; }
    if (i < f.n) {
        const Real4 xi = ld_gather(f.pos + i);
        Real4 v = ld_gather(f.vel + i);
        const unsigned st = f.state[i];
        ViscosityF<MODE> fv(f, c, xi, v, (Real)10.0 * sp.viscosity * c.V * c.density0);
        neighbor_sweep<DFSPH_U2>(tab_rows(f.tab_f, f.Kf, i), f.tcnt_f[i >> 5], fv);
        Real ax = sp.gx + fv.ax, ay = sp.gy + fv.ay, az = sp.gz + fv.az;
        if (sp.viscosity_boundary != (Real)0.0) {
            ViscosityBoundaryF<MODE> fb(bpos, c, xi, v, (Real)10.0 * sp.viscosity_boundary * c.density0 / f.density[i]);
            neighbor_sweep<DFSPH_U1>(tab_rows(f.tab_b, f.Kb, i), f.tcnt_b[i >> 5], fb);
            ax += fb.ax; ay += fb.ay; az += fb.az;
        }
        const Real tx = v.x + ax * h, ty = v.y + ay * h, tz = v.z + az * h;     // Simulation.cpp:431-439
        velmag = tx * tx + ty * ty + tz * tz;
        if (st == 0u) { v.x += h * ax; v.y += h * ay; v.z += h * az; }           // TimeStepDFSPH.cpp:192-208
        st_real4(vel_out + i, v);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) velmag = real_max(velmag, __shfl_xor_sync(0xffffffffu, velmag, d));
    __shared__ Real wm[DFSPH_BLOCK / 32];
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = velmag;
    __syncthreads();
    if (threadIdx.x == 0) {
        Real m = wm[0];
        for (int w = 1; w < DFSPH_BLOCK / 32; ++w) m = real_max(m, wm[w]);
#if DFSPH_REAL_IS_DOUBLE
        atomicMax(&ctrl->maxvel_bits, (unsigned long long)__double_as_longlong(m));
#else
        atomicMax(&ctrl->maxvel_bits, (unsigned long long)__float_as_uint(m));
#endif
    }
}

// Simulation::updateTimeStepSize (Simulation.cpp:395-413, 415-493) on the reduced maximum; one thread.
__global__ void k_update_time_step(Ctrl* ctrl, SolverParams sp)
{
    Real h = ctrl->h;
    if (ctrl->fatal) return;
    ctrl->h_step = h;
    if (sp.cfl_method == 1 || sp.cfl_method == 2) {
#if DFSPH_REAL_IS_DOUBLE
        Real maxVel = __longlong_as_double((long long)ctrl->maxvel_bits);
#else
        Real maxVel = __uint_as_float((unsigned)ctrl->maxvel_bits);
#endif
        const Real diameter = (Real)2.0 * sp.radius;
        if (maxVel < (Real)1.0e-9) maxVel = (Real)1.0e-9;
        Real hn = sp.cfl_factor * (Real)0.4 * (diameter / real_sqrt(maxVel));
        hn = real_min(hn, sp.cfl_max);
        hn = real_max(hn, sp.cfl_min);
        if (sp.cfl_method == 2) {
            // iteration-count based adaption (Simulation.cpp:399-412); iterations of the previous pressure solve
            const unsigned iterations = ctrl->iterations;
            if (iterations != 0u) {
                Real h2 = h;
                if (iterations > 10u) h2 *= (Real)0.9;
                else if (iterations < 5u) h2 *= (Real)1.1;
                hn = real_min(h2, hn);
            }
        }
        h = hn;
    }
    ctrl->h = h;
    ctrl->maxvel_bits = 0ull;
}

// ---- pressure solve init -------------------------------------------------------------------------------------------
template <int MODE>
struct VelDivF {
    struct Data { Real4 x, v; Real rx, ry, rz, g; };
    const Real4* pos; const Real4* vel; const SphConst& c;
    cudaTextureObject_t pos_tex, vel_tex;
    Real4 xi, vi;
    Real delta;
    __device__ __forceinline__ VelDivF(const FluidArrays& f, const SphConst& c_, Real4 xi_, Real4 vi_) : pos(f.pos), vel(f.vel), c(c_), pos_tex(f.pos_tex), vel_tex(f.vel_tex), xi(xi_), vi(vi_), delta(0) {}
    __device__ __forceinline__ Data load(unsigned j, int slot) const
    {
        Data d;
        d.x = gather4<DFSPH_TEX_POS_V != 0, true>(pos, pos_tex, j);
        d.v = gather4<DFSPH_TEX_VEL != 0, false>(vel, vel_tex, j);
        return d;
    }
    __device__ __forceinline__ void prep(Data& d) const
    {
        d.rx = xi.x - d.x.x; d.ry = xi.y - d.x.y; d.rz = xi.z - d.x.z;
#if DFSPH_REAL_IS_DOUBLE
        d.g = sph_gradW_scale<MODE>(c, d.rx * d.rx + d.ry * d.ry + d.rz * d.rz);
#else
        d.g = sph_V_gradW_scale<MODE>(c, d.rx * d.rx + d.ry * d.ry + d.rz * d.rz);   // V gradW scale
#endif
    }
    __device__ __forceinline__ void apply(const Data& d)
    {
        const Real rx = d.rx, ry = d.ry, rz = d.rz, g = d.g;
#if DFSPH_REAL_IS_DOUBLE
        delta += (vi.x - d.v.x) * (g * rx) + (vi.y - d.v.y) * (g * ry) + (vi.z - d.v.z) * (g * rz);
#else
        const Real s = g;   // V gradW scale
        delta += (vi.x - d.v.x) * (rx * s) + (vi.y - d.v.y) * (ry * s) + (vi.z - d.v.z) * (rz * s);
#endif
    }
};

template <int MODE>
__global__ void __launch_bounds__(DFSPH_BLOCK) k_press_init(FluidArrays f, SphConst c, Ctrl* ctrl)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f.n || ctrl->fatal) return;
    const Real h = ctrl->h;     // the NEW time step size (TimeStepDFSPH.cpp:254)
    const Real4 xi = ld_plain(f.pos + i);
    const Real4 vi = ld_gather(f.vel + i);
    VelDivF<MODE> fv(f, c, xi, vi);
    neighbor_sweep<DFSPH_U2>(tab_rows(f.tab_f, f.Kf, i), f.tcnt_f[i >> 5], fv);
    Real delta = fv.delta;
#if DFSPH_REAL_IS_DOUBLE
    delta *= c.V;
#endif
    const Real4 b = ld_gather(f.bgrad + i);
    delta += vi.x * b.x + vi.y * b.y + vi.z * b.z;

    const Real dadv = f.density[i] / c.density0 + h * delta;     // :888 / 1241
    f.density_adv[i] = dadv;
    const Real invH2 = (Real)1.0 / (h * h);
    f.factor[i] *= invH2;                                         // :283
    const Real kold = f.kappa[i];
    const Real k = dadv > (Real)1.0 ? (Real)0.5 * real_min(kold, (Real)0.00025) * invH2 : (Real)0.0;   // :291-294
    f.pos[i].w = k;
}

// ---- pressure finaliser + advection ----------------------------------------------------------------------------------
// Positions are written to pos_out (the other half of the double buffer): neighbours still read the old positions.
// out_x / out_v (step_host): the new positions and velocities additionally leave as packed 3-vectors in host row order
// (row = out_id[i], or i when out_id is null), ready for the device-to-host copies -- no separate pack kernels.
template <int MODE>
__global__ void __launch_bounds__(DFSPH_BLOCK) k_press_final(FluidArrays f, SphConst c, Ctrl* ctrl, Real4* __restrict__ pos_out,
                                                            Real* __restrict__ out_x, Real* __restrict__ out_v, const unsigned* __restrict__ out_id)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f.n) return;
    const Real h = ctrl->h;
    const Real hs = ctrl->h_step;
    const Real4 xi = ld_gather(f.pos + i);
    if (ctrl->fatal) { st_real4(pos_out + i, xi); return; }   // failed step: the state stays where the search left it
    const unsigned st = f.state[i];
    Real ax, ay, az;
    pressure_accel<MODE>(f, c, i, xi, ax, ay, az);
    if (st != 0u) { ax = ay = az = (Real)0.0; }
    st_real4(f.acc + i, make_real4(ax, ay, az, (Real)0.0));
    Real4 v = ld_gather(f.vel + i);
    v.x += h * ax; v.y += h * ay; v.z += h * az;                  // :360
    st_real4(f.vel + i, v);
    f.kappa[i] = xi.w * (h * h);                                  // :379
    Real4 xo = xi;
    if (st == 0u) { xo.x += hs * v.x; xo.y += hs * v.y; xo.z += hs * v.z; }   // :220-237 (h captured at step start)
    st_real4(pos_out + i, xo);
    if (out_x) {
        const size_t d = 3 * (size_t)(out_id ? out_id[i] : i);
        out_x[d] = xo.x; out_x[d + 1] = xo.y; out_x[d + 2] = xo.z;
        out_v[d] = v.x; out_v[d + 1] = v.y; out_v[d + 2] = v.z;
    }
}

__global__ void k_step_begin(Ctrl* ctrl)
{
    if (ctrl->fatal) return;
    ctrl->max_nbr = 0u;
    ctrl->overflow = 0u;
    ctrl->overflow_b = 0u;
}

__global__ void k_step_end(Ctrl* ctrl)
{
    if (ctrl->fatal) return;
    ctrl->time += (double)ctrl->h_step;      // TimeStepDFSPH.cpp:248
}

// After the table build: a list that did not fit makes the step -- and the context -- fail instead of running the solver on
// truncated neighbour sets.  The flag is sticky on the device, so no host round trip is needed to stop the step.
__global__ void k_check_capacity(Ctrl* ctrl, unsigned Kf, unsigned Kb)
{
    if (ctrl->overflow > Kf || ctrl->overflow_b > Kb) ctrl->fatal = 1u;
}

// Last node of the body of the solver loops' WHILE node (CUDA graph): keep iterating while the loop control says so.
__global__ void k_loop_cond(cudaGraphConditionalHandle handle, const Ctrl* __restrict__ ctrl)
{
    cudaGraphSetConditional(handle, ctrl->done ? 0u : 1u);
}

// ---- Akinci2012 boundary volume (BoundaryModel_Akinci2012.cpp:48-75) -------------------------------------------------
template <int W_MODE>
struct BoundaryVolumeF {
    Real4 xi; unsigned i; const SphConst& c; const Real4* __restrict__ bpos; Real delta;
    __device__ __forceinline__ void operator()(unsigned j)
    {
        if (j == i) return;
        const Real4 xj = ld_gather(bpos + j);
        // neighbour predicate first (only list members contribute in the reference)
        if (neighbor_predicate(xi, xj, c.R2)) {
            const Real rx = xi.x - xj.x, ry = xi.y - xj.y, rz = xi.z - xj.z;
            delta += sph_W<W_MODE>(c, rx * rx + ry * ry + rz * rz);
        }
    }
};

template <int W_MODE>
__global__ void __launch_bounds__(DFSPH_BLOCK) k_boundary_volume(unsigned nb, GridDesc g, SphConst c, Real W_zero,
    const Real4* __restrict__ bpos, const unsigned* __restrict__ bcell_start, Real* __restrict__ vol_out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const Real4 xi = ld_gather(bpos + i);
    BoundaryVolumeF<W_MODE> f{xi, i, c, bpos, W_zero};
    walk_candidates(xi, g, bcell_start, f);
    vol_out[i] = (Real)1.0 / f.delta;
}
