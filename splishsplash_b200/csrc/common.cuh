// Shared types of the B200 DFSPH hot path.  One Real per shared object (-DDFSPH_DOUBLE selects double), mirroring
// the reference's compile-time Real (SPlisHSPlasH/Common.h:6-24).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifdef DFSPH_DOUBLE
typedef double Real;
#define DFSPH_REAL_IS_DOUBLE 1
#else
typedef float Real;
#define DFSPH_REAL_IS_DOUBLE 0
#endif

// 4-wide particle record: one 16 B (float) / 32 B (double) vector access per gather.
struct alignas(4 * sizeof(Real)) Real4 { Real x, y, z, w; };
struct Real3 { Real x, y, z; };

__host__ __device__ __forceinline__ Real4 make_real4(Real x, Real y, Real z, Real w) { Real4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

// Vector loads.  ld_gather goes through the read-only (non-coherent) path: used for arrays that are not written by
// the running kernel.  ld_plain is an ordinary (L1-cached) load for arrays with a benign in-kernel writer.
__device__ __forceinline__ Real4 ld_gather(const Real4* p)
{
#if DFSPH_REAL_IS_DOUBLE
    // one 256-bit load (LDG.E.ENL2.256 on sm_100a) instead of two 128-bit ones: a scattered gather costs L1 wavefronts
    // per request, not per byte (profiles/r1_gather_microbench.md: 23 vs 31 cycles for a 32 B record)
    Real4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
#else
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    return make_real4(a.x, a.y, a.z, a.w);
#endif
}
__device__ __forceinline__ Real4 ld_plain(const Real4* p)
{
#if DFSPH_REAL_IS_DOUBLE
    Real4 r;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p) : "memory");
    return r;
#else
    const float4 a = *reinterpret_cast<const float4*>(p);
    return make_real4(a.x, a.y, a.z, a.w);
#endif
}
__device__ __forceinline__ void st_real4(Real4* p, Real4 v)
{
#if DFSPH_REAL_IS_DOUBLE
    reinterpret_cast<double2*>(p)[0] = make_double2(v.x, v.y);
    reinterpret_cast<double2*>(p)[1] = make_double2(v.z, v.w);
#else
    *reinterpret_cast<float4*>(p) = make_float4(v.x, v.y, v.z, v.w);
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// Cell grid and particle order.  Cell edge S >= support radius R (S = R * (1 + 1e-5): two particles that pass the
// predicate l2 < R*R are then at most one cell apart on every axis even with rounding in the cell computation, which is
// done in double).
//
// Particles are stored in PENCIL ORDER: the domain is cut into blocks of DFSPH_BX x DFSPH_BY x DFSPH_BZ cells, blocks
// are ranked along a z-order curve over their block coordinates (host-built rank table, so the tables stay proportional
// to the domain), and inside a block the particles are ordered by fine row -- (z, y) at half-cell resolution -- and then
// by x.  Consecutive particles (= consecutive lanes of a warp) are therefore neighbours ALONG x in a thin pencil, and
// their neighbour lists are near-translates of each other: at every step of a sweep the lanes of a warp gather from
// few distinct 128-byte lines (tools/model/wavefront_model.cpp: 7-9 lines per warp-level gather instead of 12 for a
// Morton order of the cells, on lattice-like states; equal on fully disordered ones).
//
// The cell table has one entry per (block, fine row, cell along x, x-slice): DFSPH_FR^2 * DFSPH_XBINS entries per cell.
// Inside a block the entries of a fine row are contiguous along x, so the neighbour search walks, for each of the
// (2 FR + 1)^2 fine rows around a particle, ONE run of candidates (two when the run crosses a block face), clipped to
// the x-slices that can hold a neighbour (search_kernels.cuh).
// ---------------------------------------------------------------------------------------------------------------
// all powers of two (the table arithmetic uses shifts)
#ifndef DFSPH_BX_LOG2
#define DFSPH_BX_LOG2 4
#endif
#ifndef DFSPH_BY_LOG2
#define DFSPH_BY_LOG2 2
#endif
#ifndef DFSPH_BZ_LOG2
#define DFSPH_BZ_LOG2 2
#endif
#define DFSPH_FR_LOG2 1     /* fine rows per cell in y and in z: 2 */
#ifndef DFSPH_XBINS_LOG2
#define DFSPH_XBINS_LOG2 1  /* x-slices per cell: 2 */
#endif
#define DFSPH_BX (1 << DFSPH_BX_LOG2)
#define DFSPH_BY (1 << DFSPH_BY_LOG2)
#define DFSPH_BZ (1 << DFSPH_BZ_LOG2)
#define DFSPH_FR (1 << DFSPH_FR_LOG2)
#define DFSPH_XBINS (1 << DFSPH_XBINS_LOG2)
#define DFSPH_ENTRIES_PER_BLOCK (1u << (DFSPH_BZ_LOG2 + DFSPH_FR_LOG2 + DFSPH_BY_LOG2 + DFSPH_FR_LOG2 + DFSPH_BX_LOG2 + DFSPH_XBINS_LOG2))

struct GridDesc {
    double ox, oy, oz;      // origin
    double inv_cell;        // 1 / S
    int nx, ny, nz;         // cells per axis
    int nby, nbz;           // blocks per axis (y, z)
    unsigned num_keys;      // table entries: blocks * DFSPH_ENTRIES_PER_BLOCK (the dump cell of the slab sort sits behind them)
    const unsigned* block_rank;   // [nbx*nby*nbz] position of each block along the z-order curve over blocks
    const unsigned* block_of_rank;   // inverse of block_rank: linear block index (bx nby + by) nbz + bz of the r-th block of the curve
};

__host__ __device__ __forceinline__ unsigned spread3(unsigned v)   // 3 bits -> bits 0,3,6
{
    return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4);
}
// table entry of cell (cx, cy, cz), fine row (fy, fz in 0..FR-1) and x-slice xs.  Entries with the same (cy, cz, fy, fz)
// and cx inside one block are contiguous: entry(cx + 1, .., xs = 0) = entry(cx, .., xs = 0) + DFSPH_XBINS.
__host__ __device__ __forceinline__ unsigned cell_entry(int cx, int cy, int cz, unsigned fy, unsigned fz, unsigned xs, const GridDesc& g)
{
    const unsigned ux = (unsigned)cx, uy = (unsigned)cy, uz = (unsigned)cz;
    const unsigned b = ((ux >> DFSPH_BX_LOG2) * (unsigned)g.nby + (uy >> DFSPH_BY_LOG2)) * (unsigned)g.nbz + (uz >> DFSPH_BZ_LOG2);
    const unsigned rz = ((uz & (DFSPH_BZ - 1u)) << DFSPH_FR_LOG2) | fz, ry = ((uy & (DFSPH_BY - 1u)) << DFSPH_FR_LOG2) | fy;
    const unsigned l = ((((rz << (DFSPH_BY_LOG2 + DFSPH_FR_LOG2)) | ry) << DFSPH_BX_LOG2 | (ux & (DFSPH_BX - 1u))) << DFSPH_XBINS_LOG2) | xs;
#ifdef __CUDA_ARCH__
    return __ldg(g.block_rank + b) * DFSPH_ENTRIES_PER_BLOCK + l;
#else
    return b * DFSPH_ENTRIES_PER_BLOCK + l;   // host code never needs the curve position
#endif
}
__host__ __device__ __forceinline__ int cell_coord(Real x, double o, double inv, int n)
{
    const double t = ((double)x - o) * inv;
    int c = (int)floor(t);
    c = c < 0 ? 0 : c;
    return c >= n ? n - 1 : c;
}
// cell coordinate plus the 1/8-cell sub-position (0..7) inside it
__host__ __device__ __forceinline__ int cell_coord_fine(Real x, double o, double inv, int n, unsigned& sub)
{
    const double t = ((double)x - o) * inv;
    const double fl = floor(t);
    int c = (int)fl;
    int s = (int)((t - fl) * 8.0);
    if (c < 0) { c = 0; s = 0; }
    if (c >= n) { c = n - 1; s = 7; }
    sub = (unsigned)(s < 0 ? 0 : (s > 7 ? 7 : s));
    return c;
}
// table entry of a position; `xord` receives a key that orders the particles of one entry by x (ties: source index)
__device__ __forceinline__ unsigned position_entry(const Real4& p, const GridDesc& g, unsigned& xord)
{
    unsigned sx, sy, sz;
    const int cx = cell_coord_fine(p.x, g.ox, g.inv_cell, g.nx, sx);
    const int cy = cell_coord_fine(p.y, g.oy, g.inv_cell, g.ny, sy);
    const int cz = cell_coord_fine(p.z, g.oz, g.inv_cell, g.nz, sz);
    const unsigned fb = __float_as_uint((float)p.x);                 // float order -> unsigned order
    xord = (fb & 0x80000000u) ? ~fb : (fb | 0x80000000u);
    return cell_entry(cx, cy, cz, sy >> (3 - DFSPH_FR_LOG2), sz >> (3 - DFSPH_FR_LOG2), sx >> (3 - DFSPH_XBINS_LOG2), g);
}

// ---------------------------------------------------------------------------------------------------------------
// SPH constants (host-computed in Real exactly as the reference's setRadius functions do).
// ---------------------------------------------------------------------------------------------------------------
enum KernelMode { KM_CUBIC_AVX = 0, KM_CUBIC = 1, KM_LUT = 2, KM_GENERIC = 3 };   // GENERIC: any kernel/gradKernel pair, chosen at run time

struct SphConst {
    Real R;            // support radius (4 r)
    Real R2;           // R*R in Real: neighbour predicate threshold
    Real invR;         // 1/R
    Real invR2;        // invR*invR (CubicKernel_AVX::m_invRadius2)
    Real k, l;         // cubic spline constants 8/(pi R^3), 48/(pi R^3)
    Real W_zero;       // W(0) of the solver kernel
    Real V;            // fluid particle volume (FluidModel::m_V)
    Real density0;
    Real lut_inv_step; // PrecomputedKernel::m_invStepSize
    const Real* lutW;      // [9999]  pre-averaged: 0.5*(m_W[i] + m_W[i+1]), the exact value the reference computes per call
    const Real* lutGradW;  // [9999]  pre-averaged: 0.5*(m_gradW[i] + m_gradW[i+1])
    int mode;          // KernelMode used by the solver sums
    // KM_GENERIC (next-row f2): Simulation "kernel" / "gradKernel" ids 0 cubic, 1 Wendland quintic C2, 2 Poly6, 3 Spiky,
    // 4 precomputed cubic (Simulation.cpp:306-393) and the m_k / m_l constants of kernels 1..3 (SPHKernels.h:107-118,
    // 196-204, 268-277)
    int w_kind, g_kind;
    Real gen_k[4], gen_l[4];
    Real g_a, g_b, g_ml;   // cubic gradient in the form the float sweeps evaluate: 3 l / R^2, -2 l / R^2, -l
    Real gV_a, gV_b, gV_ml; // the same three times the particle volume V (pass A / B / the init sweeps need V gradW, never gradW alone)
};

// Solver control block, resident in device memory so that no host round trip is needed inside a step.
struct Ctrl {
    Real h;                 // TimeManager time step size (updated by the CFL kernel)
    Real h_step;            // h captured at the start of the step (TimeStepDFSPH.cpp:121)
    double time;
    double err_sum;         // reduced density error of the last Jacobi pass
    double avg_err;         // pressure solve
    double avg_err_v;       // divergence solve
    unsigned long long maxvel_bits;   // CFL: max |v + a h|^2 as ordered bits (non-negative IEEE values order like integers)
    unsigned iter;          // running iteration counter of the current solve
    unsigned iterations;    // result: pressure solver iterations
    unsigned iterations_v;  // result: divergence solver iterations
    int done;               // current solve finished
    unsigned ticket;        // last-block election
    unsigned overflow;      // neighbour-table capacity exceeded (value = needed capacity)
    unsigned overflow_b;
    unsigned max_nbr;
    unsigned multi;         // 1: multi-GPU run, loop control happens in k_solve_control after the all-reduce
    unsigned long long n_global;   // particles of all ranks (divisor of the average density error)
    unsigned red_seq;       // sequence number of the fused peer-memory all-reduce
    unsigned fatal;         // sticky: a neighbour list did not fit (overflow > capacity); every later kernel of the context is a no-op
};

struct SolverParams {
    Real gx, gy, gz;
    Real max_error, max_error_v;      // percent
    unsigned min_iter, max_iter, max_iter_v;
    int cfl_method;
    Real cfl_factor, cfl_min, cfl_max;
    Real radius;                      // particle radius
    int viscosity_method;             // 0 none, 1 Viscosity_Standard
    Real viscosity, viscosity_boundary;
};

#ifndef DFSPH_BLOCK
#define DFSPH_BLOCK 256
#endif
#define DFSPH_TILE 32
#define DFSPH_PAD 4u     /* neighbour lists are padded to multiples of this (>= the sweep unroll factors) */
