// Neighbourhood search on the device (replaces CompactNSearch::find_neighbors / z_sort /
// sort_field called from SPlisHSPlasH/Simulation.cpp:606-642 and SPlisHSPlasH/FluidModel.cpp:329-360):
//   table entry (pencil order, common.cuh) -> counting sort (histogram, exclusive scan = cell-start table, scatter,
//   per-entry fix-up that orders every entry by x and makes the permutation deterministic) -> reorder of the persistent
//   particle arrays -> per-particle neighbour table for the solver sweeps, laid out warp-tile interleaved so that every
//   table read is a coalesced 128 B line.  Predicate (CompactNSearch contract, see oracle/standin/CompactNSearch.h):
//   l2 = dx*dx; l2 += dy*dy; l2 += dz*dz  with every operation rounded in Real (no FMA), neighbour iff l2 < R*R.
#pragma once
#include "common.cuh"

// ---- exclusive scan of a uint32 array (cell counts -> cell starts) ------------------------------------------------
#define SCAN_ITEMS 8
#define SCAN_BLOCK 256
#define SCAN_CHUNK (SCAN_ITEMS * SCAN_BLOCK)

__device__ __forceinline__ unsigned warp_incl_scan(unsigned v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, total in *total
__device__ __forceinline__ unsigned block_excl_scan(unsigned v, unsigned* total)
{
    __shared__ unsigned warp_sums[SCAN_BLOCK / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned incl = warp_incl_scan(v, lane);
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        unsigned s = lane < (SCAN_BLOCK / 32) ? warp_sums[lane] : 0u;
        s = warp_incl_scan(s, lane);
        if (lane < (SCAN_BLOCK / 32)) warp_sums[lane] = s;
    }
    __syncthreads();
    const unsigned base = w > 0 ? warp_sums[w - 1] : 0u;
    *total = warp_sums[SCAN_BLOCK / 32 - 1];
    __syncthreads();
    return base + incl - v;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_partials(const unsigned* __restrict__ in, unsigned n, unsigned* __restrict__ partial)
{
    const unsigned base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
    unsigned s = 0;
    if (base + SCAN_ITEMS <= n) {
        const uint4 a = *reinterpret_cast<const uint4*>(in + base);
        const uint4 b = *reinterpret_cast<const uint4*>(in + base + 4);
        s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    } else {
        for (unsigned k = 0; k < SCAN_ITEMS; ++k) if (base + k < n) s += in[base + k];
    }
    unsigned total;
    block_excl_scan(s, &total);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

// single block: exclusive scan of the per-chunk totals (in place); writes grand total to partial[nparts]
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_spine(unsigned* partial, unsigned nparts)
{
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < nparts; base += SCAN_BLOCK) {
        const unsigned i = base + threadIdx.x;
        const unsigned v = i < nparts ? partial[i] : 0u;
        unsigned total;
        const unsigned ex = block_excl_scan(v, &total);
        const unsigned c = carry;
        if (i < nparts) partial[i] = ex + c;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[nparts] = carry;
}

// out has n+1 entries; out[n] = grand total
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_apply(const unsigned* __restrict__ in, unsigned n, const unsigned* __restrict__ partial,
                                                             unsigned nparts, unsigned* __restrict__ out)
{
    const unsigned base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
    unsigned v[SCAN_ITEMS];
    unsigned s = 0;
#pragma unroll
    for (unsigned k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        s += v[k];
    }
    unsigned total;
    unsigned ex = block_excl_scan(s, &total) + partial[blockIdx.x];
#pragma unroll
    for (unsigned k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = partial[nparts];
}

// ---- counting sort ---------------------------------------------------------------------------------------------
// pos4.w is not used by the search.
__global__ void __launch_bounds__(DFSPH_BLOCK) k_cell_hash(const Real4* __restrict__ pos, unsigned n, GridDesc g,
                                                             unsigned* __restrict__ cell_count, unsigned* __restrict__ key_out,
                                                             unsigned* __restrict__ rank_out, unsigned* __restrict__ fine_out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Real4 p = pos[i];
    unsigned xord;
    const unsigned key = position_entry(p, g, xord);
    key_out[i] = key;
    fine_out[i] = xord;
    rank_out[i] = atomicAdd(cell_count + key, 1u);
}

__global__ void __launch_bounds__(DFSPH_BLOCK) k_cell_scatter(const unsigned* __restrict__ key, const unsigned* __restrict__ rank, unsigned n,
                                                                const unsigned* __restrict__ cell_start, unsigned* __restrict__ sorted_idx)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sorted_idx[cell_start[key[i]] + rank[i]] = i;
}

// The atomic ranks above depend on thread scheduling.  Sorting every table entry's segment by (x, source index) makes
// the permutation -- and with it every floating-point summation order downstream -- reproducible run to run, and
// completes the pencil order (a fine row is sorted by x across its entries).
__global__ void __launch_bounds__(DFSPH_BLOCK) k_cell_fix_order(const unsigned* __restrict__ cell_start, unsigned num_keys,
                                                                  const unsigned* __restrict__ fine, unsigned* __restrict__ sorted_idx)
{
    const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= num_keys) return;
    const unsigned s = cell_start[c], e = cell_start[c + 1];
    for (unsigned a = s + 1; a < e; ++a) {
        const unsigned v = sorted_idx[a];
        const unsigned long long kv = ((unsigned long long)fine[v] << 32) | v;
        unsigned b = a;
        while (b > s) {
            const unsigned w = sorted_idx[b - 1];
            if ((((unsigned long long)fine[w] << 32) | w) <= kv) break;
            sorted_idx[b] = w;
            --b;
        }
        sorted_idx[b] = v;
    }
}

// Gather the persistent per-particle state into sorted order (replaces PointSet::sort_field on x, v, id, state,
// kappa, kappa_v: FluidModel.cpp:338-346, SimulationDataDFSPH.cpp:97-98).
__global__ void __launch_bounds__(DFSPH_BLOCK) k_reorder(unsigned n, const unsigned* __restrict__ sorted_idx,
    const Real4* __restrict__ pos_in, const Real4* __restrict__ vel_in, const Real* __restrict__ kappa_in, const Real* __restrict__ kappav_in,
    const unsigned* __restrict__ id_in, const unsigned* __restrict__ state_in,
    Real4* __restrict__ pos_out, Real4* __restrict__ vel_out, Real* __restrict__ kappa_out, Real* __restrict__ kappav_out,
    unsigned* __restrict__ id_out, unsigned* __restrict__ state_out, Real4* __restrict__ acc_sentinel)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0) {
        // sentinel particle [n]: far away, at rest, no pressure -> contributes exactly 0 to every sum; the padded
        // slots of the neighbour table point at it
        st_real4(pos_out + n, make_real4((Real)1.0e15, (Real)1.0e15, (Real)1.0e15, (Real)0.0));
        st_real4(vel_out + n, make_real4((Real)0.0, (Real)0.0, (Real)0.0, (Real)0.0));
        st_real4(acc_sentinel + n, make_real4((Real)0.0, (Real)0.0, (Real)0.0, (Real)0.0));
    }
    const unsigned s = sorted_idx[i];
    st_real4(pos_out + i, ld_gather(pos_in + s));
    st_real4(vel_out + i, ld_gather(vel_in + s));
    kappa_out[i] = kappa_in[s];
    kappav_out[i] = kappav_in[s];
    id_out[i] = id_in[s];
    state_out[i] = state_in[s];
}

// Boundary particles are static: sorted once. bpos4 = (x, y, z, V_b)
__global__ void __launch_bounds__(DFSPH_BLOCK) k_reorder_boundary(unsigned n, const unsigned* __restrict__ sorted_idx,
    const Real4* __restrict__ in, const unsigned* __restrict__ orig_in, Real4* __restrict__ out, unsigned* __restrict__ orig_out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0) st_real4(out + n, make_real4((Real)1.0e15, (Real)1.0e15, (Real)1.0e15, (Real)0.0));   // sentinel, V_b = 0
    const unsigned s = sorted_idx[i];
    st_real4(out + i, ld_gather(in + s));
    orig_out[i] = orig_in[s];
}

// ---- neighbour table -------------------------------------------------------------------------------------------
// Table layout: entry k of particle i lives at  tab[((i/32) * K + k) * 32 + (i%32)]  (warp-tile interleaved).
__device__ __forceinline__ bool neighbor_predicate(Real4 a, Real4 b, Real R2)
{
#if DFSPH_REAL_IS_DOUBLE
    const double dx = __dsub_rn(a.x, b.x), dy = __dsub_rn(a.y, b.y), dz = __dsub_rn(a.z, b.z);
    double l2 = __dmul_rn(dx, dx);
    l2 = __dadd_rn(l2, __dmul_rn(dy, dy));
    l2 = __dadd_rn(l2, __dmul_rn(dz, dz));
#else
    const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
    float l2 = __fmul_rn(dx, dx);
    l2 = __fadd_rn(l2, __fmul_rn(dy, dy));
    l2 = __fadd_rn(l2, __fmul_rn(dz, dz));
#endif
    return l2 < R2;
}

// Candidate walk.  Calls f(k) for every slot k of the point set behind `cs` (a cell-start table in pencil order) that can
// hold a neighbour of position xi: for each of the (2 FR + 1)^2 fine rows around the particle, the distance of the
// particle to the row bounds |dx| of any neighbour in it; rows that cannot hold one are skipped and the others are
// clipped to the x-slices inside the bound.  Inside a block a fine row is one contiguous run of slots sorted by x (two
// runs when the clipped range crosses a block face); both runs are walked by ONE flat loop so that lanes with one and
// with two runs stay in step.  The clipping is conservative (margins far above the rounding of the exact predicate and
// of the double-precision cell coordinates); the caller's predicate decides membership.  This leaves ~45 % of the
// candidates of a plain 27-cell walk.
template <class F>
__device__ __forceinline__ void walk_candidates(const Real4 xi, const GridDesc& g, const unsigned* __restrict__ cs, F& f)
{
    // cell coordinates in double exactly as the cell sort computes them (clamped to the grid)
    const double tx = ((double)xi.x - g.ox) * g.inv_cell, ty = ((double)xi.y - g.oy) * g.inv_cell, tz = ((double)xi.z - g.oz) * g.inv_cell;
    const double flx = floor(tx), fly = floor(ty), flz = floor(tz);
    int cx = (int)flx, cy = (int)fly, cz = (int)flz;
    float uy = (float)((ty - fly) * (double)DFSPH_FR), uz = (float)((tz - flz) * (double)DFSPH_FR);   // inside the cell, fine-row units
    float txs = (float)(tx * (double)DFSPH_XBINS);                                                     // x in slice units
    if (cx < 0) { cx = 0; txs = 0.0f; } if (cx >= g.nx) { cx = g.nx - 1; txs = (float)(g.nx * DFSPH_XBINS) - 0.5f; }
    if (cy < 0) { cy = 0; uy = 0.0f; } if (cy >= g.ny) { cy = g.ny - 1; uy = (float)DFSPH_FR - 0.001f; }
    if (cz < 0) { cz = 0; uz = 0.0f; } if (cz >= g.nz) { cz = g.nz - 1; uz = (float)DFSPH_FR - 0.001f; }
    const int iy = __float2int_rd(uy), iz = __float2int_rd(uz);
    const int gy = cy * DFSPH_FR + iy, gz = cz * DFSPH_FR + iz;                 // fine row of the particle
    const float fyf = uy - (float)iy, fzf = uz - (float)iz;                     // position inside the fine row [0, 1)
    const int x0 = cx > 0 ? cx - 1 : 0, x1 = cx + 1 < g.nx ? cx + 1 : g.nx - 1;
    const int f0 = x0 * DFSPH_XBINS, f1 = x1 * DFSPH_XBINS + DFSPH_XBINS - 1;
    const int nry = g.ny * DFSPH_FR, nrz = g.nz * DFSPH_FR;
    // lengths in units of the cell edge S = R (1 + 1e-5): R < S, so the bound R^2 < 1 is conservative; m covers the roundings
    // (of the float arithmetic here, of rsqrt.approx, and of the predicate itself)
    const float m = 2.0e-4f, inv_fr = 1.0f / (float)DFSPH_FR;
    // z outermost / x innermost = ascending entry order inside a block: lists come out (nearly) sorted by address
    for (int dz = -DFSPH_FR; dz <= DFSPH_FR; ++dz) {
        const int rz = gz + dz;
        if (rz < 0 || rz >= nrz) continue;
        const float az = dz < 0 ? fzf + (float)(-dz - 1) : (1.0f - fzf) + (float)(dz - 1);
        const float pz = dz == 0 ? 0.0f : fmaxf(az * inv_fr - m, 0.0f);
        const float wz2 = 1.0f - pz * pz;
        if (wz2 <= 0.0f) continue;
        const unsigned czr = (unsigned)rz >> DFSPH_FR_LOG2, fzr = (unsigned)rz & (DFSPH_FR - 1u);
#pragma unroll
        for (int dy = -DFSPH_FR; dy <= DFSPH_FR; ++dy) {
            const int ry = gy + dy;
            const float ay = dy < 0 ? fyf + (float)(-dy - 1) : (1.0f - fyf) + (float)(dy - 1);
            const float py = dy == 0 ? 0.0f : fmaxf(ay * inv_fr - m, 0.0f);
            const float w2 = wz2 - py * py;
            if (ry < 0 || ry >= nry || w2 <= 0.0f) continue;                      // outside the grid / the whole row is farther than R
            const float ws = (w2 * fast_rsqrt(w2) + m) * (float)DFSPH_XBINS + 0.02f;   // |dx| bound in slices (+ float rounding of txs)
            int fa = __float2int_rd(txs - ws), fb = __float2int_rd(txs + ws);
            fa = fa < f0 ? f0 : fa; fb = fb > f1 ? f1 : fb;
            if (fa > fb) continue;
            const unsigned ca = (unsigned)fa >> DFSPH_XBINS_LOG2, cb = (unsigned)fb >> DFSPH_XBINS_LOG2;
            const unsigned cyr = (unsigned)ry >> DFSPH_FR_LOG2, fyr = (unsigned)ry & (DFSPH_FR - 1u);
            // up to two runs: [s1, e1) in ca's block and [s2, e2) in the next block along x
            const unsigned ka = cell_entry((int)ca, (int)cyr, (int)czr, fyr, fzr, 0u, g);
            unsigned s1 = __ldg(cs + ka + ((unsigned)fa - (ca << DFSPH_XBINS_LOG2))), e1, s2 = 0u, e2 = 0u;
            if ((ca >> DFSPH_BX_LOG2) == (cb >> DFSPH_BX_LOG2)) {
                e1 = __ldg(cs + ka + ((unsigned)fb - (ca << DFSPH_XBINS_LOG2)) + 1u);
            } else {
                const unsigned xs = (cb >> DFSPH_BX_LOG2) << DFSPH_BX_LOG2;          // first cell of the next block
                e1 = __ldg(cs + ka + ((xs - ca) << DFSPH_XBINS_LOG2));                // start of the entry behind the row's last cell in this block
                const unsigned kb = cell_entry((int)xs, (int)cyr, (int)czr, fyr, fzr, 0u, g);
                s2 = __ldg(cs + kb);
                e2 = __ldg(cs + kb + ((unsigned)fb - (xs << DFSPH_XBINS_LOG2)) + 1u);
            }
            unsigned rem = (e1 - s1) + (e2 - s2);
            unsigned k = s1;
            if (s1 == e1) { k = s2; e1 = 0xffffffffu; }
            for (; rem > 0u; --rem) {
                const unsigned kk = k;
                ++k;
                if (k == e1) k = s2;
                f(kk);
            }
        }
    }
}

// Functor of the table build: exact predicate, predicated (branch-free) store into the warp-tile interleaved table.
// SELF: searching set == found set (skip j == i).
// PERM: the found set is not stored in cell order; `perm` maps cell-table slots to its particles (ghost set) and the
// table receives `base + particle`.  `cnt` continues an existing list (ghosts are appended to the fluid list).
template <bool SELF, bool PERM>
struct BuildF {
    Real4 xi; unsigned i; Real R2;
    const Real4* __restrict__ other_pos; const unsigned* __restrict__ perm;
    unsigned* my; unsigned K, base, cnt;
    __device__ __forceinline__ void operator()(unsigned k)
    {
        const unsigned j = PERM ? __ldg(perm + k) : k;
        const Real4 xj = ld_gather(other_pos + j);
        const bool hit = neighbor_predicate(xi, xj, R2) && !(SELF && j == i);
        // a divergent branch per candidate costs more than the store: predicate it
        unsigned* dst = my + (size_t)cnt * DFSPH_TILE;
        asm volatile("{ .reg .pred p; setp.ne.u32 p, %0, 0; @p st.global.u32 [%1], %2; }"
                     :: "r"((unsigned)(hit && cnt < K)), "l"(dst), "r"(base + j) : "memory");
        cnt += hit ? 1u : 0u;
    }
};

template <bool SELF, bool PERM>
__device__ __forceinline__ unsigned search_cells(const Real4 xi, unsigned i, const GridDesc& g, Real R2,
    const Real4* __restrict__ other_pos, const unsigned* __restrict__ other_cell_start,
    unsigned* __restrict__ tab, unsigned K, unsigned tile, unsigned lane,
    unsigned cnt = 0, const unsigned* __restrict__ perm = nullptr, unsigned base = 0)
{
    BuildF<SELF, PERM> f{xi, i, R2, other_pos, perm, tab + (size_t)tile * K * DFSPH_TILE + lane, K, base, cnt};
    walk_candidates(xi, g, other_cell_start, f);
    return f.cnt;
}

// Static boundaries: marks every cell that has a boundary particle in its 3x3x3 neighbourhood, so that the table build
// only walks the boundary set for fluid particles near a wall (one byte per cell, linear index (cx ny + cy) nz + cz).
__global__ void __launch_bounds__(DFSPH_BLOCK) k_mark_boundary_cells(unsigned nb, GridDesc g, const Real4* __restrict__ bpos, unsigned char* __restrict__ near)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const Real4 p = bpos[i];
    const int cx = cell_coord(p.x, g.ox, g.inv_cell, g.nx), cy = cell_coord(p.y, g.oy, g.inv_cell, g.ny), cz = cell_coord(p.z, g.oz, g.inv_cell, g.nz);
    for (int x = max(cx - 1, 0); x <= min(cx + 1, g.nx - 1); ++x)
        for (int y = max(cy - 1, 0); y <= min(cy + 1, g.ny - 1); ++y)
            for (int z = max(cz - 1, 0); z <= min(cz + 1, g.nz - 1); ++z)
                near[((size_t)x * g.ny + y) * g.nz + z] = 1;
}

// ghost_lo / ghost_hi: (multi-GPU) only particles within one cell of a slab face can have ghost neighbours
#ifndef DFSPH_BUILD_MIN_BLOCKS
#define DFSPH_BUILD_MIN_BLOCKS 5   /* 48 registers: the kernel is latency-bound between the table loads of a row and its candidates (2.29 -> 2.13 ms at 10 M) */
#endif
__global__ void __launch_bounds__(DFSPH_BLOCK, DFSPH_BUILD_MIN_BLOCKS) k_build_neighbors(unsigned n, GridDesc g, Real R2,
    const Real4* __restrict__ pos, const unsigned* __restrict__ cell_start,
    const Real4* __restrict__ bpos, const unsigned* __restrict__ bcell_start, unsigned nb, const unsigned char* __restrict__ bnear,
    unsigned* __restrict__ tab_f, unsigned Kf, unsigned* __restrict__ tab_b, unsigned Kb,
    unsigned* __restrict__ cnt_f, unsigned* __restrict__ cnt_b, unsigned* __restrict__ tcnt_f, unsigned* __restrict__ tcnt_b, Ctrl* ctrl,
    unsigned ng, const unsigned* __restrict__ gcell_start, const unsigned* __restrict__ gperm, int slab_axis, double ghost_lo, double ghost_hi)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned tile = i >> 5, lane = i & 31u;
    const Real4 xi = ld_gather(pos + i);
    unsigned cf = search_cells<true, false>(xi, i, g, R2, pos, cell_start, tab_f, Kf, tile, lane);
    // multi-GPU: ghost particles of the neighbouring slabs live behind the owned ones at pos[n .. n+ng)
    if (ng > 0) {
        const double a = slab_axis == 0 ? (double)xi.x : (slab_axis == 1 ? (double)xi.y : (double)xi.z);
        if (a < ghost_lo || a > ghost_hi) cf = search_cells<false, true>(xi, i, g, R2, pos + n, gcell_start, tab_f, Kf, tile, lane, cf, gperm, n);
    }
    unsigned cb = 0;
    if (nb > 0) {
        const int cx = cell_coord(xi.x, g.ox, g.inv_cell, g.nx), cy = cell_coord(xi.y, g.oy, g.inv_cell, g.ny), cz = cell_coord(xi.z, g.oz, g.inv_cell, g.nz);
        if (bnear[((size_t)cx * g.ny + cy) * g.nz + cz]) cb = search_cells<false, false>(xi, i, g, R2, bpos, bcell_start, tab_b, Kb, tile, lane);
    }
    const unsigned sf = cf < Kf ? cf : Kf, sb = cb < Kb ? cb : Kb;
    cnt_f[i] = sf;
    cnt_b[i] = sb;
    if (cf > Kf) atomicMax(&ctrl->overflow, cf);
    if (cb > Kb) atomicMax(&ctrl->overflow_b, cb);
    // pad every list of the warp tile with the sentinel index up to the tile maximum (rounded to DFSPH_PAD; Kf and
    // Kb are multiples of DFSPH_PAD): the solver sweeps then run warp-uniform loops without tail handling
    const unsigned mask = __activemask();
    const unsigned mf = (__reduce_max_sync(mask, sf) + (DFSPH_PAD - 1u)) & ~(DFSPH_PAD - 1u);
    const unsigned mb = (__reduce_max_sync(mask, sb) + (DFSPH_PAD - 1u)) & ~(DFSPH_PAD - 1u);
    unsigned* pf = tab_f + (size_t)tile * Kf * DFSPH_TILE + lane;
    for (unsigned k = sf; k < mf; ++k) pf[(size_t)k * DFSPH_TILE] = n + ng;   // sentinel sits behind the ghosts
    unsigned* pb = tab_b + (size_t)tile * Kb * DFSPH_TILE + lane;
    for (unsigned k = sb; k < mb; ++k) pb[(size_t)k * DFSPH_TILE] = nb;
    if (lane == (unsigned)(__ffs(mask) - 1)) {
        tcnt_f[tile] = mf;
        tcnt_b[tile] = mb;
    }
    const unsigned mx = __reduce_max_sync(mask, cf);
    if (lane == (unsigned)(__ffs(mask) - 1) && mx > 0) atomicMax(&ctrl->max_nbr, mx);
}

