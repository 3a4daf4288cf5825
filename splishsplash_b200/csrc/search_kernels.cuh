// Neighbourhood search on the device (replaces CompactNSearch::find_neighbors / z_sort /
// sort_field called from SPlisHSPlasH/Simulation.cpp:606-642 and SPlisHSPlasH/FluidModel.cpp:329-360):
//   cell key (blocked z-order) -> counting sort (histogram, exclusive scan = cell-start table, scatter, per-cell
//   fix-up that makes the permutation deterministic) -> reorder of the persistent particle arrays -> per-particle
//   neighbour table for the solver sweeps, laid out warp-tile interleaved so that every table read is a coalesced
//   128 B line.  Predicate (CompactNSearch contract, see oracle/standin/CompactNSearch.h):
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
    unsigned sx, sy, sz;
    const int cx = cell_coord_fine(p.x, g.ox, g.inv_cell, g.nx, sx);
    const int cy = cell_coord_fine(p.y, g.oy, g.inv_cell, g.ny, sy);
    const int cz = cell_coord_fine(p.z, g.oz, g.inv_cell, g.nz, sz);
    const unsigned key = cell_key(cx, cy, cz, g);
    key_out[i] = key;
    fine_out[i] = spread3(sx) | (spread3(sy) << 1) | (spread3(sz) << 2);
    rank_out[i] = atomicAdd(cell_count + key, 1u);
}

__global__ void __launch_bounds__(DFSPH_BLOCK) k_cell_scatter(const unsigned* __restrict__ key, const unsigned* __restrict__ rank, unsigned n,
                                                                const unsigned* __restrict__ cell_start, unsigned* __restrict__ sorted_idx)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sorted_idx[cell_start[key[i]] + rank[i]] = i;
}

// The atomic ranks above depend on thread scheduling.  Sorting every cell segment by (sub-cell Morton code, source
// index) makes the permutation -- and with it every floating-point summation order downstream -- reproducible run to
// run, and continues the z-order curve below the cell level.
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

// One thread per particle of the searching set; walks the 27 cells around its cell in the `other` set's cell table.
// SELF: searching set == found set (skip j == i).
// PERM: the found set is not stored in cell order; `perm` maps cell-table slots to its particles (ghost set) and the
// table receives `base + particle`.  `cnt` continues an existing list (ghosts are appended to the fluid list).
template <bool SELF, bool PERM>
__device__ __forceinline__ unsigned search_cells(const Real4 xi, unsigned i, const GridDesc& g, Real R2,
    const Real4* __restrict__ other_pos, const unsigned* __restrict__ other_cell_start,
    unsigned* __restrict__ tab, unsigned K, unsigned tile, unsigned lane,
    unsigned cnt = 0, const unsigned* __restrict__ perm = nullptr, unsigned base = 0)
{
    const int cx = cell_coord(xi.x, g.ox, g.inv_cell, g.nx);
    const int cy = cell_coord(xi.y, g.oy, g.inv_cell, g.ny);
    const int cz = cell_coord(xi.z, g.oz, g.inv_cell, g.nz);
    unsigned* my = tab + (size_t)tile * K * DFSPH_TILE + lane;
    // z outermost / x innermost = ascending Morton order inside a block: lists come out (nearly) sorted by address
    for (int dz = -1; dz <= 1; ++dz) {
        const int z = cz + dz;
        if (z < 0 || z >= g.nz) continue;
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = cy + dy;
            if (y < 0 || y >= g.ny) continue;
            for (int dx = -1; dx <= 1; ++dx) {
                const int x = cx + dx;
                if (x < 0 || x >= g.nx) continue;
                const unsigned key = cell_key(x, y, z, g);
                const unsigned s = __ldg(other_cell_start + key), e = __ldg(other_cell_start + key + 1);
                for (unsigned k = s; k < e; ++k) {
                    const unsigned j = PERM ? __ldg(perm + k) : k;
                    const Real4 xj = ld_gather(other_pos + j);
                    if (neighbor_predicate(xi, xj, R2) && !(SELF && j == i)) {
                        if (cnt < K) my[(size_t)cnt * DFSPH_TILE] = base + j;
                        ++cnt;
                    }
                }
            }
        }
    }
    return cnt;
}

__global__ void __launch_bounds__(DFSPH_BLOCK) k_build_neighbors(unsigned n, GridDesc g, Real R2,
    const Real4* __restrict__ pos, const unsigned* __restrict__ cell_start,
    const Real4* __restrict__ bpos, const unsigned* __restrict__ bcell_start, unsigned nb,
    unsigned* __restrict__ tab_f, unsigned Kf, unsigned* __restrict__ tab_b, unsigned Kb,
    unsigned* __restrict__ cnt_f, unsigned* __restrict__ cnt_b, unsigned* __restrict__ tcnt_f, unsigned* __restrict__ tcnt_b, Ctrl* ctrl,
    unsigned ng, const unsigned* __restrict__ gcell_start, const unsigned* __restrict__ gperm)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned tile = i >> 5, lane = i & 31u;
    const Real4 xi = ld_gather(pos + i);
    unsigned cf = search_cells<true, false>(xi, i, g, R2, pos, cell_start, tab_f, Kf, tile, lane);
    // multi-GPU: ghost particles of the neighbouring slabs live behind the owned ones at pos[n .. n+ng)
    if (ng > 0) cf = search_cells<false, true>(xi, i, g, R2, pos + n, gcell_start, tab_f, Kf, tile, lane, cf, gperm, n);
    unsigned cb = 0;
    if (nb > 0) cb = search_cells<false, false>(xi, i, g, R2, bpos, bcell_start, tab_b, Kb, tile, lane);
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

