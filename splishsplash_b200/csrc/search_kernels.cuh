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

// Single-pass exclusive scan (chained scan with decoupled look-back): every block publishes its chunk total, then the
// inclusive prefix up to its end, in one 64-bit word per chunk -- (epoch, state, value) -- and the blocks are numbered by an
// atomic ticket in the order they start, so a block only ever waits for blocks that are already running.  One read and one
// write of the array instead of two reads, one write and a serial spine kernel (0.31 -> 0.09 ms for the 50 M-entry table
// of the 10 M-particle scene).  `epoch` must differ from call to call on the same `status` array (no reset needed).
#define SCAN1_ITEMS 16
#define SCAN1_CHUNK (SCAN1_ITEMS * SCAN_BLOCK)
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_lookback(const unsigned* __restrict__ in, unsigned n, unsigned* __restrict__ out,
                                                                unsigned long long* status, unsigned* ticket, unsigned epoch)
{
    __shared__ unsigned s_bid, s_prefix;
    if (threadIdx.x == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned bid = s_bid;
    const unsigned base = bid * SCAN1_CHUNK + threadIdx.x * SCAN1_ITEMS;
    unsigned v[SCAN1_ITEMS];
    unsigned s = 0;
    if (base + SCAN1_ITEMS <= n) {
#pragma unroll
        for (unsigned k = 0; k < SCAN1_ITEMS; k += 4) {
            const uint4 a = *reinterpret_cast<const uint4*>(in + base + k);
            v[k] = a.x; v[k + 1] = a.y; v[k + 2] = a.z; v[k + 3] = a.w;
        }
    } else {
#pragma unroll
        for (unsigned k = 0; k < SCAN1_ITEMS; ++k) v[k] = base + k < n ? in[base + k] : 0u;
    }
#pragma unroll
    for (unsigned k = 0; k < SCAN1_ITEMS; ++k) s += v[k];
    unsigned total;
    unsigned ex = block_excl_scan(s, &total);
    const unsigned long long tag = (unsigned long long)epoch << 34;
    if (threadIdx.x < 32) {
        // warp 0 looks back 32 chunks at a time
        const unsigned lane = threadIdx.x;
        unsigned prefix = 0u;
        if (bid == 0u) {
            if (lane == 0u) atomicExch(status, tag | (2ull << 32) | total);
        } else {
            if (lane == 0u) atomicExch(status + bid, tag | (1ull << 32) | total);      // chunk total is known
            int j0 = (int)bid - 1;                                                      // lane l inspects chunk j0 - l
            while (true) {
                const int j = j0 - (int)lane;
                unsigned long long st = tag | (2ull << 32);                              // before chunk 0: an inclusive prefix of 0
                if (j >= 0) do { st = *reinterpret_cast<volatile unsigned long long*>(status + j); } while ((st >> 34) != epoch);
                const unsigned incl = __ballot_sync(0xffffffffu, ((st >> 32) & 3ull) == 2ull);
                const unsigned first = incl ? (unsigned)__ffs(incl) - 1u : 32u;          // nearest chunk with an inclusive prefix
                unsigned v0 = lane <= first ? (unsigned)st : 0u;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v0 += __shfl_xor_sync(0xffffffffu, v0, d);
                prefix += v0;
                if (incl) break;
                j0 -= 32;
            }
            if (lane == 0u) atomicExch(status + bid, tag | (2ull << 32) | (unsigned long long)(prefix + total));
        }
        if (lane == 0u) {
            s_prefix = prefix;
            if (bid == gridDim.x - 1u) out[n] = prefix + total;             // grand total
        }
    }
    __syncthreads();
    ex += s_prefix;
    if (base + SCAN1_ITEMS <= n) {
#pragma unroll
        for (unsigned k = 0; k < SCAN1_ITEMS; k += 4) {
            uint4 o;
            o.x = ex; ex += v[k]; o.y = ex; ex += v[k + 1]; o.z = ex; ex += v[k + 2]; o.w = ex; ex += v[k + 3];
            *reinterpret_cast<uint4*>(out + base + k) = o;
        }
    } else {
#pragma unroll
        for (unsigned k = 0; k < SCAN1_ITEMS; ++k) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
    }
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
    {   // most of a table is air: a CTA whose entries hold nothing leaves after two loads
        const unsigned c0 = blockIdx.x * blockDim.x, c1 = min(c0 + blockDim.x, num_keys);
        if (__ldg(cell_start + c0) == __ldg(cell_start + c1)) return;
    }
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

// k_cell_fix_order + k_reorder in one pass for the fluid set: the thread of scattered slot a finds the rank of its particle
// inside its table entry by comparing (x-order key, source index) with the entry's other particles -- entries hold one to
// three particles, and every thread works, where the per-entry insertion sort left most lanes idle behind dependent
// loads -- and moves the particle's state straight to its final slot.  Same permutation as the two kernels.
// n = sorted points (multi-GPU: including the leavers, which sit in the dump cell behind everything else), n_keep = owned ones.
__global__ void __launch_bounds__(DFSPH_BLOCK) k_fix_reorder(unsigned n, unsigned n_keep, const unsigned* __restrict__ scattered_idx, const unsigned* __restrict__ key,
    const unsigned* __restrict__ fine, const unsigned* __restrict__ cell_start, unsigned* __restrict__ sorted_idx_out,
    const Real4* __restrict__ pos_in, const Real4* __restrict__ vel_in, const Real* __restrict__ kappa_in, const Real* __restrict__ kappav_in,
    const unsigned* __restrict__ id_in, const unsigned* __restrict__ state_in,
    Real4* __restrict__ pos_out, Real4* __restrict__ vel_out, Real* __restrict__ kappa_out, Real* __restrict__ kappav_out,
    unsigned* __restrict__ id_out, unsigned* __restrict__ state_out, Real4* __restrict__ acc_sentinel)
{
    const unsigned a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    if (a == 0) {   // sentinel particle [n_keep], see k_reorder
        st_real4(pos_out + n_keep, make_real4((Real)1.0e15, (Real)1.0e15, (Real)1.0e15, (Real)0.0));
        st_real4(vel_out + n_keep, make_real4((Real)0.0, (Real)0.0, (Real)0.0, (Real)0.0));
        st_real4(acc_sentinel + n_keep, make_real4((Real)0.0, (Real)0.0, (Real)0.0, (Real)0.0));
    }
    if (a >= n_keep) return;    // slots of the dump cell: particles that left the slab
    const unsigned v = scattered_idx[a];
    const Real4 p = ld_gather(pos_in + v), u = ld_gather(vel_in + v);
    const Real k1 = kappa_in[v], k2 = kappav_in[v];
    const unsigned id = id_in[v], st = state_in[v];
    const unsigned c = key[v];
    const unsigned s = __ldg(cell_start + c), e = __ldg(cell_start + c + 1);
    const unsigned long long kv = ((unsigned long long)fine[v] << 32) | v;
    unsigned rank = 0;
    for (unsigned b = s; b < e; ++b) {
        const unsigned w = scattered_idx[b];
        rank += ((((unsigned long long)fine[w] << 32) | w) < kv) ? 1u : 0u;
    }
    const unsigned i = s + rank;
    sorted_idx_out[i] = v;
    st_real4(pos_out + i, p);
    st_real4(vel_out + i, u);
    kappa_out[i] = k1; kappav_out[i] = k2; id_out[i] = id; state_out[i] = st;
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

// ---- table build from shared-memory tiles: one CTA per block part ---------------------------------------------------------
// In pencil order the particles of a block are one contiguous range, and everything a particle of the block can have as
// a neighbour lies in (TB_SY x TB_SZ) fine rows: the CTA's own rows plus a halo of one cell (DFSPH_FR fine rows) in y and z,
// each row extended by one cell (DFSPH_XBINS slices) into the blocks before and behind it along x.  Each of these rows is
// up to three contiguous runs of the sorted point array (previous block | this block | next block).  The CTA copies
// them into shared memory back to back -- (x, y, z, point index) per record, so one LDS.128 yields everything a
// candidate test needs -- together with a table of slice boundaries (rs[row][slice] = shared-memory slot).  A thread
// then walks the fine rows around its particle exactly as walk_candidates does, but a row is ONE run of shared-memory
// slots: no block-rank lookup, no two-run loop, no global-memory latency per candidate, and the row set-up is branch-free.
// Candidate order (rows z-major, x ascending inside a row) and predicate are those of the one-thread walk, so the lists
// come out identical.  The same pass runs a second time over the static boundary points for the CTAs that have any in
// reach (one uniform test per CTA instead of a divergent one per particle).  A block part whose rows do not fit the
// staging buffer (far denser than the rest state) falls back to the one-thread walk.
#if DFSPH_REAL_IS_DOUBLE
#define DFSPH_TB_TZ 2          /* fine rows (z) of the block per CTA */
#define DFSPH_TB_CAP 3072      /* staged records (32 B each) */
#define DFSPH_TB_THREADS 256
#else
#define DFSPH_TB_TZ 4
#define DFSPH_TB_CAP 5120      /* staged records (16 B each); the rest state needs ~3500 */
#define DFSPH_TB_THREADS 512
#endif
#define TB_RY (DFSPH_BY * DFSPH_FR)              /* fine rows of a block in y */
#define TB_RZ (DFSPH_BZ * DFSPH_FR)
#define TB_SL (DFSPH_BX * DFSPH_XBINS)           /* x-slices of a block row = table entries per row */
#define TB_SY (TB_RY + 2 * DFSPH_FR)             /* staged rows in y */
#define TB_SZ (DFSPH_TB_TZ + 2 * DFSPH_FR)
#define TB_ROWS (TB_SY * TB_SZ)
#define TB_SSL (TB_SL + 2 * DFSPH_XBINS)         /* staged slices per row */
#define TB_PARTS (TB_RZ / DFSPH_TB_TZ)           /* CTAs per block */
#define TB_SMEM_BYTES ((size_t)DFSPH_TB_CAP * sizeof(Real4) + (size_t)TB_ROWS * (TB_SSL + 1) * sizeof(unsigned short) + 2048 /* TB_MAXP row ids */)
#define TB_NONE 0xffffffffu

struct TileGeom {
    int bx, by, bz;          // block
    int gy0, gz0;            // fine row of staged row (0, 0)
    unsigned nbx;
    double X0, Y0, Z0;       // lower corner of staged slice 0 / staged row (0, 0)
    float sx, syz;           // staged x-slices / fine rows per unit length
    float x_lo, x_hi;        // the grid's x-range in staged slice units (out-of-grid particles are filed in the edge cells)
};
#define TB_MAXP 2048         /* particles of one CTA that can be walked from the tile (own-row table); more: one-thread walk */
struct TileRuns {
    unsigned seg_e[3][TB_ROWS];      // first table entry of the run (previous | this | next block), TB_NONE: no such block
    unsigned seg_gs[3][TB_ROWS];     // first point of the run
    unsigned seg_len[3][TB_ROWS];
    unsigned rowoff[TB_ROWS + 1];    // shared-memory slot of the row's first record
};

__device__ __forceinline__ Real4 tile_record(const Real4* __restrict__ pts, unsigned j)
{
    Real4 t = ld_gather(pts + j);
#if DFSPH_REAL_IS_DOUBLE
    t.w = __longlong_as_double((long long)j);
#else
    t.w = __uint_as_float(j);
#endif
    return t;
}
__device__ __forceinline__ unsigned tile_record_index(const Real4& r)
{
#if DFSPH_REAL_IS_DOUBLE
    return (unsigned)__double_as_longlong(r.w);
#else
    return __float_as_uint(r.w);
#endif
}

__device__ __forceinline__ TileGeom tile_geometry(const GridDesc& g, unsigned nbx, unsigned r, unsigned zp)
{
    const unsigned lin = __ldg(g.block_of_rank + r);
    TileGeom tg;
    tg.bz = (int)(lin % (unsigned)g.nbz); tg.by = (int)((lin / (unsigned)g.nbz) % (unsigned)g.nby); tg.bx = (int)(lin / ((unsigned)g.nbz * (unsigned)g.nby));
    tg.gy0 = tg.by * TB_RY - DFSPH_FR; tg.gz0 = tg.bz * TB_RZ + (int)zp * DFSPH_TB_TZ - DFSPH_FR;
    tg.nbx = nbx;
    const double S = 1.0 / g.inv_cell;
    const int xs0 = tg.bx * TB_SL - DFSPH_XBINS;             // global x-slice of staged slice 0
    tg.X0 = g.ox + (double)xs0 * (S / (double)DFSPH_XBINS);
    tg.Y0 = g.oy + (double)tg.gy0 * (S / (double)DFSPH_FR);
    tg.Z0 = g.oz + (double)tg.gz0 * (S / (double)DFSPH_FR);
    tg.sx = (float)(g.inv_cell * (double)DFSPH_XBINS);
    tg.syz = (float)(g.inv_cell * (double)DFSPH_FR);
    tg.x_lo = (float)(-xs0);                                  // particles left of the grid are filed in slice 0 ...
    tg.x_hi = (float)(g.nx * DFSPH_XBINS - xs0) - 0.5f;       // ... and those right of it in the last slice
    return tg;
}

// The runs (previous | this | next block) behind staged row q of the point set with cell table cs.
__device__ __forceinline__ void tile_row_runs(const GridDesc& g, const TileGeom& tg, const unsigned* __restrict__ cs, unsigned q, unsigned e[3], unsigned gs[3], unsigned len[3])
{
    const unsigned nby = (unsigned)g.nby, nbz = (unsigned)g.nbz;
    const int gy = tg.gy0 + (int)(q % TB_SY), gz = tg.gz0 + (int)(q / TB_SY);
    const bool row_ok = gy >= 0 && gz >= 0 && (gy >> (DFSPH_BY_LOG2 + DFSPH_FR_LOG2)) < (int)nby && (gz >> (DFSPH_BZ_LOG2 + DFSPH_FR_LOG2)) < (int)nbz;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const int sbx = tg.bx - 1 + p;
        e[p] = TB_NONE; gs[p] = 0u; len[p] = 0u;
        if (row_ok && sbx >= 0 && sbx < (int)tg.nbx) {
            const unsigned sl = ((unsigned)sbx * nby + (unsigned)(gy >> (DFSPH_BY_LOG2 + DFSPH_FR_LOG2))) * nbz + (unsigned)(gz >> (DFSPH_BZ_LOG2 + DFSPH_FR_LOG2));
            const unsigned row = (((unsigned)gz & (TB_RZ - 1u)) << (DFSPH_BY_LOG2 + DFSPH_FR_LOG2)) | ((unsigned)gy & (TB_RY - 1u));
            const unsigned e0 = __ldg(g.block_rank + sl) * DFSPH_ENTRIES_PER_BLOCK + row * TB_SL;
            e[p] = p == 0 ? e0 + (TB_SL - DFSPH_XBINS) : e0;
            const unsigned ee = p == 1 ? e0 + TB_SL : e[p] + DFSPH_XBINS;
            gs[p] = __ldg(cs + e[p]);
            len[p] = __ldg(cs + ee) - gs[p];
        }
    }
}

// Static boundaries: which block parts have boundary points in their staged rows at all (run once per boundary / grid;
// the table build then skips the boundary pass of every other part without a single load or barrier).
__global__ void __launch_bounds__(128) k_mark_boundary_parts(GridDesc g, unsigned nbx, unsigned nparts, const unsigned* __restrict__ bcs, unsigned char* __restrict__ part_near)
{
    const unsigned part = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
    if (part >= nparts) return;
    const TileGeom tg = tile_geometry(g, nbx, part / TB_PARTS, part % TB_PARTS);
    unsigned any = 0u;
    for (unsigned q = lane; q < TB_ROWS; q += 32u) {
        unsigned e[3], gs[3], len[3];
        tile_row_runs(g, tg, bcs, q, e, gs, len);
        any |= len[0] | len[1] | len[2];
    }
    any = __reduce_or_sync(0xffffffffu, any);
    if (lane == 0u) part_near[part] = any ? 1 : 0;
}

// Stages the rows of the point set (pts, cs) around the CTA.  Returns the number of staged records (uniform over the CTA);
// nothing is copied when that number is 0 or exceeds DFSPH_TB_CAP.  Ends with a barrier.
// rowq != nullptr (fluid pass): rowq[i - t0] receives the staged row of every particle i of the CTA's own rows.
__device__ __forceinline__ unsigned tile_stage(const GridDesc& g, const TileGeom& tg, const Real4* __restrict__ pts, const unsigned* __restrict__ cs,
                                               Real4* srec, unsigned short* rs, TileRuns& tr, unsigned char* rowq, unsigned t0)
{
    __syncthreads();   // the previous pass is done with the buffers
    for (unsigned q = threadIdx.x; q < TB_ROWS; q += blockDim.x) {
        unsigned e[3], gs[3], len[3];
        tile_row_runs(g, tg, cs, q, e, gs, len);
#pragma unroll
        for (int p = 0; p < 3; ++p) { tr.seg_e[p][q] = e[p]; tr.seg_gs[p][q] = gs[p]; tr.seg_len[p][q] = len[p]; }
    }
    __syncthreads();
    if (threadIdx.x < 32) {   // exclusive scan of the row lengths
        constexpr unsigned PER = (TB_ROWS + 31) / 32;
        unsigned v[PER], sum = 0u;
#pragma unroll
        for (unsigned k = 0; k < PER; ++k) {
            const unsigned q = threadIdx.x * PER + k;
            v[k] = q < TB_ROWS ? tr.seg_len[0][q] + tr.seg_len[1][q] + tr.seg_len[2][q] : 0u;
            sum += v[k];
        }
        unsigned ex = warp_incl_scan(sum, (int)threadIdx.x) - sum;
#pragma unroll
        for (unsigned k = 0; k < PER; ++k) {
            const unsigned q = threadIdx.x * PER + k;
            if (q < TB_ROWS) tr.rowoff[q] = ex;
            ex += v[k];
        }
        if (threadIdx.x == 31) tr.rowoff[TB_ROWS] = ex;
    }
    __syncthreads();
    const unsigned staged = tr.rowoff[TB_ROWS];
    if (staged == 0u || staged > DFSPH_TB_CAP) return staged;
    // slice boundaries: a warp per row, a lane per slice
    {
        const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, nw = blockDim.x >> 5;
        for (unsigned q = warp; q < TB_ROWS; q += nw) {
            const unsigned ro = tr.rowoff[q], lL = tr.seg_len[0][q], lC = tr.seg_len[1][q];
#pragma unroll
            for (unsigned la = lane; la < TB_SSL + 1; la += 32u) {
                const unsigned p = la < DFSPH_XBINS ? 0u : (la < DFSPH_XBINS + TB_SL ? 1u : 2u);
                const unsigned k = p == 0u ? la : (p == 1u ? la - DFSPH_XBINS : la - (DFSPH_XBINS + TB_SL));
                unsigned off = ro + (p >= 1u ? lL : 0u) + (p >= 2u ? lC : 0u);
                const unsigned e = tr.seg_e[p][q];
                if (e != TB_NONE) off += __ldg(cs + e + k) - tr.seg_gs[p][q];
                rs[q * (TB_SSL + 1) + la] = (unsigned short)off;
            }
        }
    }
    // records: a warp per row, several rows in flight
    {
        const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, nw = blockDim.x >> 5;
        constexpr int U = 3;
        for (unsigned q0 = warp; q0 < TB_ROWS; q0 += nw * U) {
            Real4 rec[U];
            bool have[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned q = q0 + (unsigned)u * nw;
                have[u] = false;
                if (q < TB_ROWS) {
                    const unsigned lL = tr.seg_len[0][q], lC = tr.seg_len[1][q], tot = lL + lC + tr.seg_len[2][q];
                    if (lane < tot) {
                        const unsigned j = lane < lL ? tr.seg_gs[0][q] + lane : (lane < lL + lC ? tr.seg_gs[1][q] + (lane - lL) : tr.seg_gs[2][q] + (lane - lL - lC));
                        rec[u] = tile_record(pts, j);
                        have[u] = true;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned q = q0 + (unsigned)u * nw;
                if (have[u]) srec[tr.rowoff[q] + lane] = rec[u];
                if (q < TB_ROWS) {
                    const unsigned lL = tr.seg_len[0][q], lC = tr.seg_len[1][q], tot = lL + lC + tr.seg_len[2][q];
                    for (unsigned k = lane + 32u; k < tot; k += 32u) {   // rows longer than a warp
                        const unsigned j = k < lL ? tr.seg_gs[0][q] + k : (k < lL + lC ? tr.seg_gs[1][q] + (k - lL) : tr.seg_gs[2][q] + (k - lL - lC));
                        srec[tr.rowoff[q] + k] = tile_record(pts, j);
                    }
                    const unsigned qy = q % TB_SY, qz = q / TB_SY;
                    if (rowq && qy >= DFSPH_FR && qy < DFSPH_FR + TB_RY && qz >= DFSPH_FR && qz < DFSPH_FR + DFSPH_TB_TZ) {   // own row: its centre run are CTA particles
                        const unsigned first = tr.seg_gs[1][q] - t0;
                        for (unsigned k = lane; k < lC; k += 32u) if (first + k < TB_MAXP) rowq[first + k] = (unsigned char)q;
                    }
                }
            }
        }
    }
    __syncthreads();
    return staged;
}

// Where a particle sits relative to the staged rows (its own row q comes from the staging pass, i.e. from the cell sort
// itself; the fractions only feed the conservative row bounds, so float arithmetic relative to the tile corner suffices).
struct TileWalker {
    float fyf, fzf, txs;     // position inside its fine row [0, 1] (y, z); x in staged slice units
    int row;                 // index of the particle's own row in rs
};
__device__ __forceinline__ TileWalker tile_locate(const Real4& xi, unsigned q, const TileGeom& tg)
{
    TileWalker w;
    const float xr = (float)((double)xi.x - tg.X0) * tg.sx, yr = (float)((double)xi.y - tg.Y0) * tg.syz, zr = (float)((double)xi.z - tg.Z0) * tg.syz;
    const unsigned qy = q % TB_SY, qz = q / TB_SY;
    w.fyf = fminf(fmaxf(yr - (float)qy, 0.0f), 1.0f);
    w.fzf = fminf(fmaxf(zr - (float)qz, 0.0f), 1.0f);
    w.txs = fminf(fmaxf(xr, tg.x_lo), tg.x_hi);
    w.row = (int)q * (TB_SSL + 1);
    return w;
}

// Shared-memory accesses of the walk through explicit 32-bit shared addresses: with generic pointers the compiler
// re-derives the shared window base (S2R SR_CgaCtaId + LEA) in every row, ~8 instructions of 45.
__device__ __forceinline__ Real4 tile_lds_record(unsigned a)
{
    Real4 r;
#if DFSPH_REAL_IS_DOUBLE
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(a));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+16];" : "=d"(r.z), "=d"(r.w) : "r"(a));
#else
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a));
#endif
    return r;
}
__device__ __forceinline__ unsigned tile_lds_u16(unsigned a)
{
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return (unsigned)v;
}

// Walks the staged candidates of one particle; writes the list to my[k * DFSPH_TILE] and returns its length.  The length may
// exceed K: hits beyond the capacity all land in the last slot (the step fails on the reported overflow anyway), which
// costs one integer min per candidate instead of a capacity test per row and a second loop.
template <bool SELF>
__device__ __forceinline__ unsigned tile_walk(const Real4& xi, unsigned i, Real R2, const TileWalker& w, unsigned srec_a, unsigned rs_a,
                                              unsigned* __restrict__ my, unsigned K)
{
    const float m = 2.0e-4f, inv_fr = 1.0f / (float)DFSPH_FR;
    float py2[2 * DFSPH_FR + 1];
#pragma unroll
    for (int dy = -DFSPH_FR; dy <= DFSPH_FR; ++dy) {
        const float ay = dy < 0 ? w.fyf + (float)(-dy - 1) : (1.0f - w.fyf) + (float)(dy - 1);
        const float py = dy == 0 ? 0.0f : fmaxf(ay * inv_fr - m, 0.0f);
        py2[dy + DFSPH_FR] = py * py;
    }
    unsigned off = 0u;                                   // element offset of the next free slot: my[off]
    const unsigned off_last = (K - 1u) * DFSPH_TILE;
#pragma unroll 1
    for (int dz = -DFSPH_FR; dz <= DFSPH_FR; ++dz) {
        const float az = dz < 0 ? w.fzf + (float)(-dz - 1) : (1.0f - w.fzf) + (float)(dz - 1);
        const float pz = dz == 0 ? 0.0f : fmaxf(az * inv_fr - m, 0.0f);
        const float wz2 = 1.0f - pz * pz;
        if (wz2 <= 0.0f) continue;
        const unsigned rrow_a = rs_a + 2u * (unsigned)(w.row + dz * (TB_SY * (TB_SSL + 1)));
#pragma unroll
        for (int dy = -DFSPH_FR; dy <= DFSPH_FR; ++dy) {
            // branch-free row set-up: a row that is out of reach gets an empty run
            const float w2 = wz2 - py2[dy + DFSPH_FR];
            const float w2c = fmaxf(w2, 1.0e-12f);
            const float ws = (w2c * fast_rsqrt(w2c) + m) * (float)DFSPH_XBINS + 0.02f;   // |dx| bound in slices (+ float rounding of txs)
            int fa = __float2int_rd(w.txs - ws), fb = __float2int_rd(w.txs + ws);
            fa = fa < 0 ? 0 : fa; fb = fb > TB_SSL - 1 ? TB_SSL - 1 : fb;     // the staged slices are all a neighbour can be in
            const unsigned ra = rrow_a + (unsigned)(dy * (TB_SSL + 1) * 2);
            const unsigned s = tile_lds_u16(ra + 2u * (unsigned)fa);
            unsigned e = tile_lds_u16(ra + 2u * (unsigned)fb + 2u);
            e = w2 > 0.0f ? e : s;
            unsigned sp = srec_a + s * (unsigned)sizeof(Real4);
            const unsigned sp_end = srec_a + e * (unsigned)sizeof(Real4);
#pragma unroll 1
            for (; sp < sp_end; sp += (unsigned)sizeof(Real4)) {
                const Real4 xj = tile_lds_record(sp);
                const unsigned j = tile_record_index(xj);
                const bool hit = neighbor_predicate(xi, xj, R2) && !(SELF && j == i);
                // predicated store (a divergent branch per candidate costs more than the store)
                const unsigned ow = off < off_last ? off : off_last;
                asm volatile("{ .reg .pred p; .reg .u64 a; setp.ne.u32 p, %1, 0; mad.wide.u32 a, %2, 4, %3; @p st.global.u32 [a], %4; @p add.u32 %0, %0, 32; }"
                             : "+r"(off) : "r"((unsigned)hit), "r"(ow), "l"(my), "r"(j) : "memory");
            }
        }
    }
    return off / DFSPH_TILE;
}

#define TB_INCOMPLETE 0xffffffffu   /* tcnt marker: the tile is shared by two CTAs (or ghosts follow); k_build_neighbors<false> finishes it */

// One pass (fluid or boundary lists) over the CTA's particles [t0, t1), a warp per table tile.  Tiles that lie completely
// inside the range are finished here: counts clamped, lists padded with `sentinel` up to the tile maximum (rounded to
// DFSPH_PAD), tcnt written, overflow raised.  Tiles shared with the neighbouring CTA get the raw counts and the
// TB_INCOMPLETE marker.  mode 0: nothing staged (all lists empty), 1: walk the tile, 2: one-thread walk over global memory.
template <bool SELF>
__device__ __forceinline__ void tile_pass(int mode, unsigned t0, unsigned t1, unsigned n, bool finish_tiles, const GridDesc& g, const TileGeom& tg, Real R2,
    const Real4* __restrict__ pos, const Real4* __restrict__ pts, const unsigned* __restrict__ cs, unsigned srec_a, unsigned rs_a, const unsigned char* rowq,
    unsigned* __restrict__ tab, unsigned K, unsigned* __restrict__ cnt, unsigned* __restrict__ tcnt, unsigned sentinel, unsigned* overflow, unsigned* max_out,
    int slab_axis, double ghost_lo, double ghost_hi)
{
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, nw = blockDim.x >> 5;
    unsigned wmax = 0u;
    for (unsigned base = (t0 & ~31u) + warp * 32u; base < t1; base += nw * 32u) {
        const unsigned i = base + lane;
        const bool active = i >= t0 && i < t1;
        unsigned* my = tab + (size_t)(base >> 5) * K * DFSPH_TILE + lane;
        unsigned c = 0u;
        bool ghosts = false;      // multi-GPU: this particle can have ghost neighbours, which k_build_neighbors<false> appends
        if (active && (mode != 0 || slab_axis >= 0)) {
            const Real4 xi = ld_gather(pos + i);
            if (slab_axis >= 0) {
                const double a = slab_axis == 0 ? (double)xi.x : (slab_axis == 1 ? (double)xi.y : (double)xi.z);
                ghosts = a < ghost_lo || a > ghost_hi;
            }
            if (mode == 0) { }
            else if (mode == 1) c = tile_walk<SELF>(xi, i, R2, tile_locate(xi, rowq[i - t0], tg), srec_a, rs_a, my, K);
            else c = search_cells<SELF, false>(xi, i, g, R2, pts, cs, tab, K, base >> 5, lane);
        }
        const bool whole = finish_tiles && base >= t0 && (base + 32u <= t1 || t1 == n)   // no other CTA has particles in this tile
                           && !__any_sync(0xffffffffu, ghosts);
        if (whole) {
            const unsigned sc = c < K ? c : K;
            if (c > K) atomicMax(overflow, c);
            const unsigned mt = (__reduce_max_sync(0xffffffffu, sc) + (DFSPH_PAD - 1u)) & ~(DFSPH_PAD - 1u);
            if (i < n) {
                cnt[i] = sc;
                for (unsigned k = sc; k < mt; ++k) my[(size_t)k * DFSPH_TILE] = sentinel;
            }
            if (lane == 0u) tcnt[base >> 5] = mt;
            wmax = max(wmax, c);
        } else {
            if (active) cnt[i] = c;
            if (lane == 0u) tcnt[base >> 5] = TB_INCOMPLETE;
        }
    }
    if (max_out) {
        wmax = __reduce_max_sync(0xffffffffu, wmax);
        if (lane == 0u && wmax > 0u) atomicMax(max_out, wmax);
    }
}

// slab_axis >= 0 (multi-GPU): tiles with a particle within one cell of a slab face (outside [ghost_lo, ghost_hi]) are left
// to k_build_neighbors<false>, which appends the ghost neighbours first.
__global__ void __launch_bounds__(DFSPH_TB_THREADS, 2) k_build_tiles(GridDesc g, unsigned nbx, Real R2, unsigned n, int finish_tiles,
    const Real4* __restrict__ pos, const unsigned* __restrict__ cs, unsigned* __restrict__ tab_f, unsigned Kf, unsigned* __restrict__ cnt_f, unsigned* __restrict__ tcnt_f,
    const Real4* __restrict__ bpos, const unsigned* __restrict__ bcs, unsigned nb, unsigned* __restrict__ tab_b, unsigned Kb, unsigned* __restrict__ cnt_b, unsigned* __restrict__ tcnt_b,
    unsigned sentinel_f, const unsigned char* __restrict__ bpart_near, Ctrl* ctrl, int slab_axis, double ghost_lo, double ghost_hi)
{
    extern __shared__ __align__(16) unsigned char tb_smem[];
    Real4* srec = reinterpret_cast<Real4*>(tb_smem);
    unsigned short* rs = reinterpret_cast<unsigned short*>(srec + DFSPH_TB_CAP);    // [TB_ROWS][TB_SSL + 1]
    unsigned char* rowq = reinterpret_cast<unsigned char*>(rs + TB_ROWS * (TB_SSL + 1));   // [TB_MAXP]
    __shared__ TileRuns tr;

    const unsigned r = blockIdx.x / TB_PARTS, zp = blockIdx.x % TB_PARTS;
    const unsigned e_blk = r * DFSPH_ENTRIES_PER_BLOCK + zp * (DFSPH_TB_TZ * TB_RY * TB_SL);
    const unsigned t0 = __ldg(cs + e_blk), t1 = __ldg(cs + e_blk + DFSPH_TB_TZ * TB_RY * TB_SL);   // this CTA's particles
    if (t0 == t1) return;
    const TileGeom tg = tile_geometry(g, nbx, r, zp);
    const bool fits = t1 - t0 <= TB_MAXP;
    const unsigned srec_a = (unsigned)__cvta_generic_to_shared(srec), rs_a = (unsigned)__cvta_generic_to_shared(rs);

    // ---- fluid-fluid lists ----------------------------------------------------------------------------------------------
    const unsigned staged = tile_stage(g, tg, pos, cs, srec, rs, tr, rowq, t0);
    tile_pass<true>(staged <= DFSPH_TB_CAP && fits ? 1 : 2, t0, t1, n, finish_tiles != 0, g, tg, R2, pos, pos, cs, srec_a, rs_a, rowq,
                    tab_f, Kf, cnt_f, tcnt_f, sentinel_f, &ctrl->overflow, &ctrl->max_nbr, slab_axis, ghost_lo, ghost_hi);
    // ---- boundary lists (nb == 0: all empty) ------------------------------------------------------------------------------
    const unsigned staged_b = (nb == 0u || !bpart_near[blockIdx.x]) ? 0u : tile_stage(g, tg, bpos, bcs, srec, rs, tr, nullptr, t0);
    // (the own-row table rowq is a by-product of the fluid staging: if that did not fit, the boundary pass cannot walk the tile either)
    tile_pass<false>(staged_b == 0u ? 0 : (staged_b <= DFSPH_TB_CAP && staged <= DFSPH_TB_CAP && fits ? 1 : 2), t0, t1, n, finish_tiles != 0, g, tg, R2, pos, bpos, bcs, srec_a, rs_a, rowq,
                     tab_b, Kb, cnt_b, tcnt_b, nb, &ctrl->overflow_b, nullptr, -1, 0.0, 0.0);
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
// FLUID_PASS = false: the fluid and boundary lists and their raw lengths (cnt_f, cnt_b) are already there (k_build_tiles); this
// kernel appends the ghosts (multi-GPU), pads the tiles and reports overflows.
template <bool FLUID_PASS>
__global__ void __launch_bounds__(DFSPH_BLOCK, DFSPH_BUILD_MIN_BLOCKS) k_build_neighbors(unsigned n, GridDesc g, Real R2,
    const Real4* __restrict__ pos, const unsigned* __restrict__ cell_start,
    const Real4* __restrict__ bpos, const unsigned* __restrict__ bcell_start, unsigned nb, const unsigned char* __restrict__ bnear,
    unsigned* __restrict__ tab_f, unsigned Kf, unsigned* __restrict__ tab_b, unsigned Kb,
    unsigned* __restrict__ cnt_f, unsigned* __restrict__ cnt_b, unsigned* __restrict__ tcnt_f, unsigned* __restrict__ tcnt_b, Ctrl* ctrl,
    unsigned ng, const unsigned* __restrict__ gcell_start, const unsigned* __restrict__ gperm, int slab_axis, double ghost_lo, double ghost_hi,
    const unsigned* __restrict__ ghost_block_rank)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned tile = i >> 5, lane = i & 31u;
    if (!FLUID_PASS && tcnt_f[tile] != TB_INCOMPLETE) return;   // finished by k_build_tiles
    const Real4 xi = ld_gather(pos + i);
    unsigned cf = FLUID_PASS ? search_cells<true, false>(xi, i, g, R2, pos, cell_start, tab_f, Kf, tile, lane) : cnt_f[i];
    // multi-GPU: ghost particles of the neighbouring slabs live behind the owned ones at pos[n .. n+ng)
    if (ng > 0) {
        const double a = slab_axis == 0 ? (double)xi.x : (slab_axis == 1 ? (double)xi.y : (double)xi.z);
        if (a < ghost_lo || a > ghost_hi) {
            // the ghost set has its own, compact cell table: same geometry, but only the blocks in ghost reach of the slab faces
            // have rows of their own (every other block shares one empty block), see build of ghost_block_rank
            GridDesc gg = g;
            gg.block_rank = ghost_block_rank;
            cf = search_cells<false, true>(xi, i, gg, R2, pos + n, gcell_start, tab_f, Kf, tile, lane, cf, gperm, n);
        }
    }
    unsigned cb = 0;
    if (!FLUID_PASS) cb = nb > 0 ? cnt_b[i] : 0u;
    else if (nb > 0) {
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

