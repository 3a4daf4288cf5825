// Multi-GPU slab decomposition along one coordinate axis (SURVEY.md row e): one context per GPU / process, ghost particles of one
// support radius appended behind the owned particles, refreshed over NCCL send/recv; particle migration each step;
// the density-error sum and the CFL maximum are all-reduced so that every rank takes identical loop decisions.
// NCCL is resolved at run time with dlopen (the process usually already holds torch's libnccl.so.2).
#pragma once
#include "common.cuh"
#include <nccl.h>
#include <dlfcn.h>
#include <string>

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;

    bool load(std::string& err)
    {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define NCCL_SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(lib, name)); if (!field) { err = std::string("libnccl lacks ") + name; return false; }
        NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        NCCL_SYM(CommInitRank, "ncclCommInitRank")
        NCCL_SYM(CommDestroy, "ncclCommDestroy")
        NCCL_SYM(Send, "ncclSend")
        NCCL_SYM(Recv, "ncclRecv")
        NCCL_SYM(GroupStart, "ncclGroupStart")
        NCCL_SYM(GroupEnd, "ncclGroupEnd")
        NCCL_SYM(AllReduce, "ncclAllReduce")
        NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
        return true;
    }
};

// device-side counters of the exchange, read by the host once per phase
struct ExchangeCounts {
    unsigned leave_l, leave_r;     // owned particles that crossed the left / right slab face
    unsigned exp_l, exp_r;         // owned particles within one cell of the left / right face (ghost exports)
    unsigned exp_all, pad0, pad1, pad2;   // union of the two export sets (processed first by the solver passes)
};

// coordinate along the slab axis (0 = x, 1 = y, 2 = z)
__device__ __forceinline__ double slab_coord(const Real4& p, int axis) { return (double)(axis == 0 ? p.x : (axis == 1 ? p.y : p.z)); }

// ---- migration --------------------------------------------------------------------------------------------------
struct MigrantAux { Real kappa, kappa_v; unsigned id, state; };

// Owned particles outside [lo, hi) are packed for the neighbour rank (atomic compaction).  They are dropped from the
// owned set by the cell sort, which files them under the dump key `num_keys` (see k_cell_hash_slab).
__global__ void __launch_bounds__(DFSPH_BLOCK) k_pack_leavers(unsigned n, const Real4* __restrict__ pos, const Real4* __restrict__ vel,
    const Real* __restrict__ kappa, const Real* __restrict__ kappa_v, const unsigned* __restrict__ id, const unsigned* __restrict__ state,
    double lo, double hi, int axis, int has_left, int has_right, unsigned cap,
    Real4* __restrict__ pl, Real4* __restrict__ vl, MigrantAux* __restrict__ al,
    Real4* __restrict__ pr, Real4* __restrict__ vr, MigrantAux* __restrict__ ar, ExchangeCounts* cnt)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Real4 p = pos[i];
    const double x = slab_coord(p, axis);
    if (has_left && x < lo) {
        const unsigned k = atomicAdd(&cnt->leave_l, 1u);
        if (k < cap) { pl[k] = p; vl[k] = vel[i]; al[k] = MigrantAux{kappa[i], kappa_v[i], id[i], state[i]}; }
    } else if (has_right && x >= hi) {
        const unsigned k = atomicAdd(&cnt->leave_r, 1u);
        if (k < cap) { pr[k] = p; vr[k] = vel[i]; ar[k] = MigrantAux{kappa[i], kappa_v[i], id[i], state[i]}; }
    }
}

__global__ void __launch_bounds__(DFSPH_BLOCK) k_unpack_arrivals(unsigned count, unsigned base, const MigrantAux* __restrict__ aux,
    Real* __restrict__ kappa, Real* __restrict__ kappa_v, unsigned* __restrict__ id, unsigned* __restrict__ state)
{
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const MigrantAux a = aux[k];
    kappa[base + k] = a.kappa; kappa_v[base + k] = a.kappa_v; id[base + k] = a.id; state[base + k] = a.state;
}

// cell hash with the slab filter: leavers get the dump key (sorted behind every real cell, then cut off)
__global__ void __launch_bounds__(DFSPH_BLOCK) k_cell_hash_slab(const Real4* __restrict__ pos, unsigned n, GridDesc g, double lo, double hi, int axis,
    int has_left, int has_right, unsigned* __restrict__ cell_count, unsigned* __restrict__ key_out, unsigned* __restrict__ rank_out,
    unsigned* __restrict__ fine_out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Real4 p = pos[i];
    unsigned xord;
    unsigned key = position_entry(p, g, xord);
    const double x = slab_coord(p, axis);
    if ((has_left && x < lo) || (has_right && x >= hi)) key = g.num_keys;
    key_out[i] = key;
    fine_out[i] = xord;
    rank_out[i] = atomicAdd(cell_count + key, 1u);
}

// ---- ghost exports -------------------------------------------------------------------------------------------------
// Owned (sorted) particles within `width` of a slab face are exported to that neighbour.
__global__ void __launch_bounds__(DFSPH_BLOCK) k_select_exports(unsigned n, const Real4* __restrict__ pos, double lo, double hi, double width, int axis,
    int has_left, int has_right, unsigned cap, unsigned* __restrict__ exp_l, unsigned* __restrict__ exp_r,
    unsigned* __restrict__ exp_all, unsigned char* __restrict__ is_export, ExchangeCounts* cnt)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = slab_coord(pos[i], axis);
    const bool l = has_left && x < lo + width, r = has_right && x >= hi - width;
    if (l) { const unsigned k = atomicAdd(&cnt->exp_l, 1u); if (k < cap) exp_l[k] = i; }
    if (r) { const unsigned k = atomicAdd(&cnt->exp_r, 1u); if (k < cap) exp_r[k] = i; }
    if (l || r) { const unsigned k = atomicAdd(&cnt->exp_all, 1u); if (k < 2u * cap) exp_all[k] = i; }
    is_export[i] = (l || r) ? 1 : 0;
}

__global__ void __launch_bounds__(DFSPH_BLOCK) k_copy_real4(const Real4* __restrict__ src, Real4* __restrict__ dst, unsigned n)
{
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) st_real4(dst + k, ld_plain(src + k));
}

// gather one Real4 field of the export lists into the two send buffers
__global__ void __launch_bounds__(DFSPH_BLOCK) k_pack_exports(const Real4* __restrict__ src, const unsigned* __restrict__ exp_l, unsigned nl,
    const unsigned* __restrict__ exp_r, unsigned nr, Real4* __restrict__ out_l, Real4* __restrict__ out_r)
{
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nl) st_real4(out_l + k, ld_plain(src + exp_l[k]));
    else if (k < nl + nr) st_real4(out_r + (k - nl), ld_plain(src + exp_r[k - nl]));
}

__global__ void k_write_sentinel(Real4* pos, Real4* vel, Real4* acc, unsigned at)
{
    st_real4(pos + at, make_real4((Real)1.0e15, (Real)1.0e15, (Real)1.0e15, (Real)0.0));
    st_real4(vel + at, make_real4((Real)0.0, (Real)0.0, (Real)0.0, (Real)0.0));
    st_real4(acc + at, make_real4((Real)0.0, (Real)0.0, (Real)0.0, (Real)0.0));
}

__global__ void k_zero_counts(ExchangeCounts* c) { c->leave_l = c->leave_r = c->exp_l = c->exp_r = c->exp_all = 0u; }

// ---- NVLink peer-to-peer ghost refresh ---------------------------------------------------------------------------------
// The owning rank stores its export values straight into the neighbour's ghost slots (peer memory mapped through CUDA
// IPC) and then publishes a sequence number in the neighbour's flag word; the consumer spins on its local flag.  No
// host round trip, no NCCL launch: ~10 us per refresh instead of ~40 us (pack kernel + grouped ncclSend/ncclRecv).
// Slot safety: a producer can only be one refresh ahead of its consumer because every solver iteration ends in an
// all-reduce (see DESIGN.md "Multi-GPU").
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// What a push needs to know about this step, kept in device memory so that the launches can live in a CUDA graph: export
// counts, the element offsets of this rank's slots in the neighbours' ghost regions, and the owned particle count (the sweep
// kernels of the slab loop graphs read it instead of a kernel argument, because migration changes it every step).
struct PushDesc { unsigned nl, nr; unsigned long long off_l, off_r; unsigned n_local, pad; };

__global__ void k_set_push_desc(PushDesc* pd, unsigned nl, unsigned nr, unsigned long long off_l, unsigned long long off_r, unsigned n_local)
{
    pd->nl = nl; pd->nr = nr; pd->off_l = off_l; pd->off_r = off_r; pd->n_local = n_local; pd->pad = 0u;
}

// Sequence numbers live on the device (`seq_word`, one per kind of refresh, next to the flag words): the last block of a push
// advances the word and publishes the new value in the neighbours' flags; a wait reads the word -- it runs behind the local
// push of the same kind in stream order, and all ranks push in lockstep, so it expects exactly the neighbours' matching push.
// No host-side counter: pushes and waits can be nodes of a graph (the solver loops on slabs).
__global__ void __launch_bounds__(DFSPH_BLOCK) k_push_exports(const Real4* __restrict__ src,
    const unsigned* __restrict__ exp_l, Real4* __restrict__ dst_l_base, unsigned* flag_l,
    const unsigned* __restrict__ exp_r, Real4* __restrict__ dst_r_base, unsigned* flag_r,
    const PushDesc* __restrict__ pd, unsigned* seq_word, unsigned* ticket)
{
    const unsigned nl = pd->nl, tot = nl + pd->nr;
    Real4* dst_l = dst_l_base + pd->off_l;
    Real4* dst_r = dst_r_base + pd->off_r;
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < tot; k += gridDim.x * blockDim.x) {
        if (k < nl) st_real4(dst_l + k, ld_plain(src + exp_l[k]));
        else st_real4(dst_r + (k - nl), ld_plain(src + exp_r[k - nl]));
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {          // every block's stores are fenced before its ticket: all data is visible
            *ticket = 0u;
            const unsigned seq = *seq_word + 1u;
            *seq_word = seq;
            __threadfence_system();
            if (flag_l) st_release_sys(flag_l, seq);
            if (flag_r) st_release_sys(flag_r, seq);
        }
    }
}

__global__ void k_wait_flags(const unsigned* f_left, const unsigned* f_right, const unsigned* seq_word)
{
    const unsigned seq = *seq_word;
    if (f_left) while ((int)(ld_acquire_sys(f_left) - seq) < 0) { }
    if (f_right) while ((int)(ld_acquire_sys(f_right) - seq) < 0) { }
    __threadfence_system();
}
