// Microbenchmark for a possible round-2 redesign of the DFSPH sweeps: per-neighbour gathers from GLOBAL memory through
// L1 (the current design: one scattered LDG.128 per pair) against gathers from a SHARED-MEMORY copy of the CTA's
// spatial block + halo (one LDS.128 per pair, 16-bit block-local indices).  Synthetic jittered lattice, 8 particles per
// cell of edge R, block = 4x4x4 cells (512 particles, one CTA), halo = 6x6x6 cells (1728 records = 27 KB).
// The inner loop is pass A of the solver (a_i -= V (k_i + k_j) gradW_ij) so the instruction mix is realistic.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int NB = 16;            // blocks per axis
constexpr int BC = 4;             // cells per block axis
constexpr int NC = NB * BC;       // cells per axis
constexpr int PPC = 8;            // particles per cell
constexpr int BP = BC * BC * BC * PPC;        // 512 particles per block
constexpr int HC = BC + 2;                    // halo cells per axis
constexpr int HP = HC * HC * HC * PPC;        // 1728 halo records
constexpr int K = 64;             // table capacity per particle
constexpr float R = 0.1f;

__host__ __device__ inline unsigned cell_base(int cx, int cy, int cz)   // first particle of a cell in block-major order
{
    const int bx = cx / BC, by = cy / BC, bz = cz / BC, lx = cx % BC, ly = cy % BC, lz = cz % BC;
    const unsigned b = (unsigned)((bz * NB + by) * NB + bx);
    const unsigned l = (unsigned)((lz * BC + ly) * BC + lx);
    return (b * (BC * BC * BC) + l) * PPC;
}

// one thread per particle: neighbour lists (global indices and block-local halo indices), warp-tile interleaved
__global__ void k_build(const float4* __restrict__ pos, unsigned n, unsigned* __restrict__ tab_g, unsigned short* __restrict__ tab_l,
                        unsigned* __restrict__ tcnt)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned b = i / BP, li = i % BP;
    const int bx = b % NB, by = (b / NB) % NB, bz = b / (NB * NB);
    const int lc = li / PPC;
    const int cx = bx * BC + lc % BC, cy = by * BC + (lc / BC) % BC, cz = bz * BC + lc / (BC * BC);
    const float4 pi = pos[i];
    unsigned cnt = 0;
    unsigned* tg = tab_g + (size_t)(i >> 5) * K * 32 + (i & 31);
    unsigned short* tl = tab_l + (size_t)(i >> 5) * K * 32 + (i & 31);
    for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
        const int x = cx + dx, y = cy + dy, z = cz + dz;
        if (x < 0 || y < 0 || z < 0 || x >= NC || y >= NC || z >= NC) continue;
        const unsigned base = cell_base(x, y, z);
        const int hx = x - (bx * BC - 1), hy = y - (by * BC - 1), hz = z - (bz * BC - 1);
        const unsigned hbase = (unsigned)((hz * HC + hy) * HC + hx) * PPC;
        for (int p = 0; p < PPC; ++p) {
            const unsigned j = base + p;
            if (j == i) continue;
            const float4 pj = pos[j];
            const float ddx = pi.x - pj.x, ddy = pi.y - pj.y, ddz = pi.z - pj.z;
            if (ddx * ddx + ddy * ddy + ddz * ddz < R * R && cnt < K) { tg[(size_t)cnt * 32] = j; tl[(size_t)cnt * 32] = (unsigned short)(hbase + p); ++cnt; }
        }
    }
    const unsigned mx = (__reduce_max_sync(0xffffffffu, cnt) + 3u) & ~3u;
    for (unsigned k = cnt; k < mx; ++k) { tg[(size_t)k * 32] = n; tl[(size_t)k * 32] = (unsigned short)HP; }
    if ((i & 31) == 0) tcnt[i >> 5] = mx;
}

__device__ __forceinline__ void pair(const float4& pi, const float4& pj, float& ax, float& ay, float& az)
{
    const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
    const float r2 = dx * dx + dy * dy + dz * dz;
    float inv; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(r2));
    const float q = r2 * inv * (1.0f / R), v = 1.0f - q;
    const float g1 = 3.0f * q - 2.0f, g2 = -inv * v * v;
    float g = q <= 1.0f ? g2 : 0.0f; g = q <= 0.5f ? g1 : g; g = r2 > 1e-18f ? g : 0.0f;
    const float s = (pi.w + pj.w) * g;
    ax -= s * dx; ay -= s * dy; az -= s * dz;
}

// current design: scattered LDG.128 per pair
__global__ void __launch_bounds__(256, 8) k_global(const float4* __restrict__ pos, const unsigned* __restrict__ tab, const unsigned* __restrict__ tcnt,
                                                   float4* __restrict__ out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const float4 pi = pos[i];
    const unsigned m = tcnt[i >> 5];
    const unsigned* t = tab + (size_t)(i >> 5) * K * 32 + (i & 31);
    float ax = 0, ay = 0, az = 0;
    for (unsigned k = 0; k < m; k += 4) {
        const unsigned j0 = t[(size_t)k * 32], j1 = t[(size_t)(k + 1) * 32], j2 = t[(size_t)(k + 2) * 32], j3 = t[(size_t)(k + 3) * 32];
        const float4 p0 = __ldg(pos + j0), p1 = __ldg(pos + j1), p2 = __ldg(pos + j2), p3 = __ldg(pos + j3);
        pair(pi, p0, ax, ay, az); pair(pi, p1, ax, ay, az); pair(pi, p2, ax, ay, az); pair(pi, p3, ax, ay, az);
    }
    out[i] = make_float4(ax, ay, az, 0.0f);
}

// variant: S lanes per particle (lane s of a particle takes the list entries k = s mod S), so that the lanes of one
// request fetch S CONSECUTIVE list entries of 32/S particles (consecutive entries are usually neighbours in memory)
template <int S>
__global__ void k_build_split(const unsigned* __restrict__ tab_g, const unsigned* __restrict__ tcnt, unsigned n, unsigned* __restrict__ tab_s,
                              unsigned* __restrict__ tcnt_s)
{
    // re-layout of the S=1 table: tile = 32/S particles; row r of a tile holds entries k = r*S + s of particle p at lane p*S + s
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;   // particle
    constexpr unsigned P = 32 / S;
    const unsigned* tg = tab_g + (size_t)(i >> 5) * K * 32 + (i & 31);
    unsigned cnt = 0;
    for (unsigned k = 0; k < tcnt[i >> 5]; ++k) if (tg[(size_t)k * 32] != n) ++cnt;
    unsigned rows = (cnt + S - 1) / S;
    // max over the P particles of the tile (P consecutive lanes of this warp)
    for (unsigned o = 1; o < P; o <<= 1) rows = max(rows, __shfl_xor_sync(0xffffffffu, rows, o));
    rows = (rows + 3u) & ~3u;
    unsigned* ts = tab_s + (size_t)(i / P) * K * 32 + (i % P) * S;
    for (unsigned k = 0; k < rows * S; ++k) ts[(size_t)(k / S) * 32 + (k % S)] = k < cnt ? tg[(size_t)k * 32] : n;
    if (i % P == 0) tcnt_s[i / P] = rows;
}

template <int S>
__global__ void __launch_bounds__(256, 8) k_global_split(const float4* __restrict__ pos, const unsigned* __restrict__ tab, const unsigned* __restrict__ tcnt,
                                                         float4* __restrict__ out)
{
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;   // lane-slot
    const unsigned i = t / S;
    const float4 pi = pos[i];
    const unsigned m = tcnt[t >> 5];
    const unsigned* tb = tab + (size_t)(t >> 5) * K * 32 + (t & 31);
    float ax = 0, ay = 0, az = 0;
    for (unsigned k = 0; k < m; k += 4) {
        const unsigned j0 = tb[(size_t)k * 32], j1 = tb[(size_t)(k + 1) * 32], j2 = tb[(size_t)(k + 2) * 32], j3 = tb[(size_t)(k + 3) * 32];
        const float4 p0 = __ldg(pos + j0), p1 = __ldg(pos + j1), p2 = __ldg(pos + j2), p3 = __ldg(pos + j3);
        pair(pi, p0, ax, ay, az); pair(pi, p1, ax, ay, az); pair(pi, p2, ax, ay, az); pair(pi, p3, ax, ay, az);
    }
    for (unsigned o = 1; o < S; o <<= 1) { ax += __shfl_xor_sync(0xffffffffu, ax, o); ay += __shfl_xor_sync(0xffffffffu, ay, o); az += __shfl_xor_sync(0xffffffffu, az, o); }
    if (t % S == 0) out[i] = make_float4(ax, ay, az, 0.0f);
}

// candidate design: CTA = spatial block, halo staged in shared memory, LDS.128 per pair
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_shared(const float4* __restrict__ pos, const unsigned short* __restrict__ tab, const unsigned* __restrict__ tcnt,
                                                    float4* __restrict__ out, unsigned n)
{
    __shared__ float4 sh[HP + 1];
    const unsigned b = blockIdx.x;
    const int bx = b % NB, by = (b / NB) % NB, bz = b / (NB * NB);
    for (unsigned h = threadIdx.x; h < HP; h += THREADS) {
        const unsigned hc = h / PPC, p = h % PPC;
        const int x = bx * BC - 1 + (int)(hc % HC), y = by * BC - 1 + (int)((hc / HC) % HC), z = bz * BC - 1 + (int)(hc / (HC * HC));
        float4 v = make_float4(1e15f, 1e15f, 1e15f, 0.0f);
        if (x >= 0 && y >= 0 && z >= 0 && x < NC && y < NC && z < NC) v = __ldg(pos + cell_base(x, y, z) + p);
        sh[h] = v;
    }
    if (threadIdx.x == 0) sh[HP] = make_float4(1e15f, 1e15f, 1e15f, 0.0f);
    __syncthreads();
    for (unsigned li = threadIdx.x; li < BP; li += THREADS) {
        const unsigned i = b * BP + li;
        const unsigned lc = li / PPC;
        const unsigned hself = (((lc / (BC * BC)) + 1) * HC + ((lc / BC) % BC) + 1) * HC + (lc % BC) + 1;
        const float4 pi = sh[hself * PPC + li % PPC];
        const unsigned m = tcnt[i >> 5];
        const unsigned short* t = tab + (size_t)(i >> 5) * K * 32 + (i & 31);
        float ax = 0, ay = 0, az = 0;
        for (unsigned k = 0; k < m; k += 4) {
            const unsigned j0 = t[(size_t)k * 32], j1 = t[(size_t)(k + 1) * 32], j2 = t[(size_t)(k + 2) * 32], j3 = t[(size_t)(k + 3) * 32];
            const float4 p0 = sh[j0], p1 = sh[j1], p2 = sh[j2], p3 = sh[j3];
            pair(pi, p0, ax, ay, az); pair(pi, p1, ax, ay, az); pair(pi, p2, ax, ay, az); pair(pi, p3, ax, ay, az);
        }
        out[i] = make_float4(ax, ay, az, 0.0f);
    }
}

int main()
{
    const unsigned n = (unsigned)NC * NC * NC * PPC;
    std::vector<float4> h(n + 1);
    srand(7);
    for (int cz = 0; cz < NC; ++cz) for (int cy = 0; cy < NC; ++cy) for (int cx = 0; cx < NC; ++cx) {
        const unsigned base = cell_base(cx, cy, cz);
        for (int p = 0; p < PPC; ++p) {
            const float jx = (rand() / (float)RAND_MAX - 0.5f) * 0.3f, jy = (rand() / (float)RAND_MAX - 0.5f) * 0.3f, jz = (rand() / (float)RAND_MAX - 0.5f) * 0.3f;
            h[base + p] = make_float4((cx + 0.25f + 0.5f * (p & 1) + jx * 0.5f) * R, (cy + 0.25f + 0.5f * ((p >> 1) & 1) + jy * 0.5f) * R,
                                      (cz + 0.25f + 0.5f * ((p >> 2) & 1) + jz * 0.5f) * R, 1e-3f * (rand() / (float)RAND_MAX));
        }
    }
    h[n] = make_float4(1e15f, 1e15f, 1e15f, 0.0f);
    float4 *pos, *out_g, *out_s; unsigned* tab_g; unsigned short* tab_l; unsigned* tcnt;
    CK(cudaMalloc(&pos, (size_t)(n + 1) * 16)); CK(cudaMalloc(&out_g, (size_t)n * 16)); CK(cudaMalloc(&out_s, (size_t)n * 16));
    CK(cudaMalloc(&tab_g, (size_t)n * K * 4)); CK(cudaMalloc(&tab_l, (size_t)n * K * 2)); CK(cudaMalloc(&tcnt, (size_t)(n / 32) * 4));
    CK(cudaMemcpy(pos, h.data(), (size_t)(n + 1) * 16, cudaMemcpyHostToDevice));
    k_build<<<n / 256, 256>>>(pos, n, tab_g, tab_l, tcnt);
    CK(cudaDeviceSynchronize());
    std::vector<unsigned> hc(n / 32); CK(cudaMemcpy(hc.data(), tcnt, (size_t)(n / 32) * 4, cudaMemcpyDeviceToHost));
    double slots = 0; for (unsigned v : hc) slots += v;
    printf("particles %u, padded slots per particle %.1f\n", n, slots / (n / 32));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char* name, auto launch) {
        float best = 1e9f;
        for (int r = 0; r < 6; ++r) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms; }
        CK(cudaGetLastError());
        printf("%-44s %.3f ms  (%.2f ns per particle)\n", name, best, best * 1e6 / n);
    };
    timeit("global LDG.128 gathers (current design)", [&] { k_global<<<n / 256, 256>>>(pos, tab_g, tcnt, out_g); });
    timeit("shared-memory tile, 512 threads / block", [&] { k_shared<512><<<n / BP, 512>>>(pos, tab_l, tcnt, out_s, n); });
    timeit("shared-memory tile, 256 threads / block", [&] { k_shared<256><<<n / BP, 256>>>(pos, tab_l, tcnt, out_s, n); });
    {
        unsigned *tab_s, *tcnt_s; float4* out_p;
        CK(cudaMalloc(&tab_s, (size_t)n * K * 4 * 4)); CK(cudaMalloc(&tcnt_s, (size_t)(n / 8) * 4)); CK(cudaMalloc(&out_p, (size_t)n * 16));
        auto stats = [&](int S) { std::vector<unsigned> c(n / (32 / S)); CK(cudaMemcpy(c.data(), tcnt_s, c.size() * 4, cudaMemcpyDeviceToHost)); double t = 0; for (unsigned v : c) t += v; printf("  S=%d: rows per tile %.1f -> padded slots per particle %.1f\n", S, t / c.size(), t / c.size() * S); };
        k_build_split<2><<<n / 256, 256>>>(tab_g, tcnt, n, tab_s, tcnt_s); CK(cudaDeviceSynchronize()); stats(2);
        timeit("global LDG.128, 2 lanes per particle", [&] { k_global_split<2><<<n * 2 / 256, 256>>>(pos, tab_s, tcnt_s, out_p); });
        k_build_split<4><<<n / 256, 256>>>(tab_g, tcnt, n, tab_s, tcnt_s); CK(cudaDeviceSynchronize()); stats(4);
        timeit("global LDG.128, 4 lanes per particle", [&] { k_global_split<4><<<n * 4 / 256, 256>>>(pos, tab_s, tcnt_s, out_p); });
        std::vector<float4> a(n), b(n);
        CK(cudaMemcpy(a.data(), out_g, (size_t)n * 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), out_p, (size_t)n * 16, cudaMemcpyDeviceToHost));
        double md = 0, mx = 0; for (unsigned i = 0; i < n; ++i) { md = fmax(md, fabs(a[i].x - b[i].x) + fabs(a[i].y - b[i].y) + fabs(a[i].z - b[i].z)); mx = fmax(mx, fabs(a[i].x)); }
        printf("  max |S=1 - S=4| = %g (max |a.x| = %g)\n", md, mx);
    }
    std::vector<float4> a(n), b(n);
    CK(cudaMemcpy(a.data(), out_g, (size_t)n * 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), out_s, (size_t)n * 16, cudaMemcpyDeviceToHost));
    double md = 0; for (unsigned i = 0; i < n; ++i) md = fmax(md, fabs(a[i].x - b[i].x) + fabs(a[i].y - b[i].y) + fabs(a[i].z - b[i].z));
    printf("max |global - shared| = %g\n", md);
    return 0;
}
