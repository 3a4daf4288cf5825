// Microbenchmark: cost of per-lane scattered gathers through L1 on sm_100a, by access width and address pattern.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k_gather(const float4* __restrict__ data, const unsigned* __restrict__ idx, int iters, float* out)
{
    // each warp owns a window of `idx` entries: idx[(warp*iters + k)*32 + lane]
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const unsigned* my = idx + (size_t)warp * iters * 32 + lane;
    float acc = 0.f;
#pragma unroll 4
    for (int k = 0; k < iters; ++k) {
        const unsigned j = __ldg(my + (size_t)k * 32);
        if (MODE == 0) { const float4 v = __ldg(data + j); acc += v.x + v.y + v.z + v.w; }                       // LDG.128
        if (MODE == 1) { const float4 a = __ldg(data + 2 * j); const float4 b = __ldg(data + 2 * j + 1); acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w; }  // 2x LDG.128 same 32B record
        if (MODE == 2) { float a0,a1,a2,a3,a4,a5,a6,a7; asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a0),"=f"(a1),"=f"(a2),"=f"(a3),"=f"(a4),"=f"(a5),"=f"(a6),"=f"(a7) : "l"(data + 2 * j)); acc += a0+a1+a2+a3+a4+a5+a6+a7; }  // LDG.256
        if (MODE == 3) { const float2 v = __ldg(reinterpret_cast<const float2*>(data) + 2 * j); acc += v.x + v.y; }   // LDG.64 (first 8 B of a 16 B record)
        if (MODE == 4) { const float v = __ldg(reinterpret_cast<const float*>(data) + 4 * j); acc += v; }              // LDG.32
    }
    if (acc == 123.456f) out[0] = acc;
}

int main()
{
    const int nwarps = 148 * 64 * 4, iters = 64;
    const size_t nidx = (size_t)nwarps * iters * 32;
    const unsigned ndata = 1u << 22;   // 4M records
    float4* data; unsigned* idx; float* out;
    cudaMalloc(&data, (size_t)ndata * 32); cudaMemset(data, 0, (size_t)ndata * 32);
    cudaMalloc(&idx, nidx * 4); cudaMalloc(&out, 4);
    std::vector<unsigned> h(nidx);
    const char* pat_name[] = {"coalesced (lane l -> base+l)", "random in 384-record window (SPH-like)", "pairs share a 32B sector, random pairs in window",
                              "8-lane groups in one 128B line, lines random in window", "random in 64-record window"};
    for (int pat = 0; pat < 5; ++pat) {
        srand(1);
        for (int w = 0; w < nwarps; ++w) {
            const unsigned base = ((unsigned)w * 256u) % (ndata / 2 - 1024);
            for (int k = 0; k < iters; ++k)
                for (int l = 0; l < 32; ++l) {
                    unsigned j;
                    if (pat == 0) j = base + ((k * 32 + l) % 384);
                    else if (pat == 1) j = base + rand() % 384;
                    else if (pat == 2) { static unsigned p; if ((l & 1) == 0) p = (rand() % 192) * 2; j = base + p + (l & 1); }
                    else if (pat == 3) { static unsigned p; if ((l & 7) == 0) p = (rand() % 48) * 8; j = base + p + (l & 7); }
                    else j = base + rand() % 64;
                    h[((size_t)w * iters + k) * 32 + l] = j;
                }
        }
        cudaMemcpy(idx, h.data(), nidx * 4, cudaMemcpyHostToDevice);
        const char* mode_name[] = {"LDG.128 (16B rec)", "2xLDG.128 (32B rec)", "LDG.256 (32B rec)", "LDG.64", "LDG.32"};
        for (int mode = 0; mode < 5; ++mode) {
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            float best = 1e9f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(a);
                const int blocks = nwarps * 32 / 256;
                if (mode == 0) k_gather<0><<<blocks, 256>>>(data, idx, iters, out);
                if (mode == 1) k_gather<1><<<blocks, 256>>>(data, idx, iters, out);
                if (mode == 2) k_gather<2><<<blocks, 256>>>(data, idx, iters, out);
                if (mode == 3) k_gather<3><<<blocks, 256>>>(data, idx, iters, out);
                if (mode == 4) k_gather<4><<<blocks, 256>>>(data, idx, iters, out);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
            }
            // cycles per gather request per SM at ~1.9 GHz
            const double req_per_sm = (double)nwarps * iters / 148.0;
            printf("pattern[%s] %-22s %.3f ms  -> %.2f SM-cycles per warp-request (@1.9GHz)\n", pat_name[pat], mode_name[mode], best, best * 1e-3 * 1.9e9 / req_per_sm);
        }
    }
    return 0;
}
