// Microbenchmark: can the texture path take half of the scattered float4 gathers off the LSU data pipe on sm_100a?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(const float4* __restrict__ a, const float4* __restrict__ b, cudaTextureObject_t ta, cudaTextureObject_t tb,
                                         const unsigned* __restrict__ idx, int iters, float* out)
{
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const unsigned* my = idx + (size_t)warp * iters * 32 + lane;
    float acc = 0.f;
#pragma unroll 4
    for (int k = 0; k < iters; ++k) {
        const unsigned j = __ldg(my + (size_t)k * 32);
        float4 u, v;
        if (MODE == 0) { u = __ldg(a + j); v = __ldg(b + j); }                                   // both through LSU
        if (MODE == 1) { u = __ldg(a + j); v = tex1Dfetch<float4>(tb, (int)j); }                 // one LSU, one TEX
        if (MODE == 2) { u = tex1Dfetch<float4>(ta, (int)j); v = tex1Dfetch<float4>(tb, (int)j); } // both TEX
        if (MODE == 3) { u = __ldg(a + j); v = make_float4(0, 0, 0, 0); }                        // single LSU gather
        if (MODE == 4) { u = tex1Dfetch<float4>(ta, (int)j); v = make_float4(0, 0, 0, 0); }      // single TEX gather
        if (MODE == 5) { if (k & 1) u = tex1Dfetch<float4>(ta, (int)j); else u = __ldg(a + j); v = make_float4(0, 0, 0, 0); }  // alternate
        acc += u.x + u.y + u.z + u.w + v.x + v.y + v.z + v.w;
    }
    if (acc == 123.456f) out[0] = acc;
}

static cudaTextureObject_t make_tex(float4* p, size_t n)
{
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = p;
    rd.res.linear.desc = cudaCreateChannelDesc<float4>(); rd.res.linear.sizeInBytes = n * sizeof(float4);
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t t = 0; cudaCreateTextureObject(&t, &rd, &td, nullptr); return t;
}

int main()
{
    const int nwarps = 148 * 64 * 4, iters = 64;
    const size_t nidx = (size_t)nwarps * iters * 32;
    const unsigned ndata = 1u << 22;
    float4 *a, *b; unsigned* idx; float* out;
    cudaMalloc(&a, (size_t)ndata * 16); cudaMalloc(&b, (size_t)ndata * 16); cudaMemset(a, 0, (size_t)ndata * 16); cudaMemset(b, 0, (size_t)ndata * 16);
    cudaMalloc(&idx, nidx * 4); cudaMalloc(&out, 4);
    cudaTextureObject_t ta = make_tex(a, ndata), tb = make_tex(b, ndata);
    std::vector<unsigned> h(nidx);
    const int windows[] = {64, 384};
    for (int wi = 0; wi < 2; ++wi) {
        srand(1);
        for (int w = 0; w < nwarps; ++w) {
            const unsigned base = ((unsigned)w * 256u) % (ndata - 1024);
            for (int k = 0; k < iters; ++k) for (int l = 0; l < 32; ++l) h[((size_t)w * iters + k) * 32 + l] = base + rand() % windows[wi];
        }
        cudaMemcpy(idx, h.data(), nidx * 4, cudaMemcpyHostToDevice);
        const char* names[] = {"2x LDG.128", "LDG.128 + TEX", "2x TEX", "1x LDG.128", "1x TEX", "alternate LDG/TEX (1 gather)"};
        for (int mode = 0; mode < 6; ++mode) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            float best = 1e9f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                const int blocks = nwarps * 32 / 256;
                if (mode == 0) k<0><<<blocks, 256>>>(a, b, ta, tb, idx, iters, out);
                if (mode == 1) k<1><<<blocks, 256>>>(a, b, ta, tb, idx, iters, out);
                if (mode == 2) k<2><<<blocks, 256>>>(a, b, ta, tb, idx, iters, out);
                if (mode == 3) k<3><<<blocks, 256>>>(a, b, ta, tb, idx, iters, out);
                if (mode == 4) k<4><<<blocks, 256>>>(a, b, ta, tb, idx, iters, out);
                if (mode == 5) k<5><<<blocks, 256>>>(a, b, ta, tb, idx, iters, out);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            const double req_per_sm = (double)nwarps * iters / 148.0;
            printf("window %3d  %-30s %.3f ms -> %.2f SM-cycles per iteration\n", windows[wi], names[mode], best, best * 1e-3 * 1.9e9 / req_per_sm);
        }
    }
    return 0;
}
