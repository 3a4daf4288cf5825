// Microbenchmark: issue rate of packed fp32 (FFMA2 / fma.rn.f32x2) against scalar FFMA on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s)
{
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    const float2 m = make_float2(s, s * 0.5f), c = make_float2(0.25f, -0.125f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
            else a[i] = __ffma2_rn(a[i], m, c);
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
    if (r == 123.456f) out[0] = r;
}
int main()
{
    float* out; cudaMalloc(&out, 4);
    const int iters = 4096, blocks = 148 * 8;
    for (int mode = 0; mode < 2; ++mode) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<blocks, 256>>>(out, iters, 0.999f); else k<1><<<blocks, 256>>>(out, iters, 0.999f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double fma = (double)blocks * 256 * iters * 16;
        printf("%s: %.3f ms, %.1f Gfma/s (%.1f TFLOP/s)\n", mode == 0 ? "scalar FFMA" : "packed FFMA2", best, fma / best * 1e-6, 2 * fma / best * 1e-9);
    }
    return 0;
}
