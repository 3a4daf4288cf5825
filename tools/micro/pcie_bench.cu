// Raw pinned H2D / D2H bandwidth of the box (context for bench.py's e2e number).
#include <cstdio>
#include <cuda_runtime.h>
int main()
{
    const size_t bytes = 240ull << 20;
    void *h, *d; cudaMallocHost(&h, bytes); cudaMalloc(&d, bytes);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int dir = 0; dir < 2; ++dir) {
        float best = 1e9f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(a, s);
            if (dir == 0) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s); else cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s);
            cudaEventRecord(b, s); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        printf("%s 240 MiB pinned: %.2f ms = %.1f GB/s\n", dir == 0 ? "H2D" : "D2H", best, bytes / best / 1e6);
    }
    // both directions at once on two streams
    cudaStream_t s2; cudaStreamCreate(&s2); void *h2, *d2; cudaMallocHost(&h2, bytes); cudaMalloc(&d2, bytes);
    cudaEventRecord(a, s); cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s); cudaMemcpyAsync(h2, d2, bytes, cudaMemcpyDeviceToHost, s2);
    cudaStreamSynchronize(s2); cudaEventRecord(b, s); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); printf("H2D + D2H concurrently: %.2f ms = %.1f GB/s aggregate\n", ms, 2.0 * bytes / ms / 1e6);
    return 0;
}
