// ubench_l2_partial.cu -- does data written with 8-byte partial-sector stores by one kernel hit in L2 when the next
// kernel reads it?  Measures the latency of a batch of 8 independent coalesced 256-byte warp loads (the access pattern
// of the NMS fold workers) after (a) strided 8-byte writes by many CTAs, (b) fully coalesced writes, (c) a previous read.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_l2_partial tools/ubench_l2_partial.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
constexpr int ROWS = 6000, CB = 94;

__global__ void write_partial(u64* m) {            // grid (CB, ROWS/64): thread = row, block column = word -> 8 B at stride 752 B
    const int row = blockIdx.y * 64 + threadIdx.x, col = blockIdx.x;
    if (row < ROWS) m[(size_t)row * CB + col] = (u64)row * 1315423911ULL + col;
}
__global__ void write_full(u64* m) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t)ROWS * CB) m[i] = i * 2654435761ULL;
}
__global__ void read_probe(const u64* __restrict__ m, long long* out, int row0) {
    const int lane = threadIdx.x;
    long long total = 0; u64 sink = 0;
    for (int it = 0; it < 16; ++it) {
        const u64* base = m + (size_t)(row0 + it * 64) * CB + 8 + lane;
        const long long t0 = clock64();
        u64 v[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = __ldcg(base + (size_t)(t * 3) * CB);
#pragma unroll
        for (int t = 0; t < 8; ++t) sink |= v[t];
        if (sink == 0x1234567ULL) out[2] = 1;
        total += clock64() - t0;
    }
    if (lane == 0) { out[0] = total / 16; out[1] = (long long)sink; }
}
__global__ void prefetch_rows(const u64* m, int row0) {      // one warp: prefetch.global.L2 of every line read_probe will touch
    const int lane = threadIdx.x;
    for (int it = 0; it < 16; ++it)
        for (int t = 0; t < 8; ++t) {
            const u64* p = m + (size_t)(row0 + it * 64 + t * 3) * CB + 8 + lane;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        }
}
int main() {
    u64* m; long long* out; long long h[3];
    cudaMalloc(&m, (size_t)ROWS * CB * 8); cudaMalloc(&out, 64);
    const char* names[] = {"after partial-sector writes", "after full coalesced writes", "after a previous read (same rows)", "after partial writes, other rows", "after partial writes + prefetch.L2 kernel"};
    for (int rep = 0; rep < 2; ++rep)
        for (int mode = 0; mode < 5; ++mode) {
            if (mode == 0 || mode == 3 || mode == 4) write_partial<<<dim3(CB, (ROWS + 63) / 64), 64>>>(m);
            if (mode == 1) write_full<<<(ROWS * CB + 255) / 256, 256>>>(m);
            if (mode == 2) read_probe<<<1, 32>>>(m, out, 100);
            if (mode == 4) prefetch_rows<<<1, 32>>>(m, 100);
            read_probe<<<1, 32>>>(m, out, mode == 3 ? 3000 : 100);
            cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
            printf("%-40s batch of 8 loads: %lld cycles\n", names[mode], h[0]);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
