// Micro-benchmark: throughput of fp32 global reductions (scalar / v2 / v4) and of shared-memory RMW,
// to choose the RoIAlign-backward accumulation strategy.   nvcc -arch=sm_100a -O3 -o ubench_red ubench_red.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int V>
__global__ void red_kernel(float* buf, size_t n_elems, int iters, unsigned seed) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        unsigned h = hash32(tid * 977u + it * 7919u + seed);
        size_t idx = ((size_t)h % (n_elems / V)) * V;
        float* p = buf + idx;
        if (V == 1) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.0f) : "memory");
        if (V == 2) asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(1.0f), "f"(2.0f) : "memory");
        if (V == 4) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.0f), "f"(2.0f), "f"(3.0f), "f"(4.0f) : "memory");
    }
}
// lanes = 32 consecutive channels of one random "cell" (coalesced 128 B red per warp), scalar vs v4 (8 lanes x 16 B)
template <int V>
__global__ void red_cell_kernel(float* buf, size_t n_cells, int iters, unsigned seed) {
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
        unsigned h = hash32(warp * 977u + it * 7919u + seed);
        if (V == 1) { float* p = buf + ((size_t)h % n_cells) * 32 + lane; asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.0f) : "memory"); }
        if (V == 4) { unsigned h2 = hash32(h + (lane >> 3)); float* p = buf + ((size_t)h2 % n_cells) * 32 + (lane & 7) * 4;
                      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.0f), "f"(2.0f), "f"(3.0f), "f"(4.0f) : "memory"); }
    }
}
// shared-memory RMW: LDS.128 + 4 FADD + STS.128 on conflict-free addresses (each quarter-warp its own 128 B cell)
__global__ void smem_rmw_kernel(float* out, int iters) {
    extern __shared__ float4 s[];
    const int n4 = 96 * 1024 / 16;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) s[i] = make_float4(0, 0, 0, 0);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, q = lane >> 3, i8 = lane & 7;
    unsigned h = hash32(blockIdx.x * 131u + warp);
    for (int it = 0; it < iters; ++it) {
        h = hash32(h + it);
        const int cell = (((h >> 2) % (n4 / 8 / 8)) * 8 + warp) ;      // warp-exclusive cells
        float4* p = s + (size_t)(cell * 4 + q) % (n4 / 8) * 8 + i8;
        float4 v = *p; v.x += 1.f; v.y += 2.f; v.z += 3.f; v.w += 4.f; *p = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s[0].x;
}
template <typename F> float time_ms(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    const size_t n = 13926400;           // one cfg2 dX (55.7 MB)
    float* buf; cudaMalloc(&buf, n * 4); cudaMemset(buf, 0, n * 4);
    const int blocks = 148 * 8, threads = 256, iters = 64;
    const double ops = (double)blocks * threads * iters;
    float ms;
    ms = time_ms([&] { red_kernel<1><<<blocks, threads>>>(buf, n, iters, 1); }); printf("random scalar red : %.1f Gop/s  (%.1f Gfloat/s) %.3f ms\n", ops / ms / 1e6, ops / ms / 1e6, ms);
    ms = time_ms([&] { red_kernel<2><<<blocks, threads>>>(buf, n, iters, 2); }); printf("random v2 red     : %.1f Gop/s  (%.1f Gfloat/s)\n", ops / ms / 1e6, 2 * ops / ms / 1e6);
    ms = time_ms([&] { red_kernel<4><<<blocks, threads>>>(buf, n, iters, 3); }); printf("random v4 red     : %.1f Gop/s  (%.1f Gfloat/s)\n", ops / ms / 1e6, 4 * ops / ms / 1e6);
    ms = time_ms([&] { red_cell_kernel<1><<<blocks, threads>>>(buf, n / 32, iters, 4); }); printf("cell(32ch) scalar  : %.1f Gfloat/s\n", ops / ms / 1e6);
    ms = time_ms([&] { red_cell_kernel<4><<<blocks, threads>>>(buf, n / 32, iters, 5); }); printf("cell(32ch) v4      : %.1f Gfloat/s\n", 4 * ops / ms / 1e6);
    float* o; cudaMalloc(&o, 148 * 2 * 4);
    cudaFuncSetAttribute(smem_rmw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    const int it2 = 4096;
    ms = time_ms([&] { smem_rmw_kernel<<<148 * 2, 256, 96 * 1024>>>(o, it2); });
    printf("smem RMW (LDS.128+STS.128): %.1f Gfloat/s chip, %.2f cycles/warp-RMW/SM @1.9GHz\n", 148.0 * 2 * 256 * it2 * 4 / ms / 1e6,
           ms * 1e-3 * 1.9e9 / ((double)2 * 8 * it2));
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
