// tma_probe4.cu -- does TMA accept the 4-D "x-quad interleaved" view of an NCHW map?
//   dims (x%4 : 4, c : C, x/4 : W/4, y : H), strides (-, H*W*4, 16, W*4) bytes, box (4, BC, XQ, 1)
//   -> shared memory [XQ x-quads][BC channels][4 floats]: every 16-byte piece is 4 consecutive columns of one channel, the
//   innermost start coordinate is always 0 (no alignment fault), out-of-range rows / quads / channels are zero-filled.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tma_probe4 tools/tma_probe4.cu -lcuda
//   ./tma_probe4 XQ0 Y C0 W [XQ] [BC]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tmap, int xq, int y, int c, int box_words, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned dst = (smem_u32(smem) + 127u) & ~127u;
    float* dstp = (float*)(smem + (dst - smem_u32(smem)));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(box_words * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(dst), "l"((unsigned long long)&tmap), "r"(smem_u32(&bar)), "r"(0), "r"(c), "r"(xq), "r"(y) : "memory");
    }
    unsigned ok = 0;
    long long t0 = clock64();
    while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        if (clock64() - t0 > 2000000000LL) { if (threadIdx.x == 0) out[0] = -12345.f; return; }
    }
    for (int i = threadIdx.x; i < box_words; i += blockDim.x) out[i] = dstp[i];
}

int main(int argc, char** argv) {
    const int XQ0 = argc > 1 ? atoi(argv[1]) : 14, Y = argc > 2 ? atoi(argv[2]) : 3, C0 = argc > 3 ? atoi(argv[3]) : 32;
    const int W = argc > 4 ? atoi(argv[4]) : 272, XQ = argc > 5 ? atoi(argv[5]) : 16, BC = argc > 6 ? atoi(argv[6]) : 32;
    const int H = 16, C = 80;
    std::vector<float> h((size_t)C * H * W);
    for (int c = 0; c < C; ++c) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) h[((size_t)c * H + y) * W + x] = c * 10000.f + y * 100.f + x * 0.25f;
    float *d, *out;
    cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    const int box_words = 4 * BC * XQ;
    cudaMalloc(&out, box_words * 4); cudaMemset(out, 0, box_words * 4);
    CUtensorMap tm;
    cuuint64_t dims[4] = {4, (cuuint64_t)C, (cuuint64_t)(W / 4), (cuuint64_t)H};
    cuuint64_t strides[3] = {(cuuint64_t)W * H * 4, 16, (cuuint64_t)W * 4};
    cuuint32_t box[4] = {4, (cuuint32_t)BC, (cuuint32_t)XQ, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("xq0=%d W=%d: encode failed %d\n", XQ0, W, (int)r); return 2; }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    probe<<<1, 128, box_words * 4 + 256>>>(tm, XQ0, Y, C0, box_words, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("xq0=%d y=%d c=%d W=%d: FAULT %s\n", XQ0, Y, C0, W, cudaGetErrorString(e)); return 1; }
    std::vector<float> o(box_words);
    cudaMemcpy(o.data(), out, box_words * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int q = 0; q < XQ; ++q) for (int k = 0; k < BC; ++k) for (int j = 0; j < 4; ++j) {
        const int c = C0 + k, x = 4 * (XQ0 + q) + j;
        const float want = (c < C && x >= 0 && x < W / 4 * 4 && Y >= 0 && Y < H) ? h[((size_t)c * H + Y) * W + x] : 0.f;
        const float got = o[(q * BC + k) * 4 + j];
        if (got != want) { if (bad < 3) printf("   mismatch q=%d k=%d j=%d got %f want %f\n", q, k, j, got, want); ++bad; }
    }
    printf("xq0=%d y=%d c=%d W=%d box=(4,%d,%d,1): %s (bad=%d, first=%f)\n", XQ0, Y, C0, W, BC, XQ, bad ? "WRONG" : "ok", bad, o[0]);
    return bad ? 3 : 0;
}
