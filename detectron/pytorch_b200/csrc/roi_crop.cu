// roi_crop.cu -- RoICrop: bilinear sampling of an NCHW image from an explicit (y,x) grid.
//
// Semantics: lib/model/roi_crop/src/roi_crop_cuda_kernel.cu (reference) getTopLeft :11-22,
// bilinearSamplingFromGrid :47-109, backwardBilinearSampling :111-194: align_corners-style mapping
// (g+1)(size-1)/2, zero padding per tap, RoI b reads image b / (R/N), image gradient by fp32 atomics,
// grid gradient never written (stays zero).  Forward values are bit-identical to the reference.
//
// Layout choice: the sampling geometry of an output pixel is the same for every channel, so one
// thread owns one output pixel (b, y, x), computes its 4 tap offsets + weights once, and walks the
// channel slab; adjacent threads are adjacent x, so output stores are coalesced.
#include "common.cuh"

namespace b200 {

constexpr int kCropThreads = 256;

struct CropTaps {
    int   off;                    // yTop * W + xLeft (may be negative; only used for taps that are in)
    bool  tl, tr, bl, br;
    float w_tl, w_tr, w_bl, w_br;
};

__device__ __forceinline__ void crop_topleft(float g, int size, int& point, float& weight) {
    const float coord = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)(size - 1)), 0.5f);
    const float fl = floorf(coord);
    point = (int)fl;
    weight = __fadd_rn(__fsub_rn(fl, coord), 1.f);
}

__device__ __forceinline__ CropTaps crop_taps(const float* __restrict__ g, int H, int W) {
    int yt, xl; float yW, xW;
    crop_topleft(g[1], W, xl, xW);
    crop_topleft(g[0], H, yt, yW);
    const bool xin0 = xl >= 0 && xl <= W - 1, xin1 = xl + 1 >= 0 && xl + 1 <= W - 1;
    const bool yin0 = yt >= 0 && yt <= H - 1, yin1 = yt + 1 >= 0 && yt + 1 <= H - 1;
    CropTaps t;
    t.off = yt * W + xl;
    t.tl = xin0 && yin0; t.tr = xin1 && yin0; t.bl = xin0 && yin1; t.br = xin1 && yin1;
    const float omx = __fsub_rn(1.f, xW), omy = __fsub_rn(1.f, yW);
    t.w_tl = __fmul_rn(xW, yW); t.w_tr = __fmul_rn(omx, yW);
    t.w_bl = __fmul_rn(xW, omy); t.w_br = __fmul_rn(omx, omy);
    return t;
}

template <bool BACKWARD>
__global__ void __launch_bounds__(kCropThreads)
roi_crop_kernel(const float* __restrict__ in,        // fwd: image       bwd: grad_output
                const float* __restrict__ grids,
                float* __restrict__ out,             // fwd: output      bwd: grad_image (pre-zeroed)
                int N, int C, int H, int W, int R, int oh, int ow, int c_per_cta) {
    const int per = R / N;                            // reference :217 roiPerImage = ob / ib
    const long pix = (long)blockIdx.x * kCropThreads + threadIdx.x;   // (b, y, x)
    const long npix = (long)R * oh * ow;
    if (pix >= npix) return;
    const int b = (int)(pix / ((long)oh * ow));
    const int yx = (int)(pix % ((long)oh * ow));
    const CropTaps t = crop_taps(grids + pix * 2, H, W);
    const int c0 = blockIdx.y * c_per_cta, c1 = min(C, c0 + c_per_cta);
    const bool any = t.tl || t.tr || t.bl || t.br;
    const int bi = per > 0 ? b / per : 0;
    const bool img_ok = bi < N;
    for (int c = c0; c < c1; ++c) {
        const size_t oidx = ((size_t)b * C + c) * oh * ow + yx;
        if (!any || !img_ok) {                        // reference `continue`s and leaves the zero-filled output
            if (!BACKWARD) out[oidx] = 0.f;
            continue;
        }
        const long base = ((long)bi * C + c) * H * W + t.off;
        if (!BACKWARD) {
            const float vtl = t.tl ? __ldg(in + base) : 0.f, vtr = t.tr ? __ldg(in + base + 1) : 0.f;
            const float vbl = t.bl ? __ldg(in + base + W) : 0.f, vbr = t.br ? __ldg(in + base + W + 1) : 0.f;
            out[oidx] = __fmaf_rn(t.w_br, vbr, __fmaf_rn(t.w_bl, vbl, __fmaf_rn(t.w_tl, vtl, __fmul_rn(t.w_tr, vtr))));
        } else {
            const float go = in[oidx];
            if (t.tl) atomicAdd(out + base, __fmul_rn(go, t.w_tl));
            if (t.tr) atomicAdd(out + base + 1, __fmul_rn(go, t.w_tr));
            if (t.bl) atomicAdd(out + base + W, __fmul_rn(go, t.w_bl));
            if (t.br) atomicAdd(out + base + W + 1, __fmul_rn(go, t.w_br));
        }
    }
}

static void crop_grid(int C, long npix, dim3& grid, int& cpc) {
    const long xb = (npix + kCropThreads - 1) / kCropThreads;
    cpc = C;
    while (cpc > 4 && xb * ((C + cpc - 1) / cpc) < 8L * kNumSMs) cpc /= 2;
    grid = dim3((unsigned)xb, (unsigned)((C + cpc - 1) / cpc));
}

int roi_crop_forward(const float* image, const float* grids, int N, int C, int H, int W, int R, int oh, int ow,
                     float* output, cudaStream_t stream) {
    const long npix = (long)R * oh * ow;
    if (npix == 0 || C == 0) return B200_ROI_OK;
    if (N <= 0) return B200_ROI_EINVAL;
    dim3 grid; int cpc;
    crop_grid(C, npix, grid, cpc);
    roi_crop_kernel<false><<<grid, kCropThreads, 0, stream>>>(image, grids, output, N, C, H, W, R, oh, ow, cpc);
    return finish_launch();
}

int roi_crop_backward(const float* grad_output, const float* grids, int N, int C, int H, int W, int R, int oh, int ow,
                      float* grad_image, float* grad_grids, cudaStream_t stream) {
    cudaError_t err = cudaMemsetAsync(grad_image, 0, sizeof(float) * (size_t)N * C * H * W, stream);
    if (err != cudaSuccess) return (int)err;
    int launches = 1;
    if (grad_grids != nullptr) {
        err = cudaMemsetAsync(grad_grids, 0, sizeof(float) * (size_t)R * oh * ow * 2, stream);
        if (err != cudaSuccess) return (int)err;
        ++launches;
    }
    const long npix = (long)R * oh * ow;
    if (npix == 0 || C == 0) return B200_ROI_OK;
    if (N <= 0) return B200_ROI_EINVAL;
    dim3 grid; int cpc;
    crop_grid(C, npix, grid, cpc);
    roi_crop_kernel<true><<<grid, kCropThreads, 0, stream>>>(grad_output, grids, grad_image, N, C, H, W, R, oh, ow, cpc);
    (void)launches;
    return finish_launch();
}

}  // namespace b200
