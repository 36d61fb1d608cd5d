// roi_align_generic.cu -- Caffe2-exact RoIAlign, shape-generic path (any N,C,H,W,P, sampling_ratio
// incl. adaptive 0).  RoI-centric: one CTA per (RoI, channel slab); the per-axis sample tables
// (cell indices + weights, identical for all channels) are built once per CTA in shared memory and
// re-used by every channel, which removes the reference's per-output index arithmetic
// (3 int div/mod + 2 IEEE divisions per sample per output element).
//
// Semantics: lib/modeling/roi_xfrom/roi_align/src/roi_align_kernel.cu (reference) :16-63, :65-121,
// :150-193, :195-270.  Forward results are bit-identical to the reference kernel (same rounding
// recipe, same summation order); the backward adds the same per-tap terms with fp32 atomics, like
// the reference (order undefined there as well).
#include "common.cuh"

namespace b200 {

constexpr int kTabMax = 224;         // per-axis sample-table capacity (P * grid); larger RoIs compute on the fly
constexpr int kGenThreads = 256;

struct __align__(16) TabEntry {
    int   low;      // low cell index; high = low + dhigh
    int   dhigh;    // 0 or 1, or -1 if the sample is outside the map on this axis
    float l, h;
};

__device__ __forceinline__ TabEntry make_entry(float start, float bin, int p, int i, int grid, int size) {
    AxisTap t = xfrom_axis(xfrom_coord(start, bin, p, i, grid), size);
    TabEntry e;
    e.low = t.low;
    e.dhigh = t.valid ? (t.high - t.low) : -1;
    e.l = t.l;
    e.h = t.h;
    return e;
}

template <bool BACKWARD>
__global__ void __launch_bounds__(kGenThreads)
roi_align_generic_kernel(const float* __restrict__ in,      // fwd: bottom_data   bwd: top_diff
                         const float* __restrict__ rois,
                         float* __restrict__ out,           // fwd: top_data      bwd: bottom_diff (pre-zeroed)
                         float scale, int N, int C, int H, int W, int PH, int PW, int sr, int c_per_cta,
                         const int* __restrict__ row_map) {                 // RoI -> row of top_data / top_diff (null: identity)
    __shared__ TabEntry ytab[kTabMax];
    __shared__ TabEntry xtab[kTabMax];

    const int n = blockIdx.x;
    const int c0 = blockIdx.y * c_per_cta;
    const int c1 = min(C, c0 + c_per_cta);
    const XfromRoi g = xfrom_roi(rois + 5 * (size_t)n, scale, PH, PW, sr);
    const int gh = g.grid_h, gw = g.grid_w;
    const int ny = PH * gh, nx = PW * gw;
    const bool tab = (ny <= kTabMax) && (nx <= kTabMax) && gh > 0 && gw > 0;
    const bool batch_ok = (g.batch >= 0 && g.batch < N);

    if (tab) {
        for (int s = threadIdx.x; s < ny; s += kGenThreads) ytab[s] = make_entry(g.start_h, g.bin_h, s / gh, s % gh, gh, H);
        for (int s = threadIdx.x; s < nx; s += kGenThreads) xtab[s] = make_entry(g.start_w, g.bin_w, s / gw, s % gw, gw, W);
    }
    __syncthreads();

    const float count = (float)(gh * gw);
    const int bins = PH * PW;
    const int total = (c1 - c0) * bins;
    for (int idx = threadIdx.x; idx < total; idx += kGenThreads) {
        const int c = c0 + idx / bins;
        const int bin = idx % bins;
        const int ph = bin / PW, pw = bin % PW;
        const size_t oidx = ((size_t)(row_map ? row_map[n] : n) * C + c) * bins + bin;       // (row, c, ph, pw)
        if (!batch_ok) {                                            // reference would read out of bounds
            if (!BACKWARD) out[oidx] = 0.f;
            continue;
        }
        const size_t plane_off = ((size_t)g.batch * C + c) * H * W;
        float acc = 0.f;
        float gtop = 0.f;
        if (BACKWARD) gtop = in[oidx];
        for (int iy = 0; iy < gh; ++iy) {
            const TabEntry ey = tab ? ytab[ph * gh + iy] : make_entry(g.start_h, g.bin_h, ph, iy, gh, H);
            for (int ix = 0; ix < gw; ++ix) {
                const TabEntry ex = tab ? xtab[pw * gw + ix] : make_entry(g.start_w, g.bin_w, pw, ix, gw, W);
                if (ey.dhigh < 0 || ex.dhigh < 0) continue;          // sample outside: contributes +0
                const float w1 = __fmul_rn(ey.h, ex.h), w2 = __fmul_rn(ey.h, ex.l);
                const float w3 = __fmul_rn(ey.l, ex.h), w4 = __fmul_rn(ey.l, ex.l);
                const size_t i1 = plane_off + (size_t)ey.low * W + ex.low;
                const size_t i2 = i1 + ex.dhigh;
                const size_t i3 = i1 + (size_t)ey.dhigh * W;
                const size_t i4 = i3 + ex.dhigh;
                if (!BACKWARD) {
                    const float v1 = __ldg(in + i1), v2 = __ldg(in + i2), v3 = __ldg(in + i3), v4 = __ldg(in + i4);
                    const float val = __fmaf_rn(v4, w4, __fmaf_rn(v3, w3, __fmaf_rn(v1, w1, __fmul_rn(v2, w2))));
                    acc = __fadd_rn(acc, val);
                } else {
                    atomicAdd(out + i1, __fdiv_rn(__fmul_rn(gtop, w1), count));
                    atomicAdd(out + i2, __fdiv_rn(__fmul_rn(gtop, w2), count));
                    atomicAdd(out + i3, __fdiv_rn(__fmul_rn(gtop, w3), count));
                    atomicAdd(out + i4, __fdiv_rn(__fmul_rn(gtop, w4), count));
                }
            }
        }
        if (!BACKWARD) out[oidx] = __fdiv_rn(acc, count);
    }
}

static int pick_c_per_cta(int num_rois, int C, int bins) {
    // aim for >= 4 waves of CTAs over 148 SMs while keeping >= one pass of kGenThreads outputs per CTA
    int c_per_cta = C;
    while (c_per_cta > 1 && (long)num_rois * ((C + c_per_cta - 1) / c_per_cta) < 8L * kNumSMs &&
           (c_per_cta / 2) * bins >= kGenThreads)
        c_per_cta /= 2;
    return c_per_cta;
}

int roi_align_forward_generic(const float* bottom, float scale, int N, int R, int H, int W, int C, int PH, int PW,
                              int sr, const float* rois, float* top, const int* row_map, cudaStream_t stream) {
    if (R == 0 || C == 0) return B200_ROI_OK;
    const int cpc = pick_c_per_cta(R, C, PH * PW);
    dim3 grid(R, (C + cpc - 1) / cpc);
    roi_align_generic_kernel<false><<<grid, kGenThreads, 0, stream>>>(bottom, rois, top, scale, N, C, H, W, PH, PW, sr, cpc, row_map);
    return finish_launch();
}

int roi_align_backward_generic(const float* top_diff, float scale, int N, int R, int H, int W, int C, int PH, int PW,
                               int sr, const float* rois, float* bottom_diff, const int* row_map, cudaStream_t stream) {
    cudaError_t err = cudaMemsetAsync(bottom_diff, 0, sizeof(float) * (size_t)N * C * H * W, stream);
    if (err != cudaSuccess) return (int)err;
    if (R == 0 || C == 0) return B200_ROI_OK;
    const int cpc = pick_c_per_cta(R, C, PH * PW);
    dim3 grid(R, (C + cpc - 1) / cpc);
    roi_align_generic_kernel<true><<<grid, kGenThreads, 0, stream>>>(top_diff, rois, bottom_diff, scale, N, C, H, W, PH, PW, sr, cpc, row_map);
    return finish_launch();   // (the cudaMemsetAsync is not one of our kernels)
}

}  // namespace b200
