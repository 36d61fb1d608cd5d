// common.cuh -- shared helpers for libb200_roi_ops.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../../include/b200_roi_ops.h"

namespace b200 {

// ---- launch accounting / error plumbing -------------------------------------------------------
extern unsigned long long g_launch_count;   // defined in api.cu

inline int finish_launch(int n_launches = 1) {
    __atomic_fetch_add(&g_launch_count, (unsigned long long)n_launches, __ATOMIC_RELAXED);   // callers may be one thread per GPU
    cudaError_t err = cudaGetLastError();
    return (err == cudaSuccess) ? B200_ROI_OK : (int)err;
}

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

// ---- dispatch switches -------------------------------------------------------------------------
// A/B and test overrides of the path selection.  Each switch is read from the environment ONCE (first use) and can
// be changed afterwards through b200_roi_ops_set_option() -- the hot path never calls getenv().
// The value is the first character of the string ('\0' = auto / default).
enum Option {
    kOptFwdPath = 0,      // B200_ROI_ALIGN_PATH      = auto | generic | tiled | stream | quad (roi_align_strip.cu, the default fast path)
    kOptBwdPath,          // B200_ROI_ALIGN_BWD_PATH  = auto | generic | nhwc | rows
    kOptBwdCpl,           // B200_ROI_ALIGN_BWD_CPL   = 4 | 2
    kOptFwdZero,          // B200_FWD_ZERO            = dense | bins
    kOptNmsScan,          // B200_NMS_SCAN            = resolver | simple
    kOptStreamStage,      // B200_STREAM_STAGE        = async (cp.async) | regs (LDG -> registers -> STS)
    kOptStreamPhases,     // B200_STREAM_PHASES       = all | prepass (timing probe: the main kernel is not launched)
    kOptFpnPath,          // B200_FPN_PATH            = fused (one launch sequence over the level table) | levels (per-level calls)
    kOptStripRowCost,     // B200_STRIP_ROWCOST       = 0..9: cost of streaming a row in the partition model = 8 d + 4 (default 24)
    kOptStripPdl,         // B200_STRIP_PDL           = 1 | 0: programmatic dependent launch of the main kernel behind the prepass
    kOptBwdTrCh,          // B200_BWD_TRCH            = 2 (256) | 1 (128) | 6 (64): channels per CTA of the backward's dY transpose
    kNumOptions
};
int option_get(Option which);

// ---- arithmetic that must round exactly like the reference kernels' SASS --------------------
// All helpers use the _rn intrinsics, which nvcc never contracts or re-associates, so the
// fusion pattern is exactly what is written here (see oracle/roi_ops_oracle.c for the recipes
// and how they were read off `cuobjdump -sass` of the reference kernels built for sm_100a).

// RoI geometry of the Caffe2-exact RoIAlign (reference roi_align_kernel.cu:75-98).
struct XfromRoi {
    int   batch;
    float start_w, start_h, bin_w, bin_h;
    int   grid_w, grid_h;
};

__device__ __forceinline__ XfromRoi xfrom_roi(const float* __restrict__ roi, float scale, int PH, int PW, int sr) {
    XfromRoi g;
    g.batch = (int)roi[0];
    g.start_w = __fmul_rn(roi[1], scale);
    g.start_h = __fmul_rn(roi[2], scale);
    float rw = fmaxf(__fmaf_rn(roi[3], scale, -g.start_w), 1.f);
    float rh = fmaxf(__fmaf_rn(roi[4], scale, -g.start_h), 1.f);
    g.bin_h = __fdiv_rn(rh, (float)PH);
    g.bin_w = __fdiv_rn(rw, (float)PW);
    g.grid_h = (sr > 0) ? sr : (int)ceilf(g.bin_h);
    g.grid_w = (sr > 0) ? sr : (int)ceilf(g.bin_w);
    return g;
}

// sample coordinate: FADD(FFMA(p, bin, start), FMUL(i + .5, bin) / grid)   (:106-110)
__device__ __forceinline__ float xfrom_coord(float start, float bin, int p, int i, int grid) {
    float base = __fmaf_rn((float)p, bin, start);
    float off = __fdiv_rn(__fmul_rn(__fadd_rn((float)i, .5f), bin), (float)grid);
    return __fadd_rn(base, off);
}

// One axis of bilinear_interpolate (:19-52): low/high cell and the two weights (l = frac, h = 1-l).
// valid == false  <=>  the reference's "outside the map" early-out for this axis.
struct AxisTap {
    int   low, high;
    float l, h;
    bool  valid;
};

__device__ __forceinline__ AxisTap xfrom_axis(float v, int size) {
    AxisTap t;
    t.valid = !(v < -1.0f || v > (float)size);
    if (v <= 0.f) v = 0.f;
    int low = (int)v;
    if (low >= size - 1) {
        t.low = t.high = size - 1;
        v = (float)(size - 1);
    } else {
        t.low = low;
        t.high = low + 1;
    }
    t.l = __fsub_rn(v, (float)t.low);
    t.h = __fsub_rn(1.f, t.l);
    return t;
}

// ---- IoU exactly as the reference's compiled Cython computes it (lib/utils/cython_bbox.pyx:52-72) ----------------------
// The generated C adds the literal 1.0 as a DOUBLE: widths / heights are (double)(float difference) + 1.0, their products and the
// union are formed in double (separate multiply and add: x86-64 baseline has no FMA) and rounded to float once on assignment
// (box_area, iw, ih, ua are float variables); the final quotient (iw * ih) / ua is float.  Up to 2 ulp away from an all-float
// evaluation -- enough to flip a threshold test.
__device__ __forceinline__ float cython_area(float x1, float y1, float x2, float y2) {
    return (float)__dmul_rn(__dadd_rn((double)__fsub_rn(x2, x1), 1.0), __dadd_rn((double)__fsub_rn(y2, y1), 1.0));
}
// b: the "boxes" row, q: the "query_boxes" row, qarea = cython_area(q)
__device__ __forceinline__ float cython_iou(float bx1, float by1, float bx2, float by2, float qx1, float qy1, float qx2, float qy2,
                                            float qarea) {
    const float iw = (float)__dadd_rn((double)__fsub_rn(fminf(bx2, qx2), fmaxf(bx1, qx1)), 1.0);
    if (!(iw > 0.f)) return 0.f;
    const float ih = (float)__dadd_rn((double)__fsub_rn(fminf(by2, qy2), fmaxf(by1, qy1)), 1.0);
    if (!(ih > 0.f)) return 0.f;
    const float inter = __fmul_rn(iw, ih);
    const double barea = __dmul_rn(__dadd_rn((double)__fsub_rn(bx2, bx1), 1.0), __dadd_rn((double)__fsub_rn(by2, by1), 1.0));
    const float ua = (float)__dsub_rn(__dadd_rn(barea, (double)qarea), (double)inter);
    return __fdiv_rn(inter, ua);
}

}  // namespace b200
