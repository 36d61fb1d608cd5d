// ref_launchers.cu -- the reference's launcher names and argument lists on top of the b200_* C ABI
// (include/b200_ref_launchers.h).  Thin by design: every function is argument re-ordering + the scratch the
// reference signature has no slot for.
#include "../../../include/b200_ref_launchers.h"
#include "../../../include/b200_roi_ops.h"
#include <limits.h>
#include <stdio.h>

namespace {
// The forward launchers carry no batch size: accept any batch index, as the reference kernels do.
constexpr int kAnyBatch = INT_MAX;

int done(int status, const char* who) {
    if (status == B200_ROI_OK) return 1;
    fprintf(stderr, "%s: %s\n", who, b200_roi_ops_strerror(status));      // the reference prints and exit(-1)s here
    return 0;
}
}  // namespace

extern "C" {

#ifndef B200_REF_LEGACY_ROI_ALIGN

int ROIAlignForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois, const int height, const int width,
                           const int channels, const int aligned_height, const int aligned_width, const int sampling_ratio,
                           const float* bottom_rois, float* top_data, cudaStream_t stream) {
    // workspace NULL -> the shape-generic kernel (bit-identical to the reference kernel), no batch-size-dependent planning
    return done(b200_roi_align_forward_ws(bottom_data, spatial_scale, kAnyBatch, num_rois, height, width, channels, aligned_height,
                                          aligned_width, sampling_ratio, bottom_rois, top_data, nullptr, 0, (b200_stream_t)stream),
                "ROIAlignForwardLaucher");
}

int ROIAlignBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size, const int num_rois, const int height,
                            const int width, const int channels, const int aligned_height, const int aligned_width,
                            const int sampling_ratio, const float* bottom_rois, float* bottom_diff, cudaStream_t stream) {
    return done(b200_roi_align_backward(top_diff, spatial_scale, batch_size, num_rois, height, width, channels, aligned_height,
                                        aligned_width, sampling_ratio, bottom_rois, bottom_diff, (b200_stream_t)stream),
                "ROIAlignBackwardLaucher");
}

int ROIPoolForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois, const int height, const int width,
                          const int channels, const int pooled_height, const int pooled_width, const float* bottom_rois,
                          float* top_data, int* argmax_data, cudaStream_t stream) {
    return done(b200_roi_pool_forward(bottom_data, spatial_scale, kAnyBatch, num_rois, height, width, channels, pooled_height,
                                      pooled_width, bottom_rois, top_data, argmax_data, (b200_stream_t)stream),
                "ROIPoolForwardLaucher");
}

int ROIPoolBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size, const int num_rois, const int height,
                           const int width, const int channels, const int pooled_height, const int pooled_width,
                           const float* bottom_rois, float* bottom_diff, const int* argmax_data, cudaStream_t stream) {
    return done(b200_roi_pool_backward(top_diff, spatial_scale, batch_size, num_rois, height, width, channels, pooled_height,
                                       pooled_width, bottom_rois, bottom_diff, argmax_data, (b200_stream_t)stream),
                "ROIPoolBackwardLaucher");
}

// Dense-layout check: images (B, C, H, W), grids (B, h, w, 2) with the glue's stride order (batch, coordinate, row,
// column), outputs (B, C, h, w) -- what RoICropFunction produces (roi_crop.py:9-12: clones + a fresh output).
static bool dense_bchw(int sb, int sc, int sh, int sw, int c, int h, int w) {
    return sw == 1 && sh == w && sc == h * w && sb == c * h * w;
}
static bool dense_grid(int gsb, int gsc, int gsh, int gsw, int h, int w) {
    return gsc == 1 && gsw == 2 && gsh == 2 * w && gsb == 2 * h * w;
}

int BilinearSamplerBHWD_updateOutput_cuda_kernel(int oc, int ow, int oh, int ob, int ic, int ih, int iw, int ib, float* inputImages,
                                                 int isb, int isc, int ish, int isw, float* grids, int gsb, int gsc, int gsh, int gsw,
                                                 float* output, int osb, int osc, int osh, int osw, cudaStream_t stream) {
    if (oc != ic || !dense_bchw(isb, isc, ish, isw, ic, ih, iw) || !dense_grid(gsb, gsc, gsh, gsw, oh, ow) ||
        !dense_bchw(osb, osc, osh, osw, oc, oh, ow))
        return 0;
    return done(b200_roi_crop_forward(inputImages, grids, ib, ic, ih, iw, ob, oh, ow, output, (b200_stream_t)stream),
                "BilinearSamplerBHWD_updateOutput_cuda_kernel");
}

int BilinearSamplerBHWD_updateGradInput_cuda_kernel(int goc, int gow, int goh, int gob, int ic, int ih, int iw, int ib,
                                                    float* inputImages, int isb, int isc, int ish, int isw, float* grids, int gsb,
                                                    int gsc, int gsh, int gsw, float* gradInputImages, int gisb, int gisc, int gish,
                                                    int gisw, float* gradGrids, int ggsb, int ggsc, int ggsh, int ggsw,
                                                    float* gradOutput, int gosb, int gosc, int gosh, int gosw, cudaStream_t stream) {
    (void)inputImages;
    if (goc != ic || !dense_bchw(isb, isc, ish, isw, ic, ih, iw) || !dense_grid(gsb, gsc, gsh, gsw, goh, gow) ||
        !dense_bchw(gisb, gisc, gish, gisw, ic, ih, iw) || !dense_grid(ggsb, ggsc, ggsh, ggsw, goh, gow) ||
        !dense_bchw(gosb, gosc, gosh, gosw, goc, goh, gow))
        return 0;
    return done(b200_roi_crop_backward(gradOutput, grids, ib, ic, ih, iw, gob, goh, gow, gradInputImages, gradGrids,
                                       (b200_stream_t)stream),
                "BilinearSamplerBHWD_updateGradInput_cuda_kernel");
}

void nms_cuda_compute(int* keep_out, int* num_out, float* boxes_host, int boxes_num, int boxes_dim, float nms_overlap_thresh) {
    // `boxes_host` is a DEVICE pointer in the reference too (nms_cuda.c:13 passes THCudaTensor_data); the name is historical.
    const size_t wsb = b200_nms_workspace_bytes(boxes_num);
    void* ws = nullptr;
    cudaStream_t stream = 0;                                     // the reference runs on the legacy default stream
    if (cudaMallocAsync(&ws, wsb, stream) != cudaSuccess) { fprintf(stderr, "nms_cuda_compute: no scratch\n"); return; }
    done(b200_nms(boxes_host, boxes_num, boxes_dim, nms_overlap_thresh, keep_out, num_out, ws, wsb, (b200_stream_t)stream),
         "nms_cuda_compute");
    cudaFreeAsync(ws, stream);
}

#else  // ---------------------------------------------------------------- legacy RoIAlign flavour (same names, no sampling_ratio)

int ROIAlignForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois, const int height, const int width,
                           const int channels, const int aligned_height, const int aligned_width, const float* bottom_rois,
                           float* top_data, cudaStream_t stream) {
    return done(b200_roi_align_legacy_forward(bottom_data, spatial_scale, kAnyBatch, num_rois, height, width, channels, aligned_height,
                                              aligned_width, bottom_rois, top_data, (b200_stream_t)stream),
                "ROIAlignForwardLaucher (legacy)");
}

int ROIAlignBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size, const int num_rois, const int height,
                            const int width, const int channels, const int aligned_height, const int aligned_width,
                            const float* bottom_rois, float* bottom_diff, cudaStream_t stream) {
    return done(b200_roi_align_legacy_backward(top_diff, spatial_scale, batch_size, num_rois, height, width, channels, aligned_height,
                                               aligned_width, bottom_rois, bottom_diff, (b200_stream_t)stream),
                "ROIAlignBackwardLaucher (legacy)");
}

#endif

}  // extern "C"
