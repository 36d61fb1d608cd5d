/*
 * b200_ref_launchers.h -- the reference's own raw-pointer launcher NAMES and ARGUMENT LISTS, exported by the small
 * compatibility libraries next to libb200_roi_ops.so, so that the reference's C glue (lib/.../src/*_cuda.c, which calls
 * these launchers after unpacking THCudaTensors) links against this implementation unchanged:
 *
 *   libb200_ref_launchers.so          ROIAlign{Forward,Backward}Laucher   (Caffe2-exact flavour)
 *                                       lib/modeling/roi_xfrom/roi_align/src/roi_align_kernel.h:13-27
 *                                     ROIPool{Forward,Backward}Laucher    lib/model/roi_pooling/src/roi_pooling_kernel.h:8-18
 *                                     BilinearSamplerBHWD_update{Output,GradInput}_cuda_kernel
 *                                                                         lib/model/roi_crop/src/roi_crop_cuda_kernel.h:6-32
 *                                     nms_cuda_compute                    lib/model/nms/src/nms_cuda_kernel.h:5-6
 *   libb200_ref_launchers_legacy.so   ROIAlign{Forward,Backward}Laucher   (legacy flavour: same names, no sampling_ratio)
 *                                       lib/model/roi_align/src/roi_align_kernel.h:13-25
 *
 * Behaviour kept from the reference: return 1 after a successful launch (the glue ignores it); work is enqueued on
 * `stream` (nms_cuda_compute: the legacy default stream, like the reference); no memory is retained.
 * Differences, all forced by the signatures: the forward launchers get no batch size, so batch indices are not range
 * checked (the reference does not check either) and the shape-generic kernels run -- the fast paths need the batch
 * size and caller scratch, i.e. the b200_* entry points of b200_roi_ops.h (INTEGRATION.md shows the one-line change
 * to the glue); scratch for the backward / NMS comes from cudaMallocAsync on the call's stream; a launch error
 * returns 0 instead of exit(-1); the RoICrop launchers accept dense tensors only (the strides the reference's Python
 * layer produces) and return 0 for anything else (the glue turns 0 into THError).
 */
#ifndef B200_REF_LAUNCHERS_H_
#define B200_REF_LAUNCHERS_H_

#include <cuda_runtime.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200_REF_API __attribute__((visibility("default")))
#else
#define B200_REF_API
#endif

#ifndef B200_REF_LEGACY_ROI_ALIGN
B200_REF_API int ROIAlignForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois, const int height,
                                        const int width, const int channels, const int aligned_height, const int aligned_width,
                                        const int sampling_ratio, const float* bottom_rois, float* top_data, cudaStream_t stream);
B200_REF_API int ROIAlignBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size, const int num_rois,
                                         const int height, const int width, const int channels, const int aligned_height,
                                         const int aligned_width, const int sampling_ratio, const float* bottom_rois,
                                         float* bottom_diff, cudaStream_t stream);
B200_REF_API int ROIPoolForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois, const int height,
                                       const int width, const int channels, const int pooled_height, const int pooled_width,
                                       const float* bottom_rois, float* top_data, int* argmax_data, cudaStream_t stream);
B200_REF_API int ROIPoolBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size, const int num_rois,
                                        const int height, const int width, const int channels, const int pooled_height,
                                        const int pooled_width, const float* bottom_rois, float* bottom_diff,
                                        const int* argmax_data, cudaStream_t stream);
B200_REF_API int BilinearSamplerBHWD_updateOutput_cuda_kernel(int oc, int ow, int oh, int ob, int ic, int ih, int iw, int ib,
                                                              float* inputImages, int isb, int isc, int ish, int isw,
                                                              float* grids, int gsb, int gsc, int gsh, int gsw,
                                                              float* output, int osb, int osc, int osh, int osw, cudaStream_t stream);
B200_REF_API int BilinearSamplerBHWD_updateGradInput_cuda_kernel(int goc, int gow, int goh, int gob, int ic, int ih, int iw, int ib,
                                                                 float* inputImages, int isb, int isc, int ish, int isw,
                                                                 float* grids, int gsb, int gsc, int gsh, int gsw,
                                                                 float* gradInputImages, int gisb, int gisc, int gish, int gisw,
                                                                 float* gradGrids, int ggsb, int ggsc, int ggsh, int ggsw,
                                                                 float* gradOutput, int gosb, int gosc, int gosh, int gosw,
                                                                 cudaStream_t stream);
B200_REF_API void nms_cuda_compute(int* keep_out, int* num_out, float* boxes_host, int boxes_num, int boxes_dim,
                                   float nms_overlap_thresh);
#else
B200_REF_API int ROIAlignForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois, const int height,
                                        const int width, const int channels, const int aligned_height, const int aligned_width,
                                        const float* bottom_rois, float* top_data, cudaStream_t stream);
B200_REF_API int ROIAlignBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size, const int num_rois,
                                         const int height, const int width, const int channels, const int aligned_height,
                                         const int aligned_width, const float* bottom_rois, float* bottom_diff, cudaStream_t stream);
#endif

#ifdef __cplusplus
}
#endif
#endif
