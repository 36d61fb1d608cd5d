"""RoIAlign over an FPN pyramid with the results delivered in restored order.

Replaces, inside Generalized_RCNN.roi_feature_transform (lib/modeling/model_builder.py:264-303, reference), the
`for lvl in range(k_min, k_max + 1)` loop of RoIAlignFunction calls, the `torch.cat(bl_out_list, dim=0)` and the
`xform_shuffled[restore_bl]` gather by one call:

    xform_out = RoIAlignFPNFunction(resolution, resolution, spatial_scales, sampling_ratio)(
        [blobs_in[k_max - lvl] for lvl in range(k_min, k_max + 1)],
        [rois_of_level(lvl) for lvl in range(k_min, k_max + 1)],          # CUDA (R_l, 5); empty levels allowed
        rpn_ret[blob_rois + '_idx_restore_int32'])

The result equals the reference flow element for element (same kernels per level; each RoI's pooled block is written
straight to its restored row, so the concatenated intermediate and the gather pass -- two extra trips of the whole
output through HBM -- disappear, and the backward reads the gradient rows in place).
"""
from detectron.pytorch_b200 import ops


class RoIAlignFPNFunction(object):
    def __init__(self, aligned_height, aligned_width, spatial_scales, sampling_ratio):
        self.aligned_height = int(aligned_height)
        self.aligned_width = int(aligned_width)
        self.spatial_scales = [float(s) for s in spatial_scales]
        self.sampling_ratio = int(sampling_ratio)

    def __call__(self, features, rois, restore):
        return ops.roi_align_fpn(list(features), list(rois), restore, self.aligned_height, self.aligned_width,
                                 self.spatial_scales, self.sampling_ratio)

    forward = __call__
