"""Drop-in for the reference's legacy `model.roi_align.functions.roi_align`
(lib/model/roi_align/functions/roi_align.py:7-47): 3-argument constructor (no sampling_ratio), one
bilinear sample per lattice corner.  See the xfrom variant for the call-protocol notes."""
from detectron.pytorch_b200 import ops as _ops


class RoIAlignFunction(object):
    def __init__(self, aligned_height, aligned_width, spatial_scale):
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)
        self.rois = None
        self.feature_size = None

    def __call__(self, features, rois):
        self.rois = rois
        self.feature_size = features.size()
        if not features.is_cuda:
            raise NotImplementedError
        return _ops._RoIAlignLegacy.apply(features, rois, self.aligned_height, self.aligned_width, self.spatial_scale)

    def forward(self, features, rois):
        self.rois = rois
        self.feature_size = features.size()
        if not features.is_cuda:
            raise NotImplementedError
        return _ops.roi_align_legacy_forward(features.detach(), rois.detach(), self.aligned_height,
                                             self.aligned_width, self.spatial_scale)

    def backward(self, grad_output):
        assert(self.feature_size is not None and grad_output.is_cuda)
        grad_input = _ops.roi_align_legacy_backward(grad_output, self.rois, tuple(self.feature_size),
                                                    self.aligned_height, self.aligned_width, self.spatial_scale)
        return grad_input, None
