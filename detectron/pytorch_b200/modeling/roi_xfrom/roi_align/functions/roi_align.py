"""Drop-in for the reference's `modeling.roi_xfrom.roi_align.functions.roi_align`
(lib/modeling/roi_xfrom/roi_align/functions/roi_align.py:7-48): same class name, constructor
arguments, call protocol and error behaviour, backed by the sm_100a kernels.

The reference class is a legacy (pre-0.4) autograd Function: it is constructed per call with the
hyper-parameters and then *called* with the tensors -- `RoIAlignFunction(h, w, scale, sr)(features,
rois)` (lib/modeling/model_builder.py:290-291).  torch >= 1.3 rejects non-static Functions, so this
is a plain callable that delegates to a static Function; `.forward` / `.backward` remain usable
directly and keep the reference's instance state (`rois`, `feature_size`).

Numerics at this surface: the forward is deterministic and bit-identical to the reference kernel,
except for bins whose samples cannot share a strip / the row ring (boxes wider than ~84 cells or taller than the ring at
`sampling_ratio` 2): those are two partial sums, equal to the reference within 1e-7 and still run-to-run identical
(`B200_ROI_ALIGN_PATH=generic` is the always-bit-exact switch).  A feature map that holds Inf / NaN can poison bins the
reference would not let that cell reach: samples the reference skips are evaluated with weight 0.  The backward matches the
reference within 1e-5 (fp32 accumulation order; the reference's atomicAdd order is itself undefined) and is not run-to-run
bit-identical on the gather path (unit order comes from integer atomics).
"""
from detectron.pytorch_b200 import ops as _ops


class RoIAlignFunction(object):
    def __init__(self, aligned_height, aligned_width, spatial_scale, sampling_ratio):
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.rois = None
        self.feature_size = None

    def __call__(self, features, rois):
        self.rois = rois
        self.feature_size = features.size()
        if not features.is_cuda:
            raise NotImplementedError
        return _ops._RoIAlign.apply(features, rois, self.aligned_height, self.aligned_width,
                                    self.spatial_scale, self.sampling_ratio)

    def forward(self, features, rois):
        self.rois = rois
        self.feature_size = features.size()
        if not features.is_cuda:
            raise NotImplementedError
        return _ops.roi_align_forward(features.detach(), rois.detach(), self.aligned_height, self.aligned_width,
                                      self.spatial_scale, self.sampling_ratio)

    def backward(self, grad_output):
        assert(self.feature_size is not None and grad_output.is_cuda)
        grad_input = _ops.roi_align_backward(grad_output, self.rois, tuple(self.feature_size), self.aligned_height,
                                             self.aligned_width, self.spatial_scale, self.sampling_ratio)
        return grad_input, None
