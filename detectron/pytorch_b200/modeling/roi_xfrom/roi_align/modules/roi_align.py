"""`RoIAlign`, `RoIAlignAvg`, `RoIAlignMax` of the Caffe2-exact flavour: same import path, constructors
`(aligned_height, aligned_width, spatial_scale, sampling_ratio)` and attributes as
lib/modeling/roi_xfrom/roi_align/modules/roi_align.py:6-45 (reference).  Avg / Max align at (h+1) x (w+1) and reduce
with a 2x2 stride-1 pool (:19-31, :33-45).  Built by detectron.pytorch_b200._modules.roi_module."""
from detectron.pytorch_b200._modules import roi_module

from ..functions.roi_align import RoIAlignFunction

_FIELDS = (("aligned_height", int), ("aligned_width", int), ("spatial_scale", float), ("sampling_ratio", int))
_GROW = ("aligned_height", "aligned_width")

RoIAlign = roi_module("RoIAlign", RoIAlignFunction, _FIELDS, doc="Caffe2-exact RoIAlign (forward(features, rois)).")
RoIAlignAvg = roi_module("RoIAlignAvg", RoIAlignFunction, _FIELDS, grow=_GROW, epilogue="avg", base=RoIAlign,
                         doc="RoIAlign at (h+1) x (w+1), then 2x2 stride-1 average pooling.")
RoIAlignMax = roi_module("RoIAlignMax", RoIAlignFunction, _FIELDS, grow=_GROW, epilogue="max", base=RoIAlign,
                         doc="RoIAlign at (h+1) x (w+1), then 2x2 stride-1 max pooling.")
