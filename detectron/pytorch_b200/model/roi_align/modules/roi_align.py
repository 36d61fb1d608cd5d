"""`RoIAlign`, `RoIAlignAvg`, `RoIAlignMax` of the legacy flavour (one sample per lattice corner, no sampling ratio):
same import path, constructors `(aligned_height, aligned_width, spatial_scale)` and attributes as
lib/model/roi_align/modules/roi_align.py:6-42 (reference).  Avg / Max align at (h+1) x (w+1) and reduce with a
2x2 stride-1 pool.  Built by detectron.pytorch_b200._modules.roi_module."""
from detectron.pytorch_b200._modules import roi_module

from ..functions.roi_align import RoIAlignFunction

_FIELDS = (("aligned_height", int), ("aligned_width", int), ("spatial_scale", float))
_GROW = ("aligned_height", "aligned_width")

RoIAlign = roi_module("RoIAlign", RoIAlignFunction, _FIELDS, doc="Legacy RoIAlign (forward(features, rois)).")
RoIAlignAvg = roi_module("RoIAlignAvg", RoIAlignFunction, _FIELDS, grow=_GROW, epilogue="avg", base=RoIAlign,
                         doc="Legacy RoIAlign at (h+1) x (w+1), then 2x2 stride-1 average pooling.")
RoIAlignMax = roi_module("RoIAlignMax", RoIAlignFunction, _FIELDS, grow=_GROW, epilogue="max", base=RoIAlign,
                         doc="Legacy RoIAlign at (h+1) x (w+1), then 2x2 stride-1 max pooling.")
