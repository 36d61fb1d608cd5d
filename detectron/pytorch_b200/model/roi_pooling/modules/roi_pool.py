"""`_RoIPooling(pooled_height, pooled_width, spatial_scale)` -- same import path, constructor and attributes as the
reference's lib/model/roi_pooling/modules/roi_pool.py:5-14; built by detectron.pytorch_b200._modules.roi_module."""
from detectron.pytorch_b200._modules import roi_module

from ..functions.roi_pool import RoIPoolFunction

_RoIPooling = roi_module(
    "_RoIPooling", RoIPoolFunction,
    fields=(("pooled_height", int), ("pooled_width", int), ("spatial_scale", float)),
    doc="Quantised max pooling of every RoI to pooled_height x pooled_width (forward(features, rois)).")
