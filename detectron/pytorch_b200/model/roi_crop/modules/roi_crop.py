"""`_RoICrop(layout='BHWD')` -- same import path and call as lib/model/roi_crop/modules/roi_crop.py:4-8 (reference);
the layout argument is accepted and ignored there too.  Built by detectron.pytorch_b200._modules.roi_module."""
from detectron.pytorch_b200._modules import _RoIModule, roi_module

from ..functions.roi_crop import RoICropFunction


class _CropBase(_RoIModule):
    def __init__(self, layout='BHWD'):
        super().__init__()
        self.layout = layout


_RoICrop = roi_module("_RoICrop", RoICropFunction, fields=(), base=_CropBase,
                      doc="Bilinear crop of input1 (N, C, H, W) at the sampling grid input2 (R, h, w, 2) in (y, x) order.")
