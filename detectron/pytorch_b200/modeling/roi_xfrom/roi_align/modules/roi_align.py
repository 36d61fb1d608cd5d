"""Drop-in for `modeling.roi_xfrom.roi_align.modules.roi_align`
(reference lib/modeling/roi_xfrom/roi_align/modules/roi_align.py:6-45)."""
from torch.nn.modules.module import Module
from torch.nn.functional import avg_pool2d, max_pool2d
from ..functions.roi_align import RoIAlignFunction


class RoIAlign(Module):
    def __init__(self, aligned_height, aligned_width, spatial_scale, sampling_ratio):
        super(RoIAlign, self).__init__()
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)

    def forward(self, features, rois):
        return RoIAlignFunction(self.aligned_height, self.aligned_width,
                                self.spatial_scale, self.sampling_ratio)(features, rois)


class RoIAlignAvg(RoIAlign):
    """RoIAlign at (h+1) x (w+1) followed by a 2x2 stride-1 average pool (reference :19-31)."""

    def forward(self, features, rois):
        x = RoIAlignFunction(self.aligned_height + 1, self.aligned_width + 1,
                             self.spatial_scale, self.sampling_ratio)(features, rois)
        return avg_pool2d(x, kernel_size=2, stride=1)


class RoIAlignMax(RoIAlign):
    """RoIAlign at (h+1) x (w+1) followed by a 2x2 stride-1 max pool (reference :33-45)."""

    def forward(self, features, rois):
        x = RoIAlignFunction(self.aligned_height + 1, self.aligned_width + 1,
                             self.spatial_scale, self.sampling_ratio)(features, rois)
        return max_pool2d(x, kernel_size=2, stride=1)
