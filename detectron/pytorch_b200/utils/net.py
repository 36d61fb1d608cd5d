"""RoI -> sampling grid helper of the RoICrop pooling mode; mirrors `affine_grid_gen`, lib/utils/net.py:110-132
(reference), which Generalized_RCNN.roi_feature_transform calls before RoICropFunction (model_builder.py:280-288).

The reference was written against torch < 1.3, whose F.affine_grid sampled corner to corner; that behaviour is
`align_corners=True` today and is what the RoICrop kernel's coordinate mapping ((g + 1) * (size - 1) / 2) assumes.
"""
import torch
import torch.nn.functional as F


def affine_grid_gen(rois, input_size, grid_size):
    """rois (R, 5) [batch, x1, y1, x2, y2] in image pixels at stride 16; input_size = (H, W) of the feature map.
    Returns the (R, grid_size, grid_size, 2) sampling grid in torch's (x, y) order."""
    rois = rois.detach()
    x1, y1, x2, y2 = (rois[:, k:k + 1] / 16.0 for k in (1, 2, 3, 4))
    height, width = float(input_size[0]), float(input_size[1])
    zero = torch.zeros_like(x1)
    theta = torch.cat([(x2 - x1) / (width - 1), zero, (x1 + x2 - width + 1) / (width - 1),
                       zero, (y2 - y1) / (height - 1), (y1 + y2 - height + 1) / (height - 1)], 1).view(-1, 2, 3)
    return F.affine_grid(theta, torch.Size((rois.size(0), 1, grid_size, grid_size)), align_corners=True)
