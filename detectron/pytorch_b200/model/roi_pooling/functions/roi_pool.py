"""Drop-in for `model.roi_pooling.functions.roi_pool`
(reference lib/model/roi_pooling/functions/roi_pool.py:6-38).

CUDA only: the reference's CPU branch feeds a non-contiguous permute to an NHWC C routine and reads
the wrong memory, so it is not reproduced; CPU input raises NotImplementedError."""
from detectron.pytorch_b200 import ops as _ops


class RoIPoolFunction(object):
    def __init__(ctx, pooled_height, pooled_width, spatial_scale):
        ctx.pooled_width = pooled_width
        ctx.pooled_height = pooled_height
        ctx.spatial_scale = spatial_scale
        ctx.feature_size = None
        ctx.argmax = None
        ctx.rois = None

    def __call__(ctx, features, rois):
        ctx.feature_size = features.size()
        ctx.rois = rois
        if not features.is_cuda:
            raise NotImplementedError("RoIPoolFunction: CUDA tensors only")
        output, argmax = _ops._RoIPool.apply(features, rois, int(ctx.pooled_height), int(ctx.pooled_width),
                                             float(ctx.spatial_scale))
        ctx.argmax = argmax
        return output

    def forward(ctx, features, rois):
        ctx.feature_size = features.size()
        ctx.rois = rois
        if not features.is_cuda:
            raise NotImplementedError("RoIPoolFunction: CUDA tensors only")
        output, ctx.argmax = _ops.roi_pool_forward(features.detach(), rois.detach(), int(ctx.pooled_height),
                                                   int(ctx.pooled_width), float(ctx.spatial_scale))
        return output

    def backward(ctx, grad_output):
        assert(ctx.feature_size is not None and grad_output.is_cuda)
        grad_input = _ops.roi_pool_backward(grad_output, ctx.argmax, ctx.rois, tuple(ctx.feature_size),
                                            int(ctx.pooled_height), int(ctx.pooled_width), float(ctx.spatial_scale))
        return grad_input, None
