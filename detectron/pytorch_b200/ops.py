"""torch.autograd.Function front-ends over the C ABI (new-style, static) + functional helpers.

These are what the reference-shaped shims in `model/` and `modeling/` delegate to.  Every op
  * requires CUDA fp32 tensors (the reference raises NotImplementedError / asserts for CPU input),
  * makes its inputs contiguous (the reference passed raw pointers and silently assumed it),
  * launches on torch's current stream of the tensor's device,
  * allocates outputs with torch.empty (the kernels define every output element; the reference
    needed `.zero_()` passes, functions/roi_align.py:23,39-40).
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda_f32(t, name):
    if not t.is_cuda:
        raise NotImplementedError("%s must be a CUDA tensor (the reference has no CPU path for this op)" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (name, t.dtype))


def _rois_ok(rois):
    if rois.dim() != 2 or rois.size(1) != 5:
        raise ValueError("rois must be (R, 5) = [batch_idx, x1, y1, x2, y2], got %s" % (tuple(rois.shape),))


# ------------------------------------------------------------------------------------------------
# RoIAlign (Caffe2-exact, sampling_ratio)
# ------------------------------------------------------------------------------------------------
def roi_align_forward(features, rois, aligned_height, aligned_width, spatial_scale, sampling_ratio):
    _need_cuda_f32(features, "features"); _need_cuda_f32(rois, "rois"); _rois_ok(rois)
    features = features.contiguous(); rois = rois.contiguous()
    N, C, H, W = features.shape
    R = rois.size(0)
    out = torch.empty((R, C, aligned_height, aligned_width), dtype=torch.float32, device=features.device)
    lib = _lib.load()
    ws_bytes = int(lib.b200_roi_align_workspace_bytes(N, R, H, W, aligned_height, aligned_width, sampling_ratio))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=features.device) if ws_bytes else None   # caching allocator, stream-ordered
    with torch.cuda.device(features.device):
        _lib.check(lib.b200_roi_align_forward_ws(features.data_ptr(), spatial_scale, N, R, H, W, C, aligned_height,
                                                 aligned_width, sampling_ratio, rois.data_ptr(), out.data_ptr(),
                                                 ws.data_ptr() if ws is not None else None, ws_bytes, _stream()),
                   "b200_roi_align_forward_ws")
    return out


def roi_align_backward(grad_output, rois, feature_size, aligned_height, aligned_width, spatial_scale, sampling_ratio):
    _need_cuda_f32(grad_output, "grad_output")
    grad_output = grad_output.contiguous(); rois = rois.contiguous()
    N, C, H, W = feature_size
    grad_input = torch.empty((N, C, H, W), dtype=torch.float32, device=grad_output.device)
    lib = _lib.load()
    ws_bytes = int(lib.b200_roi_align_backward_workspace_bytes(N, rois.size(0), C, H, W, aligned_height, aligned_width,
                                                              sampling_ratio)) if rois.size(0) > 0 else 0
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=grad_output.device) if ws_bytes else None
    with torch.cuda.device(grad_output.device):
        _lib.check(lib.b200_roi_align_backward_ws(grad_output.data_ptr(), spatial_scale, N, rois.size(0), H, W, C,
                                                  aligned_height, aligned_width, sampling_ratio, rois.data_ptr(),
                                                  grad_input.data_ptr(), ws.data_ptr() if ws is not None else None,
                                                  ws_bytes, _stream()),
                   "b200_roi_align_backward_ws")
    return grad_input


class _RoIAlign(Function):
    @staticmethod
    def forward(ctx, features, rois, aligned_height, aligned_width, spatial_scale, sampling_ratio):
        ctx.args = (int(aligned_height), int(aligned_width), float(spatial_scale), int(sampling_ratio))
        ctx.feature_size = tuple(features.shape)
        ctx.save_for_backward(rois)
        return roi_align_forward(features, rois, *ctx.args)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (rois,) = ctx.saved_tensors
        assert grad_output.is_cuda
        return roi_align_backward(grad_output, rois, ctx.feature_size, *ctx.args), None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# RoIAlign (legacy, lattice-corner samples)
# ------------------------------------------------------------------------------------------------
def roi_align_legacy_forward(features, rois, aligned_height, aligned_width, spatial_scale):
    _need_cuda_f32(features, "features"); _need_cuda_f32(rois, "rois"); _rois_ok(rois)
    features = features.contiguous(); rois = rois.contiguous()
    N, C, H, W = features.shape
    R = rois.size(0)
    out = torch.empty((R, C, aligned_height, aligned_width), dtype=torch.float32, device=features.device)
    with torch.cuda.device(features.device):
        _lib.check(_lib.load().b200_roi_align_legacy_forward(features.data_ptr(), spatial_scale, N, R, H, W, C,
                                                             aligned_height, aligned_width, rois.data_ptr(),
                                                             out.data_ptr(), _stream()),
                   "b200_roi_align_legacy_forward")
    return out


def roi_align_legacy_backward(grad_output, rois, feature_size, aligned_height, aligned_width, spatial_scale):
    _need_cuda_f32(grad_output, "grad_output")
    grad_output = grad_output.contiguous(); rois = rois.contiguous()
    N, C, H, W = feature_size
    grad_input = torch.empty((N, C, H, W), dtype=torch.float32, device=grad_output.device)
    with torch.cuda.device(grad_output.device):
        _lib.check(_lib.load().b200_roi_align_legacy_backward(grad_output.data_ptr(), spatial_scale, N, rois.size(0), H, W, C,
                                                              aligned_height, aligned_width, rois.data_ptr(),
                                                              grad_input.data_ptr(), _stream()),
                   "b200_roi_align_legacy_backward")
    return grad_input


class _RoIAlignLegacy(Function):
    @staticmethod
    def forward(ctx, features, rois, aligned_height, aligned_width, spatial_scale):
        ctx.args = (int(aligned_height), int(aligned_width), float(spatial_scale))
        ctx.feature_size = tuple(features.shape)
        ctx.save_for_backward(rois)
        return roi_align_legacy_forward(features, rois, *ctx.args)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (rois,) = ctx.saved_tensors
        assert grad_output.is_cuda
        return roi_align_legacy_backward(grad_output, rois, ctx.feature_size, *ctx.args), None, None, None, None


# ------------------------------------------------------------------------------------------------
# RoIPool
# ------------------------------------------------------------------------------------------------
def roi_pool_forward(features, rois, pooled_height, pooled_width, spatial_scale):
    _need_cuda_f32(features, "features"); _need_cuda_f32(rois, "rois"); _rois_ok(rois)
    features = features.contiguous(); rois = rois.contiguous()
    N, C, H, W = features.shape
    R = rois.size(0)
    out = torch.empty((R, C, pooled_height, pooled_width), dtype=torch.float32, device=features.device)
    argmax = torch.empty((R, C, pooled_height, pooled_width), dtype=torch.int32, device=features.device)
    with torch.cuda.device(features.device):
        _lib.check(_lib.load().b200_roi_pool_forward(features.data_ptr(), spatial_scale, N, R, H, W, C, pooled_height,
                                                     pooled_width, rois.data_ptr(), out.data_ptr(), argmax.data_ptr(),
                                                     _stream()),
                   "b200_roi_pool_forward")
    return out, argmax


def roi_pool_backward(grad_output, argmax, rois, feature_size, pooled_height, pooled_width, spatial_scale):
    _need_cuda_f32(grad_output, "grad_output")
    grad_output = grad_output.contiguous(); rois = rois.contiguous(); argmax = argmax.contiguous()
    N, C, H, W = feature_size
    grad_input = torch.empty((N, C, H, W), dtype=torch.float32, device=grad_output.device)
    with torch.cuda.device(grad_output.device):
        _lib.check(_lib.load().b200_roi_pool_backward(grad_output.data_ptr(), spatial_scale, N, rois.size(0), H, W, C,
                                                      pooled_height, pooled_width, rois.data_ptr(), grad_input.data_ptr(),
                                                      argmax.data_ptr(), _stream()),
                   "b200_roi_pool_backward")
    return grad_input


class _RoIPool(Function):
    @staticmethod
    def forward(ctx, features, rois, pooled_height, pooled_width, spatial_scale):
        ctx.args = (int(pooled_height), int(pooled_width), float(spatial_scale))
        ctx.feature_size = tuple(features.shape)
        out, argmax = roi_pool_forward(features, rois, *ctx.args)
        ctx.save_for_backward(rois, argmax)
        ctx.mark_non_differentiable(argmax)
        return out, argmax

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output, _grad_argmax):
        rois, argmax = ctx.saved_tensors
        assert grad_output.is_cuda
        return roi_pool_backward(grad_output, argmax, rois, ctx.feature_size, *ctx.args), None, None, None, None


# ------------------------------------------------------------------------------------------------
# RoICrop
# ------------------------------------------------------------------------------------------------
def roi_crop_forward(input1, input2):
    _need_cuda_f32(input1, "input1"); _need_cuda_f32(input2, "input2")
    if input2.dim() != 4 or input2.size(3) != 2:
        raise ValueError("input2 (grid) must be (R, h, w, 2) in (y, x) order, got %s" % (tuple(input2.shape),))
    assert input1.get_device() == input2.get_device(), "input1 and input2 must on the same device"
    input1 = input1.contiguous(); input2 = input2.contiguous()
    N, C, H, W = input1.shape
    R, oh, ow, _ = input2.shape
    out = torch.empty((R, C, oh, ow), dtype=torch.float32, device=input1.device)
    with torch.cuda.device(input1.device):
        _lib.check(_lib.load().b200_roi_crop_forward(input1.data_ptr(), input2.data_ptr(), N, C, H, W, R, oh, ow,
                                                     out.data_ptr(), _stream()),
                   "b200_roi_crop_forward")
    return out


def roi_crop_backward(grad_output, input2, input1_size):
    _need_cuda_f32(grad_output, "grad_output")
    grad_output = grad_output.contiguous(); input2 = input2.contiguous()
    N, C, H, W = input1_size
    R, oh, ow, _ = input2.shape
    grad_input1 = torch.empty((N, C, H, W), dtype=torch.float32, device=grad_output.device)
    grad_input2 = torch.empty_like(input2)
    with torch.cuda.device(grad_output.device):
        _lib.check(_lib.load().b200_roi_crop_backward(grad_output.data_ptr(), input2.data_ptr(), N, C, H, W, R, oh, ow,
                                                      grad_input1.data_ptr(), grad_input2.data_ptr(), _stream()),
                   "b200_roi_crop_backward")
    return grad_input1, grad_input2


class _RoICrop(Function):
    @staticmethod
    def forward(ctx, input1, input2):
        ctx.input1_size = tuple(input1.shape)
        ctx.save_for_backward(input2)
        return roi_crop_forward(input1, input2)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (input2,) = ctx.saved_tensors
        assert grad_output.is_cuda
        return roi_crop_backward(grad_output, input2, ctx.input1_size)


# ------------------------------------------------------------------------------------------------
# NMS
# ------------------------------------------------------------------------------------------------
def nms_raw(dets, thresh):
    """Returns (keep int32 (N,), num_out int32 (1,)) on the device, no host sync."""
    _need_cuda_f32(dets, "dets")
    if dets.dim() != 2 or dets.size(1) < 4:
        raise ValueError("dets must be (N, >=4) = [x1, y1, x2, y2, score], got %s" % (tuple(dets.shape),))
    dets = dets.contiguous()
    n, dim = dets.shape
    lib = _lib.load()
    keep = torch.empty((max(n, 1),), dtype=torch.int32, device=dets.device)
    num_out = torch.empty((1,), dtype=torch.int32, device=dets.device)
    ws_bytes = int(lib.b200_nms_workspace_bytes(n))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dets.device)
    with torch.cuda.device(dets.device):
        _lib.check(lib.b200_nms(dets.data_ptr(), n, dim, float(thresh), keep.data_ptr(), num_out.data_ptr(),
                                ws.data_ptr(), ws_bytes, _stream()),
                   "b200_nms")
    return keep, num_out
