"""torch front-end of oracle/_ref/*.so -- the REFERENCE's own CUDA kernels, compiled unmodified for
sm_100a by oracle/Makefile (`make ref`).  TEST / BASELINE INFRASTRUCTURE ONLY.

The THC/cffi glue of the reference cannot build on torch 2.11, so these wrappers call the
reference's `extern "C"` raw-pointer launchers directly (the layer right under that glue) and
reproduce what the reference's Python did around them (zero-filled outputs,
functions/roi_align.py:23,39-40; keep[:num_out], nms_gpu.py:8-12).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_libs = {}

_vp = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float


def available():
    return all(os.path.exists(os.path.join(_REF, "libref_%s.so" % n))
               for n in ("roi_align_xfrom", "roi_align_legacy", "roi_pooling", "roi_crop", "nms"))


def _lib(name):
    if name not in _libs:
        _libs[name] = ctypes.CDLL(os.path.join(_REF, "libref_%s.so" % name))
    return _libs[name]


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def roi_align_forward(features, rois, PH, PW, scale, sr):
    """ROIAlignForwardLaucher, lib/modeling/roi_xfrom/roi_align/src/roi_align_kernel.cu:123-142"""
    fn = _lib("roi_align_xfrom").ROIAlignForwardLaucher
    fn.argtypes = [_vp, _f, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]
    fn.restype = _i
    N, C, H, W = features.shape
    R = rois.shape[0]
    out = features.new_zeros((R, C, PH, PW))
    if R * C > 0:
        fn(features.data_ptr(), scale, R, H, W, C, PH, PW, sr, rois.data_ptr(), out.data_ptr(), _stream())
    return out


def roi_align_backward(grad_out, rois, feature_size, PH, PW, scale, sr):
    """ROIAlignBackwardLaucher :272-290 (+ the Python-side zero fill)"""
    fn = _lib("roi_align_xfrom").ROIAlignBackwardLaucher
    fn.argtypes = [_vp, _f, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]
    fn.restype = _i
    N, C, H, W = feature_size
    R = rois.shape[0]
    gin = grad_out.new_zeros((N, C, H, W))
    if R * C > 0:
        fn(grad_out.data_ptr(), scale, N, R, H, W, C, PH, PW, sr, rois.data_ptr(), gin.data_ptr(), _stream())
    return gin


def roi_align_legacy_forward(features, rois, PH, PW, scale):
    """ROIAlignForwardLaucher, lib/model/roi_align/src/roi_align_kernel.cu:73-91"""
    fn = _lib("roi_align_legacy").ROIAlignForwardLaucher
    fn.argtypes = [_vp, _f, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]
    fn.restype = _i
    N, C, H, W = features.shape
    R = rois.shape[0]
    out = features.new_zeros((R, C, PH, PW))
    if R * C > 0:
        fn(features.data_ptr(), scale, R, H, W, C, PH, PW, rois.data_ptr(), out.data_ptr(), _stream())
    return out


def roi_align_legacy_backward(grad_out, rois, feature_size, PH, PW, scale):
    fn = _lib("roi_align_legacy").ROIAlignBackwardLaucher
    fn.argtypes = [_vp, _f, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]
    fn.restype = _i
    N, C, H, W = feature_size
    R = rois.shape[0]
    gin = grad_out.new_zeros((N, C, H, W))
    if R * C > 0:
        fn(grad_out.data_ptr(), scale, N, R, H, W, C, PH, PW, rois.data_ptr(), gin.data_ptr(), _stream())
    return gin


def roi_pool_forward(features, rois, PH, PW, scale):
    """ROIPoolForwardLaucher, lib/model/roi_pooling/src/roi_pooling_kernel.cu:95-125"""
    fn = _lib("roi_pooling").ROIPoolForwardLaucher
    fn.argtypes = [_vp, _f, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]
    fn.restype = _i
    N, C, H, W = features.shape
    R = rois.shape[0]
    out = features.new_zeros((R, C, PH, PW))
    argmax = torch.zeros((R, C, PH, PW), dtype=torch.int32, device=features.device)
    if R * C > 0:
        fn(features.data_ptr(), scale, R, H, W, C, PH, PW, rois.data_ptr(), out.data_ptr(), argmax.data_ptr(), _stream())
    return out, argmax


def roi_pool_backward(grad_out, argmax, rois, feature_size, PH, PW, scale):
    """ROIPoolBackwardLaucher :205-234"""
    fn = _lib("roi_pooling").ROIPoolBackwardLaucher
    fn.argtypes = [_vp, _f, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]
    fn.restype = _i
    N, C, H, W = feature_size
    R = rois.shape[0]
    gin = grad_out.new_zeros((N, C, H, W))
    fn(grad_out.data_ptr(), scale, N, R, H, W, C, PH, PW, rois.data_ptr(), gin.data_ptr(), argmax.data_ptr(), _stream())
    return gin


def _strides4(t):
    return [int(s) for s in t.stride()]


def roi_crop_forward(img, grid):
    """BilinearSamplerBHWD_updateOutput_cuda_kernel, lib/model/roi_crop/src/roi_crop_cuda_kernel.cu:201-255,
    called with the argument order of src/roi_crop_cuda.c:23-46."""
    fn = _lib("roi_crop").BilinearSamplerBHWD_updateOutput_cuda_kernel
    fn.argtypes = [_i] * 8 + [_vp] + [_i] * 4 + [_vp] + [_i] * 4 + [_vp] + [_i] * 4 + [_vp]
    fn.restype = _i
    N, C, H, W = img.shape
    R, oh, ow, _ = grid.shape
    out = img.new_zeros((R, C, oh, ow))
    s_i, s_g, s_o = _strides4(img), _strides4(grid), _strides4(out)
    fn(C, ow, oh, R, C, H, W, N,
       img.data_ptr(), s_i[0], s_i[1], s_i[2], s_i[3],
       grid.data_ptr(), s_g[0], s_g[3], s_g[1], s_g[2],
       out.data_ptr(), s_o[0], s_o[1], s_o[2], s_o[3], _stream())
    return out


def roi_crop_backward(img, grid, grad_out):
    """BilinearSamplerBHWD_updateGradInput_cuda_kernel :257-326 (argument order of roi_crop_cuda.c:64-98)"""
    fn = _lib("roi_crop").BilinearSamplerBHWD_updateGradInput_cuda_kernel
    fn.argtypes = [_i] * 8 + ([_vp] + [_i] * 4) * 5 + [_vp]
    fn.restype = _i
    N, C, H, W = img.shape
    R, oh, ow, _ = grid.shape
    gimg = torch.zeros_like(img)
    ggrid = torch.zeros_like(grid)
    s_i, s_g, s_gi, s_gg, s_go = _strides4(img), _strides4(grid), _strides4(gimg), _strides4(ggrid), _strides4(grad_out)
    fn(C, ow, oh, R, C, H, W, N,
       img.data_ptr(), s_i[0], s_i[1], s_i[2], s_i[3],
       grid.data_ptr(), s_g[0], s_g[3], s_g[1], s_g[2],
       gimg.data_ptr(), s_gi[0], s_gi[1], s_gi[2], s_gi[3],
       ggrid.data_ptr(), s_gg[0], s_gg[3], s_gg[1], s_gg[2],
       grad_out.data_ptr(), s_go[0], s_go[1], s_go[2], s_go[3], _stream())
    return gimg, ggrid


def nms_gpu(dets, thresh):
    """nms_cuda_compute, lib/model/nms/src/nms_cuda_kernel.cu:87-161, wrapped like lib/model/nms/nms_gpu.py:7-12.
    NOTE: the reference runs on the legacy default stream with blocking copies."""
    fn = _lib("nms").nms_cuda_compute
    fn.argtypes = [_vp, _vp, _vp, _i, _i, _f]
    fn.restype = None
    n = dets.shape[0]
    keep = torch.zeros((n, 1), dtype=torch.int32, device=dets.device)
    num_out = torch.zeros((1,), dtype=torch.int32, device=dets.device)
    torch.cuda.current_stream().synchronize()
    fn(keep.data_ptr(), num_out.data_ptr(), dets.data_ptr(), n, dets.shape[1], thresh)
    return keep[:int(num_out[0])]
