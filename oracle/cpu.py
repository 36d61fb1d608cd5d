"""numpy front-end of oracle/liboracle.so (roi_ops_oracle.c).  TEST INFRASTRUCTURE ONLY.

Every function takes/returns C-contiguous numpy arrays and mirrors one reference kernel; see the
file:line citations in roi_ops_oracle.c.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)
_u64p = ctypes.POINTER(ctypes.c_uint64)


def build(force=False):
    src = os.path.join(_HERE, "roi_ops_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_roi_align_touched_cells.restype = ctypes.c_long
        _lib.oracle_nms_cuda.restype = ctypes.c_int
        _lib.oracle_nms_cython.restype = ctypes.c_int
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def set_fused(fused):
    """fused=True (default): nvcc's FMA contraction (== the GPU reference).  False: strict source
    semantics (== torchvision CPU); only for pinning the index logic."""
    lib().oracle_set_fused(int(bool(fused)))


def set_threads(n):
    lib().oracle_set_threads(int(n))


def max_threads():
    return int(lib().oracle_max_threads())


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_i32p)


def roi_align_forward(features, rois, PH, PW, scale, sr):
    features, fp = _f(features)
    rois, rp = _f(rois)
    N, C, H, W = features.shape
    R = rois.shape[0]
    out = np.empty((R, C, PH, PW), np.float32)
    lib().oracle_roi_align_forward(fp, rp, N, C, H, W, R, PH, PW, ctypes.c_float(scale), int(sr),
                                   out.ctypes.data_as(_f32p))
    return out


def roi_align_backward(top_diff, rois, feature_size, PH, PW, scale, sr, acc64=False):
    top_diff, tp = _f(top_diff)
    rois, rp = _f(rois)
    N, C, H, W = feature_size
    R = rois.shape[0]
    out = np.empty((N, C, H, W), np.float32)
    lib().oracle_roi_align_backward(tp, rp, N, C, H, W, R, PH, PW, ctypes.c_float(scale), int(sr),
                                    out.ctypes.data_as(_f32p), int(acc64))
    return out


def roi_align_touched_cells(rois, N, H, W, PH, PW, scale, sr):
    rois, rp = _f(rois)
    return int(lib().oracle_roi_align_touched_cells(rp, N, H, W, rois.shape[0], PH, PW,
                                                    ctypes.c_float(scale), int(sr)))


def roi_align_legacy_forward(features, rois, PH, PW, scale):
    features, fp = _f(features)
    rois, rp = _f(rois)
    N, C, H, W = features.shape
    R = rois.shape[0]
    out = np.empty((R, C, PH, PW), np.float32)
    lib().oracle_roi_align_legacy_forward(fp, rp, N, C, H, W, R, PH, PW, ctypes.c_float(scale),
                                          out.ctypes.data_as(_f32p))
    return out


def roi_align_legacy_backward(top_diff, rois, feature_size, PH, PW, scale, acc64=False):
    top_diff, tp = _f(top_diff)
    rois, rp = _f(rois)
    N, C, H, W = feature_size
    out = np.empty((N, C, H, W), np.float32)
    lib().oracle_roi_align_legacy_backward(tp, rp, N, C, H, W, rois.shape[0], PH, PW, ctypes.c_float(scale),
                                           out.ctypes.data_as(_f32p), int(acc64))
    return out


def roi_pool_forward(features, rois, PH, PW, scale):
    features, fp = _f(features)
    rois, rp = _f(rois)
    N, C, H, W = features.shape
    R = rois.shape[0]
    out = np.empty((R, C, PH, PW), np.float32)
    argmax = np.empty((R, C, PH, PW), np.int32)
    lib().oracle_roi_pool_forward(fp, rp, N, C, H, W, R, PH, PW, ctypes.c_float(scale),
                                  out.ctypes.data_as(_f32p), argmax.ctypes.data_as(_i32p))
    return out, argmax


def roi_pool_backward(top_diff, argmax, rois, feature_size, PH, PW, scale):
    top_diff, tp = _f(top_diff)
    argmax, ap = _i(argmax)
    rois, rp = _f(rois)
    N, C, H, W = feature_size
    out = np.empty((N, C, H, W), np.float32)
    lib().oracle_roi_pool_backward(tp, ap, rp, N, C, H, W, rois.shape[0], PH, PW, ctypes.c_float(scale),
                                   out.ctypes.data_as(_f32p))
    return out


def roi_crop_forward(img, grid):
    img, ip = _f(img)
    grid, gp = _f(grid)
    N, C, H, W = img.shape
    R, oh, ow, two = grid.shape
    assert two == 2
    out = np.empty((R, C, oh, ow), np.float32)
    lib().oracle_roi_crop_forward(ip, gp, N, C, H, W, R, oh, ow, out.ctypes.data_as(_f32p))
    return out


def roi_crop_backward(grad_out, grid, img_size, acc64=False):
    grad_out, op = _f(grad_out)
    grid, gp = _f(grid)
    N, C, H, W = img_size
    R, oh, ow, _ = grid.shape
    out = np.empty((N, C, H, W), np.float32)
    lib().oracle_roi_crop_backward(op, gp, N, C, H, W, R, oh, ow, out.ctypes.data_as(_f32p), int(acc64))
    return out


def nms_cuda(dets, thresh):
    """CUDA-semantics greedy NMS (bit-exact target). dets (N, >=4) fp32 already sorted by score."""
    dets, dp = _f(dets)
    n = dets.shape[0]
    dim = dets.shape[1] if dets.ndim == 2 else 5
    keep = np.empty((max(n, 1),), np.int32)
    k = lib().oracle_nms_cuda(dp, n, dim, ctypes.c_float(thresh), keep.ctypes.data_as(_i32p))
    return keep[:k].copy()


def nms_cuda_mask(dets, thresh):
    dets, dp = _f(dets)
    n, dim = dets.shape
    cb = (n + 63) // 64
    mask = np.zeros((n, cb), np.uint64)
    lib().oracle_nms_cuda_mask(dp, n, dim, ctypes.c_float(thresh), mask.ctypes.data_as(_u64p))
    return mask


def nms_cython(dets, thresh):
    """The reference's live CPU NMS flavour (lib/utils/cython_nms.pyx:37-87)."""
    dets, dp = _f(dets)
    n = dets.shape[0]
    keep = np.empty((max(n, 1),), np.int32)
    k = lib().oracle_nms_cython(dp, n, ctypes.c_double(thresh), keep.ctypes.data_as(_i32p))
    return keep[:k].astype(np.int64)
