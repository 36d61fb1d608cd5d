"""Drop-in for `model.nms.nms_gpu` (reference lib/model/nms/nms_gpu.py:7-12).

dets: CUDA fp32 (N,5) = [x1,y1,x2,y2,score], ALREADY sorted by score descending (the function does
not sort and never reads column 4).  Returns the kept row indices, ascending, as a CUDA int32 (K,1)
tensor -- bit-exact with the reference kernel + host scan.  Like the reference, slicing by the
device-side count costs one host sync; `nms_gpu_raw` returns (keep, num_out) without it."""
from detectron.pytorch_b200 import ops as _ops


def nms_gpu_raw(dets, thresh):
    return _ops.nms_raw(dets, thresh)


def nms_gpu(dets, thresh):
    keep, num_out = _ops.nms_raw(dets, thresh)
    keep = keep.view(-1, 1)
    keep = keep[:int(num_out[0])]
    return keep
