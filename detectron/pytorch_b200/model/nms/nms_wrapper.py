"""Drop-in for `model.nms.nms_wrapper` (reference lib/model/nms/nms_wrapper.py:11-18)."""
from .nms_gpu import nms_gpu


def nms(dets, thresh, force_cpu=False):
    """Dispatch to either CPU or GPU NMS implementations (only the GPU one exists, as in the reference)."""
    if dets.shape[0] == 0:
        return []
    return nms_gpu(dets, thresh)
