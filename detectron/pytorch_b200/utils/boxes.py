"""Host-array front door of the device NMS, shaped like the reference's `utils.boxes.nms`
(lib/utils/boxes.py:320-324: numpy (N, 5) [x1, y1, x2, y2, score] in any order -> kept indices).

`generate_proposals.py:161` and `core/test.py:764` call `box_utils.nms(dets, thresh)` with numpy arrays; re-pointing
that name at this function (`install_reference_aliases(nms=True)` or `utils.boxes.nms = nms`) moves the O(N^2) work to
`b200_nms` while the callers stay unchanged.  Flavour note: the device kernel keeps the reference CUDA kernel's
`IoU > thresh` test and fused IoU arithmetic (lib/model/nms/src/nms_cuda_kernel.cu:31-39), the Cython routine uses
`>=` (lib/utils/cython_nms.pyx:83-84); they differ only on pairs whose IoU is within one ulp of the threshold.
"""
import numpy as np
import torch

from ..model.nms.nms_gpu import nms_gpu


def nms(dets, thresh):
    """Classic greedy NMS.  Returns the kept row indices in ascending order, like cython_nms.nms
    (`np.where(suppressed == 0)[0]`, lib/utils/cython_nms.pyx:87)."""
    if dets.shape[0] == 0:
        return []
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    order = np.argsort(-dets[:, 4], kind="stable")                  # nms_gpu expects score-sorted rows
    keep = nms_gpu(torch.from_numpy(dets[order]).cuda(), float(thresh)).view(-1).cpu().numpy()
    return np.sort(order[keep]).astype(np.int64)
