"""numpy restatement of the FPN RoI bookkeeping of the reference.  TEST INFRASTRUCTURE ONLY.

  * map_rois_to_fpn_levels      lib/utils/fpn.py:11-28, boxes_area lib/utils/boxes.py:58-69
  * collect                     lib/modeling/collect_and_distribute_fpn_rpn_proposals.py:72-87 (stable sort: the
                                reference's np.argsort(-scores) leaves the order of equal scores unspecified)
  * distribute                  same file :90-119
map_rois_to_fpn_levels is pinned against the reference's own function (tests/golden/proposals.npz, key
`fpn_levels/*`, written by tests/golden/make_golden_proposals.py); collect / distribute live in a module that cannot
be imported here (it pulls in pycocotools), so they are restated and checked against brute force in the tests.
"""
import numpy as np


def map_rois_to_fpn_levels(rois, k_min, k_max, s0=224, lvl0=4):
    w = rois[:, 2] - rois[:, 0] + 1
    h = rois[:, 3] - rois[:, 1] + 1
    areas = w * h
    areas[np.where(areas < 0)[0]] = 0
    s = np.sqrt(areas)
    return np.clip(np.floor(lvl0 + np.log2(s / s0 + 1e-6)), k_min, k_max)


def collect(roi_inputs, score_inputs, post_nms_topN):
    rois = np.concatenate(roi_inputs)
    scores = np.concatenate(score_inputs).reshape(-1)
    inds = np.argsort(-scores, kind="stable")[:post_nms_topN]
    return rois[inds, :]


def distribute(rois, lvl_min, lvl_max, s0=224, lvl0=4):
    lvls = map_rois_to_fpn_levels(rois[:, 1:5].copy(), lvl_min, lvl_max, s0, lvl0)
    out = {"rois": rois}
    order = np.empty((0,))
    for lvl in range(lvl_min, lvl_max + 1):
        idx = np.where(lvls == lvl)[0]
        out["rois_fpn%d" % lvl] = rois[idx, :]
        order = np.concatenate((order, idx))
    out["rois_idx_restore_int32"] = np.argsort(order).astype(np.int32)
    return out
