"""Merge the per-level RPN proposals and hand them out to the RoI levels, on tensors (any device) -- the inference
path of `CollectAndDistributeFpnRpnProposalsOp`, lib/modeling/collect_and_distribute_fpn_rpn_proposals.py:10-119
(reference, numpy on the host).  Together with modeling/generate_proposals.py (device proposals) and
roi_align_fpn.RoIAlignFPNFunction (per-level RoIAlign written in restored order) the RoIs no longer leave the GPU
between the RPN heads and the box head; the only host read is the four per-level counts the launches need.

  collect     :72-87    concatenate the levels, keep the `post_nms_topN` best by score
  distribute  :90-119   level of every RoI (utils.fpn.map_rois_to_fpn_levels), per-level RoI blobs in ascending
                        original index, and `rois_idx_restore_int32` = the inverse of that regrouping

Equal scores: the reference's np.argsort(-scores) leaves their order unspecified; here the sort is stable (ties keep
concatenation order).  The training branch (:49-63: labels and regression targets through the data-loader code) is
outside SURVEY.md 8 and raises NotImplementedError.
"""
import torch
from torch import nn

from detectron.pytorch_b200.utils.fpn import map_rois_to_fpn_levels


def collect(roi_inputs, score_inputs, post_nms_topN):
    """roi_inputs: list of (R_l, 5) tensors, score_inputs: list of (R_l, 1) or (R_l,) tensors -> (min(R, topN), 5)."""
    rois = torch.cat([r.reshape(-1, 5) for r in roi_inputs], dim=0)
    scores = torch.cat([s.reshape(-1) for s in score_inputs], dim=0)
    order = torch.sort(scores, descending=True, stable=True).indices
    if post_nms_topN >= 0:
        order = order[:int(post_nms_topN)]
    return rois[order]


def distribute(rois, lvl_min, lvl_max, canonical_scale=224, canonical_level=4):
    """rois (R, 5) -> dict with 'rois', 'rois_fpn<lvl>' for lvl_min..lvl_max and 'rois_idx_restore_int32', exactly the
    blobs of the reference's inference path; plus 'counts' (python ints per level, one small host read)."""
    lvls = map_rois_to_fpn_levels(rois[:, 1:5], lvl_min, lvl_max, canonical_scale, canonical_level)
    order = torch.sort(lvls, stable=True).indices                      # level by level, ascending original index inside
    restore = torch.empty_like(order)
    restore[order] = torch.arange(order.numel(), device=order.device)
    counts = torch.bincount((lvls - lvl_min).long(), minlength=lvl_max - lvl_min + 1).tolist()
    out = {"rois": rois, "counts": counts}
    grouped = rois[order]
    start = 0
    for lvl, n in zip(range(lvl_min, lvl_max + 1), counts):
        out["rois_fpn%d" % lvl] = grouped[start:start + n]
        start += n
    out["rois_idx_restore_int32"] = restore.to(torch.int32)
    return out


class CollectAndDistributeFpnRpnProposalsOp(nn.Module):
    def __init__(self, rpn_min_level=2, rpn_max_level=6, roi_min_level=2, roi_max_level=5, post_nms_topN=None,
                 collect_scale=1, canonical_scale=224, canonical_level=4, cfg=None):
        """`cfg`: the reference's config object (FPN.RPN_*_LEVEL, FPN.ROI_*_LEVEL, FPN.RPN_COLLECT_SCALE,
        FPN.ROI_CANONICAL_*, TRAIN/TEST.RPN_POST_NMS_TOP_N are then read at call time); otherwise the keyword values
        (defaults = lib/core/config.py:700-723; post_nms_topN defaults to the TEST value 1000)."""
        super().__init__()
        self._cfg = cfg
        self._p = dict(rpn_min=rpn_min_level, rpn_max=rpn_max_level, roi_min=roi_min_level, roi_max=roi_max_level,
                       topn=1000 if post_nms_topN is None else post_nms_topN, scale=collect_scale, s0=canonical_scale,
                       k0=canonical_level)

    def _params(self):
        if self._cfg is None:
            return dict(self._p)
        c = self._cfg
        key = 'TRAIN' if self.training else 'TEST'
        return dict(rpn_min=c.FPN.RPN_MIN_LEVEL, rpn_max=c.FPN.RPN_MAX_LEVEL, roi_min=c.FPN.ROI_MIN_LEVEL,
                    roi_max=c.FPN.ROI_MAX_LEVEL, topn=c[key].RPN_POST_NMS_TOP_N, scale=c.FPN.RPN_COLLECT_SCALE,
                    s0=c.FPN.ROI_CANONICAL_SCALE, k0=c.FPN.ROI_CANONICAL_LEVEL)

    def forward(self, inputs, roidb=None, im_info=None):
        """inputs: [rpn_rois_fpn<min> .. rpn_rois_fpn<max>, rpn_roi_probs_fpn<min> .. rpn_roi_probs_fpn<max>] tensors."""
        if self.training:
            raise NotImplementedError("training path (labels / targets via the data loader) is outside this package")
        p = self._params()
        n = p["rpn_max"] - p["rpn_min"] + 1
        topn = int(p["topn"] * p["scale"] + 0.5)
        rois = collect(inputs[:n], inputs[n:], topn)
        return distribute(rois, p["roi_min"], p["roi_max"], p["s0"], p["k0"])
