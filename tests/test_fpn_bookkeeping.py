"""FPN RoI bookkeeping: level assignment, collect, distribute -- torch implementation (device
agnostic, exercised on CPU here) vs oracle/fpn.py, which is pinned against the reference's own utils.fpn function."""
import os

import numpy as np
import pytest
import torch

from detectron.pytorch_b200 import synthetic as S
from detectron.pytorch_b200.modeling.collect_and_distribute_fpn_rpn_proposals import (
    CollectAndDistributeFpnRpnProposalsOp, collect, distribute)
from detectron.pytorch_b200.utils.fpn import map_rois_to_fpn_levels
from oracle import fpn as OF

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "proposals.npz"))
CASES = sorted({k.split("/")[1] for k in GOLD.files if k.startswith("fpn_levels/")})


@pytest.mark.parametrize("case", CASES)
def test_level_assignment_matches_the_reference_function(case):
    rois = GOLD["fpn_levels/%s/rois" % case]; k_min, k_max = GOLD["fpn_levels/%s/k" % case]
    ref = GOLD["fpn_levels/%s/lvls" % case]
    assert np.array_equal(OF.map_rois_to_fpn_levels(rois.copy(), k_min, k_max), ref)                      # oracle == reference
    ours = map_rois_to_fpn_levels(torch.from_numpy(rois), int(k_min), int(k_max)).numpy()
    assert np.array_equal(ours, ref)                                                                        # torch == reference
    assert ref[2] == k_min and ref[6] == k_max and ref.min() >= k_min and ref.max() <= k_max              # negative area, huge box


def _levels(seed):
    rng = np.random.RandomState(seed)
    rois, scores = [], []
    for lvl in range(2, 7):
        n = int(rng.randint(0, 300))
        r = S.make_rois(n, (2, 8, 200, 336), 1.0 / 4, seed=seed * 10 + lvl).astype(np.float32) if n else np.zeros((0, 5), np.float32)
        rois.append(r); scores.append(rng.permutation(n).astype(np.float32).reshape(-1, 1) / max(n, 1) + lvl * 1e-3)
    return rois, scores


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_collect_and_distribute_match_the_restated_reference(seed):
    rois, scores = _levels(seed)
    top = 500
    ref_rois = OF.collect(rois, scores, top)
    ours = collect([torch.from_numpy(r) for r in rois], [torch.from_numpy(s) for s in scores], top)
    assert np.array_equal(ours.numpy(), ref_rois)
    ref = OF.distribute(ref_rois, 2, 5)
    out = distribute(ours, 2, 5)
    for k, v in ref.items():
        assert np.array_equal(out[k].numpy(), v), k
    assert out["counts"] == [len(ref["rois_fpn%d" % l]) for l in range(2, 6)]
    # the restore index undoes the regrouping (the reference's own sanity check, utils/fpn.py:58-59)
    stacked = torch.cat([out["rois_fpn%d" % l] for l in range(2, 6)], dim=0)
    assert torch.equal(stacked[out["rois_idx_restore_int32"].long()], ours)
    # brute force: every RoI sits in the blob of its level
    lv = OF.map_rois_to_fpn_levels(ref_rois[:, 1:5].copy(), 2, 5)
    for l in range(2, 6):
        assert np.array_equal(ref["rois_fpn%d" % l], ref_rois[lv == l])


def test_op_wrapper_inference_path_and_config_lookup():
    rois, scores = _levels(5)
    op = CollectAndDistributeFpnRpnProposalsOp(post_nms_topN=100).eval()
    blobs = op([torch.from_numpy(r) for r in rois] + [torch.from_numpy(s) for s in scores])
    assert blobs["rois"].shape == (100, 5) and sum(blobs["counts"]) == 100
    assert sorted(k for k in blobs if k.startswith("rois_fpn")) == ["rois_fpn2", "rois_fpn3", "rois_fpn4", "rois_fpn5"]
    with pytest.raises(NotImplementedError):
        op.train()([torch.zeros(0, 5)] * 5 + [torch.zeros(0, 1)] * 5)
