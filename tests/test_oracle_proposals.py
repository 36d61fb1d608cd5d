"""oracle/proposals.py against golden vectors produced by the reference's own GenerateProposalsOp (CPU, unmodified;
tests/golden/make_golden_proposals.py), and oracle_nms_cython against the reference's live Cython NMS."""
import os

import numpy as np
import pytest

from detectron.pytorch_b200 import synthetic as S
from oracle import cpu as O
from oracle import proposals as P

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "proposals.npz"))
CASES = sorted({k.split("/")[0] for k in GOLD.files if k.endswith("/params")})


def case(name):
    g = {k.split("/", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith(name + "/")}
    stride, pre, post, thresh, min_size = g["params"]
    return g, float(stride), int(pre), int(post), float(thresh), float(min_size)


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_reference_op_bit_for_bit(name):
    g, stride, pre, post, thresh, min_size = case(name)
    rois, probs = P.generate_proposals(g["scores"], g["deltas"], g["im_info"], g["anchors"], stride, pre, post, thresh,
                                       min_size, nms="cython")
    assert rois.shape == g["rois"].shape
    assert np.array_equal(rois, g["rois"])
    assert np.array_equal(probs, g["probs"])


@pytest.mark.parametrize("name", CASES)
def test_cuda_nms_semantics_give_the_same_proposals_on_these_inputs(name):
    """IoU > t (CUDA kernel) vs IoU >= t (Cython): identical unless some pair sits exactly on the threshold."""
    g, stride, pre, post, thresh, min_size = case(name)
    rois, probs = P.generate_proposals(g["scores"], g["deltas"], g["im_info"], g["anchors"], stride, pre, post, thresh,
                                       min_size, nms="cuda")
    assert np.array_equal(rois, g["rois"]) and np.array_equal(probs, g["probs"])


@pytest.mark.parametrize("n", [1, 65, 1000, 3000])
def test_nms_cython_restatement_matches_the_live_cython_routine(n):
    b = S.make_nms_boxes(n, seed=n)
    assert np.array_equal(np.asarray(O.nms_cython(b, 0.7)).reshape(-1), GOLD["cython_nms/%d/keep" % n])


def test_proposal_properties():
    g, stride, pre, post, thresh, min_size = case("min_size")
    rois, probs = P.generate_proposals(g["scores"], g["deltas"], g["im_info"], g["anchors"], stride, pre, post, thresh, min_size)
    im_h, im_w, scale = g["im_info"][0]
    assert np.all(rois[:, 1] >= 0) and np.all(rois[:, 3] <= im_w - 1) and np.all(rois[:, 2] >= 0) and np.all(rois[:, 4] <= im_h - 1)
    assert np.all(rois[:, 3] - rois[:, 1] + 1 >= min_size * scale) and np.all(rois[:, 4] - rois[:, 2] + 1 >= min_size * scale)
    assert np.all(np.diff(probs.reshape(-1)) <= 0) and len(rois) <= post
