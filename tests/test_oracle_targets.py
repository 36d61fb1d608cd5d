"""CPU: the numpy restatement of the reference's RoI label / target generation (oracle/targets.py) against the golden vectors
the UNMODIFIED reference functions produced (tests/golden/make_golden_targets.py -> targets.npz)."""
import os

import numpy as np
import pytest

from oracle import targets as OT

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "targets.npz"))
CASES = ["a", "b", "c"]


def roidb_boxes(c):
    return np.concatenate([GOLD[c + "_gt"], GOLD[c + "_prop"]]).astype(np.float32)


@pytest.mark.parametrize("c", CASES)
def test_bbox_overlaps_bit_exact(c):
    got = OT.bbox_overlaps(GOLD[c + "_prop"], GOLD[c + "_gt"])
    assert got.dtype == np.float32 and np.array_equal(got, GOLD[c + "_overlaps"])


@pytest.mark.parametrize("c", CASES)
def test_assignment_matches_roidb(c):
    mo, am, mc = OT.assign_rois(roidb_boxes(c), GOLD[c + "_gt"], GOLD[c + "_gt_classes"])
    assert np.array_equal(mo, GOLD[c + "_max_overlaps"])
    assert np.array_equal(mc, GOLD[c + "_max_classes"])
    assert np.array_equal(am, GOLD[c + "_box_to_gt"])


@pytest.mark.parametrize("c", CASES)
def test_sample_rois_matches_reference_blobs(c):
    ncls, batch, agn = [int(v) for v in GOLD[c + "_cfg"]]
    b = OT.sample_rois(roidb_boxes(c), GOLD[c + "_gt"], GOLD[c + "_gt_classes"], np.float32(1.5), 1, ncls, GOLD[c + "_fg_choice"],
                       GOLD[c + "_bg_choice"], batch_size_per_im=batch, cls_agnostic_bbox_reg=bool(agn))
    assert np.array_equal(b["labels_int32"], GOLD[c + "_labels_int32"])
    assert np.array_equal(b["rois"], GOLD[c + "_rois"])
    assert np.array_equal(b["bbox_inside_weights"], GOLD[c + "_bbox_inside_weights"])
    assert np.array_equal(b["bbox_outside_weights"], GOLD[c + "_bbox_outside_weights"])
    np.testing.assert_allclose(b["bbox_targets"], GOLD[c + "_bbox_targets"], rtol=0, atol=0)      # same numpy, same logf
