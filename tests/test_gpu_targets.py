"""GPU: RoI label / target generation on the device, through the C ABI, against the golden vectors of the
unmodified reference functions and against the numpy oracle on larger seeded inputs.
Bit-exact: overlaps, assignments, labels, index lists, rois, weights.  Regression targets: dx / dy bit-exact, dw / dh within
2e-6 relative (CUDA logf vs numpy's SIMD float32 log)."""
import os

import numpy as np
import pytest
import torch

from detectron.pytorch_b200.roi_data import fast_rcnn as FR
from oracle import targets as OT

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "targets.npz"))
CASES = ["a", "b", "c"]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def roidb_boxes(c):
    return np.concatenate([GOLD[c + "_gt"], GOLD[c + "_prop"]]).astype(np.float32)


def assert_targets_close(got, ref):
    assert np.array_equal(got == 0, ref == 0)
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-7)                # dw, dh: CUDA logf (1 ulp) vs numpy's SIMD float32 log (~4 ulp)
    cols = np.arange(got.shape[1]) % 4 < 2                                    # dx, dy: no transcendental -> exact
    assert np.array_equal(got[:, cols], ref[:, cols])


@pytest.mark.parametrize("c", CASES)
def test_bbox_overlaps_vs_reference(c):
    got = FR.bbox_overlaps(dev(GOLD[c + "_prop"]), dev(GOLD[c + "_gt"])).cpu().numpy()
    assert np.array_equal(got, GOLD[c + "_overlaps"])


def test_bbox_overlaps_large_vs_oracle_and_empty():
    rng = np.random.RandomState(0)
    a = np.sort(rng.uniform(0, 900, (3000, 2, 2)), axis=1).transpose(0, 2, 1).reshape(3000, 4)[:, [0, 2, 1, 3]].astype(np.float32)
    b = np.sort(rng.uniform(0, 900, (700, 2, 2)), axis=1).transpose(0, 2, 1).reshape(700, 4)[:, [0, 2, 1, 3]].astype(np.float32)
    assert np.array_equal(FR.bbox_overlaps(dev(a), dev(b)).cpu().numpy(), OT.bbox_overlaps(a, b))
    assert tuple(FR.bbox_overlaps(dev(a[:0]), dev(b)).shape) == (0, 700)


@pytest.mark.parametrize("c", CASES)
def test_assignment_vs_reference_roidb(c):
    mo, am, mc = FR.assign_rois(dev(roidb_boxes(c)), dev(GOLD[c + "_gt"]), dev(GOLD[c + "_gt_classes"]))
    assert np.array_equal(mo.cpu().numpy(), GOLD[c + "_max_overlaps"])
    assert np.array_equal(mc.cpu().numpy(), GOLD[c + "_max_classes"])
    assert np.array_equal(am.cpu().numpy(), GOLD[c + "_box_to_gt"])


@pytest.mark.parametrize("c", CASES)
def test_sample_rois_vs_reference_blobs(c):
    ncls, batch, agn = [int(v) for v in GOLD[c + "_cfg"]]
    b = FR.sample_rois(dev(roidb_boxes(c)), dev(GOLD[c + "_gt"]), dev(GOLD[c + "_gt_classes"]), 1.5, 1, ncls, batch_size_per_im=batch,
                       cls_agnostic_bbox_reg=bool(agn), fg_choice=GOLD[c + "_fg_choice"], bg_choice=GOLD[c + "_bg_choice"])
    assert np.array_equal(b["labels_int32"].cpu().numpy(), GOLD[c + "_labels_int32"])
    assert np.array_equal(b["rois"].cpu().numpy(), GOLD[c + "_rois"])
    assert np.array_equal(b["bbox_inside_weights"].cpu().numpy(), GOLD[c + "_bbox_inside_weights"])
    assert np.array_equal(b["bbox_outside_weights"].cpu().numpy(), GOLD[c + "_bbox_outside_weights"])
    assert_targets_close(b["bbox_targets"].cpu().numpy(), GOLD[c + "_bbox_targets"])


def test_sample_rois_device_draw_is_a_valid_minibatch():
    """Without host-provided choices the draw is a torch.randperm on the device: sizes, membership and labels must still be
    those of _sample_rois."""
    c = "c"
    ncls, batch, agn = [int(v) for v in GOLD[c + "_cfg"]]
    boxes = roidb_boxes(c)
    b = FR.sample_rois(dev(boxes), dev(GOLD[c + "_gt"]), dev(GOLD[c + "_gt_classes"]), 1.0, 0, ncls, batch_size_per_im=batch)
    keep = b["keep_inds"].cpu().numpy()
    mo = GOLD[c + "_max_overlaps"]
    n_fg = min(int(round(0.25 * batch)), int((mo >= 0.5).sum()))
    assert keep.size == n_fg + min(batch - n_fg, int(((mo < 0.5) & (mo >= 0)).sum()))
    assert len(set(keep.tolist())) == keep.size
    assert (mo[keep[:n_fg]] >= 0.5).all() and (mo[keep[n_fg:]] < 0.5).all()
    lab = b["labels_int32"].cpu().numpy()
    assert np.array_equal(lab[:n_fg], GOLD[c + "_max_classes"][keep[:n_fg]]) and (lab[n_fg:] == 0).all()


def test_no_ground_truth_and_cpu_tensors():
    boxes = dev(GOLD["a_prop"])
    mo, am, mc = FR.assign_rois(boxes, None, None)
    assert float(mo.abs().sum()) == 0 and int((am != -1).sum()) == 0 and int(mc.abs().sum()) == 0
    with pytest.raises(NotImplementedError):
        FR.bbox_overlaps(boxes.cpu(), boxes.cpu())


def test_collect_and_distribute_training_branch_vs_reference_op():
    """SURVEY 8f N1 remainder: CollectAndDistributeFpnRpnProposalsOp.forward in training mode -- proposals merged into the roidb,
    labels / targets, FPN distribution -- on the device, against the blobs of the unmodified reference op run on CPU
    (tests/golden/make_golden_targets.py --train-branch) with the same random draws."""
    from detectron.pytorch_b200.modeling.collect_and_distribute_fpn_rpn_proposals import CollectAndDistributeFpnRpnProposalsOp
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "collect_train.npz"))
    op = CollectAndDistributeFpnRpnProposalsOp(rpn_min_level=2, rpn_max_level=6, roi_min_level=2, roi_max_level=5, post_nms_topN=600)
    op.train()
    inputs = [dev(g["rpn_rois_%d" % l]) for l in range(5)] + [dev(g["rpn_probs_%d" % l]) for l in range(5)]
    roidb = [dict(boxes=g["gt_%d" % i], gt_classes=g["gt_classes_%d" % i]) for i in range(2)]
    train = dict(BATCH_SIZE_PER_IM=128, FG_FRACTION=0.25, FG_THRESH=0.5, BG_THRESH_HI=0.5, BG_THRESH_LO=0.0, NUM_CLASSES=81,
                 BBOX_REG_WEIGHTS=(10.0, 10.0, 5.0, 5.0), CLS_AGNOSTIC_BBOX_REG=False)
    choices = [(g["fg_choice_%d" % i], g["bg_choice_%d" % i]) for i in range(2)]
    blobs = op(inputs, roidb, g["im_info"], choices=choices, train=train)
    for k in ("labels_int32", "rois", "bbox_inside_weights", "bbox_outside_weights", "rois_fpn2", "rois_fpn3", "rois_fpn4", "rois_fpn5",
              "rois_idx_restore_int32"):
        assert np.array_equal(blobs[k].cpu().numpy(), g["blob_" + k]), k
    assert_targets_close(blobs["bbox_targets"].cpu().numpy(), g["blob_bbox_targets"])
    op.eval()
    assert "labels_int32" not in op(inputs)                                   # inference path unchanged


def test_topk_batched_vs_torch():
    """b200_topk_batched (radix select + in-CTA sort) against torch.topk / a stable sort: several problems of different shapes in
    one call, unique scores (indices must agree), heavy ties (values must agree, equal scores in ascending (h, w, a) index), k = n."""
    from detectron.pytorch_b200 import ops
    rng = np.random.RandomState(0)
    shapes = [(3, 200, 336), (3, 100, 168), (15, 50, 84), (3, 13, 21), (1, 7, 9)]
    ks = [2000, 2000, 12000, 3 * 13 * 21, 10]
    maps = []
    for (A, H, W) in shapes:
        n = A * H * W
        maps.append(dev(((rng.permutation(n) + 0.5) / n).astype(np.float32).reshape(A, H, W) * 2 - 1))       # unique, both signs
    order, top = ops.topk_batched_raw(maps, ks)
    off = 0
    for m, k in zip(maps, ks):
        flat = m.permute(1, 2, 0).reshape(-1)
        ref_v, ref_i = torch.topk(flat, k, largest=True, sorted=True)
        assert torch.equal(top[off:off + k], ref_v)
        assert torch.equal(order[off:off + k], ref_i)
        off += k
    # ties: few distinct values
    tie = dev(rng.randint(0, 7, (3, 40, 50)).astype(np.float32) / 7)
    o, v = ops.topk_batched_raw([tie], [1500])
    flat = tie.permute(1, 2, 0).reshape(-1)
    assert torch.equal(v, torch.topk(flat, 1500).values)
    assert torch.equal(flat[o], v)
    same = v[1:] == v[:-1]
    assert bool((o[1:][same] > o[:-1][same]).all())                   # equal scores: ascending index
    assert o.unique().numel() == 1500
