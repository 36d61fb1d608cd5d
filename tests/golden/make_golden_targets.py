"""Golden vectors for RoI label / target generation, produced by the UNMODIFIED reference code.

Run in the build container (needs /root/reference; nothing here runs on the GPU box):

    python tests/golden/make_golden_targets.py          # writes tests/golden/targets.npz

The reference functions run as they are: utils.cython_bbox.bbox_overlaps (the .pyx built by make_golden_proposals.build_cython),
datasets.json_dataset._merge_proposal_boxes_into_roidb / _add_class_assignments, roi_data.fast_rcnn._sample_rois (with its
_compute_targets / _expand_bbox_targets).  numpy.random.choice is wrapped -- from outside -- only to RECORD which positions it
picked, so that the device path can be driven with the same draw.  matplotlib and pycocotools (absent here) are imported by
json_dataset.py at module level only; empty stand-in modules are registered from outside.
"""
import os
import sys
import tempfile

import numpy as np
import scipy.sparse

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_proposals as MG  # noqa: E402


def make_case(seed, n_gt, n_prop, num_classes, im_w=1333.0, im_h=800.0):
    rng = np.random.RandomState(seed)
    cx = rng.uniform(0, im_w, n_gt); cy = rng.uniform(0, im_h, n_gt)
    w = np.exp(rng.uniform(np.log(24), np.log(500), n_gt)); h = np.exp(rng.uniform(np.log(24), np.log(500), n_gt))
    gt = np.stack([np.clip(cx - w / 2, 0, im_w - 1), np.clip(cy - h / 2, 0, im_h - 1), np.clip(cx + w / 2, 0, im_w - 1),
                   np.clip(cy + h / 2, 0, im_h - 1)], axis=1).astype(np.float32)
    gt_classes = rng.randint(1, num_classes, n_gt).astype(np.int32)
    # proposals: jittered copies of the ground truth (foreground), random boxes (background), exact copies and far-away boxes
    k = n_prop // 2
    src = rng.randint(0, n_gt, k)
    jit = gt[src] + (rng.standard_normal((k, 4)) * (gt[src, 2:3] - gt[src, 0:1] + 1) * 0.12).astype(np.float32)
    pcx = rng.uniform(0, im_w, n_prop - k); pcy = rng.uniform(0, im_h, n_prop - k)
    pw = np.exp(rng.uniform(np.log(16), np.log(600), n_prop - k)); ph = np.exp(rng.uniform(np.log(16), np.log(600), n_prop - k))
    rnd = np.stack([pcx - pw / 2, pcy - ph / 2, pcx + pw / 2, pcy + ph / 2], axis=1)
    prop = np.concatenate([jit, rnd]).astype(np.float32)
    prop[:, 0::2] = np.clip(prop[:, 0::2], 0, im_w - 1); prop[:, 1::2] = np.clip(prop[:, 1::2], 0, im_h - 1)
    prop[:, 2] = np.maximum(prop[:, 2], prop[:, 0]); prop[:, 3] = np.maximum(prop[:, 3], prop[:, 1])
    prop[0] = gt[0]                                                    # an exact duplicate of a ground-truth box
    prop[1] = (0, 0, 3, 3)                                             # touches nothing
    return gt, gt_classes, prop


def main():
    tmp = tempfile.mkdtemp(prefix="ref_cython_")
    MG.build_cython(tmp)
    cfg = MG.import_reference(tmp)[0]
    import types
    import numpy.random as npr
    # matplotlib / pycocotools are imported by json_dataset at module level and never touched by the two functions used here:
    # empty stand-in modules, installed from outside, let the unmodified file import
    for modname, attrs in (("matplotlib", {"use": lambda *a, **k: None}), ("pycocotools", {}), ("pycocotools.mask", {}),
                           ("pycocotools.coco", {"COCO": object})):
        if modname not in sys.modules:
            m = types.ModuleType(modname)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[modname] = m
    sys.modules["pycocotools"].mask = sys.modules["pycocotools.mask"]
    import datasets.json_dataset as JD
    import roi_data.fast_rcnn as FR
    import utils.boxes as box_utils
    out = {}
    cases = [("a", 3, 9, 400, 81, 512, False), ("b", 4, 1, 120, 21, 64, False), ("c", 5, 30, 2000, 81, 512, True)]
    for name, seed, n_gt, n_prop, num_classes, batch, agnostic in cases:
        gt, gt_classes, prop = make_case(seed, n_gt, n_prop, num_classes)
        cfg.MODEL.NUM_CLASSES = num_classes
        cfg.TRAIN.BATCH_SIZE_PER_IM = batch
        cfg.MODEL.CLS_AGNOSTIC_BBOX_REG = agnostic
        cfg.MODEL.MASK_ON = False; cfg.MODEL.KEYPOINTS_ON = False
        # the roidb entry as json_dataset._add_gt_annotations leaves it for an image with n_gt objects
        gt_overlaps = np.zeros((n_gt, num_classes), np.float32)
        gt_overlaps[np.arange(n_gt), gt_classes] = 1.0
        entry = dict(boxes=gt.copy(), gt_classes=gt_classes.copy(), seg_areas=np.zeros(n_gt, np.float32),
                     gt_overlaps=scipy.sparse.csr_matrix(gt_overlaps), is_crowd=np.zeros(n_gt, bool),
                     box_to_gt_ind_map=np.arange(n_gt, dtype=np.int32))
        JD._merge_proposal_boxes_into_roidb([entry], [prop])
        JD._add_class_assignments([entry])
        picks = []
        orig_choice = npr.choice

        def recording_choice(a, size=None, replace=True, p=None):
            sel = orig_choice(a, size=size, replace=replace, p=p)
            pos = {int(v): i for i, v in enumerate(np.asarray(a))}
            picks.append(np.asarray([pos[int(v)] for v in np.atleast_1d(sel)], np.int64))
            return sel
        npr.seed(seed)
        npr.choice = recording_choice
        FR.npr.choice = recording_choice
        try:
            blobs = FR._sample_rois(entry, np.float32(1.5), 1)
        finally:
            npr.choice = orig_choice; FR.npr.choice = orig_choice
        assert len(picks) == 2
        out[name + "_gt"] = gt; out[name + "_gt_classes"] = gt_classes; out[name + "_prop"] = prop
        out[name + "_cfg"] = np.asarray([num_classes, batch, int(agnostic)], np.int64)
        out[name + "_overlaps"] = box_utils.bbox_overlaps(prop, gt)
        out[name + "_max_overlaps"] = entry["max_overlaps"].astype(np.float32)
        out[name + "_max_classes"] = entry["max_classes"].astype(np.int32)
        out[name + "_box_to_gt"] = entry["box_to_gt_ind_map"].astype(np.int32)
        out[name + "_fg_choice"] = picks[0]; out[name + "_bg_choice"] = picks[1]
        for k in ("labels_int32", "rois", "bbox_targets", "bbox_inside_weights", "bbox_outside_weights"):
            out[name + "_" + k] = np.asarray(blobs[k])
        print(name, "boxes", entry["boxes"].shape, "fg", picks[0].size, "bg", picks[1].size, "rois dtype", blobs["rois"].dtype)
    np.savez_compressed(os.path.join(HERE, "targets.npz"), **out)
    print("wrote", os.path.join(HERE, "targets.npz"))


if __name__ == "__main__" and "--train-branch" not in sys.argv:
    main()


def training_branch_golden():
    """The reference's CollectAndDistributeFpnRpnProposalsOp.forward in TRAINING mode (lib/modeling/
    collect_and_distribute_fpn_rpn_proposals.py:42-63), run unmodified on CPU: two images, five RPN levels of synthetic
    proposals, FPN multilevel rois.  Writes tests/golden/collect_train.npz."""
    tmp = tempfile.mkdtemp(prefix="ref_cython_")
    MG.build_cython(tmp)
    cfg = MG.import_reference(tmp)[0]
    import types
    import numpy.random as npr
    import torch
    for modname, attrs in (("matplotlib", {"use": lambda *a, **k: None}), ("pycocotools", {}), ("pycocotools.mask", {}),
                           ("pycocotools.coco", {"COCO": object})):
        if modname not in sys.modules:
            m = types.ModuleType(modname)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[modname] = m
    sys.modules["pycocotools"].mask = sys.modules["pycocotools.mask"]
    import modeling.collect_and_distribute_fpn_rpn_proposals as CD
    import roi_data.fast_rcnn as FR
    cfg.FPN.FPN_ON = True; cfg.FPN.MULTILEVEL_ROIS = True; cfg.FPN.MULTILEVEL_RPN = True
    cfg.MODEL.NUM_CLASSES = 81; cfg.MODEL.MASK_ON = False; cfg.MODEL.KEYPOINTS_ON = False; cfg.MODEL.CLS_AGNOSTIC_BBOX_REG = False
    cfg.TRAIN.BATCH_SIZE_PER_IM = 128; cfg.TRAIN.RPN_POST_NMS_TOP_N = 600
    out = {}
    scales = np.asarray([1.5, 1.25], np.float32)
    im_info = torch.from_numpy(np.stack([[800, 1216, scales[0]], [768, 1333, scales[1]]]).astype(np.float32))
    roidb, gts = [], []
    rng = np.random.RandomState(11)
    roi_inputs, score_inputs = [], []
    all_props = []
    for i in range(2):
        gt, gt_classes, prop = make_case(20 + i, 6 + i, 500, 81, im_w=800.0, im_h=520.0)
        gts.append((gt, gt_classes))
        n_gt = gt.shape[0]
        gt_overlaps = np.zeros((n_gt, 81), np.float32); gt_overlaps[np.arange(n_gt), gt_classes] = 1.0
        roidb.append(dict(boxes=gt.copy(), gt_classes=gt_classes.copy(), seg_areas=np.zeros(n_gt, np.float32),
                          gt_overlaps=scipy.sparse.csr_matrix(gt_overlaps), is_crowd=np.zeros(n_gt, bool),
                          box_to_gt_ind_map=np.arange(n_gt, dtype=np.int32)))
        all_props.append(np.hstack([np.full((prop.shape[0], 1), i, np.float32), prop * scales[i]]).astype(np.float32))
    props = np.concatenate(all_props)
    perm = rng.permutation(props.shape[0])
    props = props[perm]
    scores = ((rng.permutation(props.shape[0]) + 0.5) / props.shape[0]).astype(np.float32)     # unique: no tie-order question
    bounds = np.linspace(0, props.shape[0], 6).astype(int)
    for l in range(5):
        roi_inputs.append(props[bounds[l]:bounds[l + 1]]); score_inputs.append(scores[bounds[l]:bounds[l + 1]].reshape(-1, 1))
    picks = []
    orig_choice = npr.choice

    def recording_choice(a, size=None, replace=True, p=None):
        sel = orig_choice(a, size=size, replace=replace, p=p)
        pos = {int(v): i for i, v in enumerate(np.asarray(a))}
        picks.append(np.asarray([pos[int(v)] for v in np.atleast_1d(sel)], np.int64))
        return sel
    op = CD.CollectAndDistributeFpnRpnProposalsOp()
    op.train()
    npr.seed(5)
    FR.npr.choice = recording_choice
    try:
        blobs = op.forward(roi_inputs + score_inputs, roidb, im_info)
    finally:
        FR.npr.choice = orig_choice
    assert len(picks) == 4
    for l in range(5):
        out["rpn_rois_%d" % l] = roi_inputs[l]; out["rpn_probs_%d" % l] = score_inputs[l]
    for i in range(2):
        out["gt_%d" % i] = gts[i][0]; out["gt_classes_%d" % i] = gts[i][1]
        out["fg_choice_%d" % i] = picks[2 * i]; out["bg_choice_%d" % i] = picks[2 * i + 1]
    out["im_info"] = im_info.numpy()
    for k, v in blobs.items():
        out["blob_" + k] = np.asarray(v)
    print("training branch blobs:", {k: np.asarray(v).shape for k, v in blobs.items()})
    np.savez_compressed(os.path.join(HERE, "collect_train.npz"), **out)


if __name__ == "__main__" and "--train-branch" in sys.argv:
    training_branch_golden()
