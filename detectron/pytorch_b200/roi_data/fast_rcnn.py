"""RoI labels and box-regression targets on the device.

Mirrors, for proposals that already live on the GPU, what the reference does per image on the host:
  * `bbox_overlaps`          lib/utils/cython_bbox.pyx:32-73
  * `assign_rois`            lib/datasets/json_dataset.py:429-490 (_merge_proposal_boxes_into_roidb) + :514-531
  * `sample_rois`            lib/roi_data/fast_rcnn.py:129-200 (_sample_rois), :203-248 (_compute_targets, _expand_bbox_targets)
Mask / keypoint targets (polygon rasterisation through pycocotools) stay on the host as in the reference.

The random choice of the reference (`npr.choice(inds, size, replace=False)` = a permutation prefix) is an INPUT here: pass
the positions it picked (`fg_choice`, `bg_choice`, e.g. from the same numpy RandomState) for a bit-identical minibatch, or
leave them None for a torch.randperm draw on the device.
"""
import ctypes

import torch

from .. import _lib
from ..ops import _need_cuda_f32, _stream


def _boxes4(t, name):
    _need_cuda_f32(t, name)
    if t.dim() != 2 or t.size(1) != 4:
        raise ValueError("%s must be (N, 4) = [x1, y1, x2, y2], got %s" % (name, tuple(t.shape)))
    return t.contiguous()


def bbox_overlaps(boxes, query_boxes):
    """(N, K) float32 IoU, +1 pixel convention -- utils.cython_bbox.bbox_overlaps on CUDA tensors."""
    boxes = _boxes4(boxes, "boxes"); query_boxes = _boxes4(query_boxes, "query_boxes")
    N, K = boxes.size(0), query_boxes.size(0)
    out = torch.empty((N, K), dtype=torch.float32, device=boxes.device)
    with torch.cuda.device(boxes.device):
        _lib.check(_lib.load().b200_bbox_overlaps(boxes.data_ptr(), N, query_boxes.data_ptr(), K, out.data_ptr(), _stream()),
                   "b200_bbox_overlaps")
    return out


def assign_rois(boxes, gt_boxes, gt_classes):
    """Best ground-truth box per box: (max_overlaps float32, box_to_gt_ind int32 (-1: none), max_classes int32 (0: none))."""
    boxes = _boxes4(boxes, "boxes")
    N = boxes.size(0)
    G = int(gt_boxes.size(0)) if gt_boxes is not None else 0
    dev = boxes.device
    max_overlaps = torch.empty((N,), dtype=torch.float32, device=dev)
    argmax = torch.empty((N,), dtype=torch.int32, device=dev)
    max_classes = torch.empty((N,), dtype=torch.int32, device=dev)
    if G:
        gt_boxes = _boxes4(gt_boxes, "gt_boxes")
        gt_classes = gt_classes.to(device=dev, dtype=torch.int32).contiguous()
    with torch.cuda.device(dev):
        _lib.check(_lib.load().b200_roi_assign(boxes.data_ptr(), N, gt_boxes.data_ptr() if G else None,
                                               gt_classes.data_ptr() if G else None, G, max_overlaps.data_ptr(), argmax.data_ptr(),
                                               max_classes.data_ptr(), _stream()), "b200_roi_assign")
    return max_overlaps, argmax, max_classes


def select_fg_bg(max_overlaps, fg_thresh, bg_thresh_hi, bg_thresh_lo):
    """np.where(max_overlaps >= fg) and np.where((max_overlaps < hi) & (max_overlaps >= lo)) as ascending int32 index tensors
    (one host read of the two counts)."""
    max_overlaps = max_overlaps.contiguous()
    N = max_overlaps.numel()
    dev = max_overlaps.device
    fg = torch.empty((N,), dtype=torch.int32, device=dev)
    bg = torch.empty((N,), dtype=torch.int32, device=dev)
    counts = torch.empty((2,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().b200_roi_select(max_overlaps.data_ptr(), N, float(fg_thresh), float(bg_thresh_hi), float(bg_thresh_lo),
                                               fg.data_ptr(), bg.data_ptr(), counts.data_ptr(), _stream()), "b200_roi_select")
    nf, nb = counts.tolist()
    return fg[:nf], bg[:nb]


def sample_rois(boxes, gt_boxes, gt_classes, im_scale, batch_idx, num_classes, batch_size_per_im=512, fg_fraction=0.25,
                fg_thresh=0.5, bg_thresh_hi=0.5, bg_thresh_lo=0.0, bbox_reg_weights=(10.0, 10.0, 5.0, 5.0),
                cls_agnostic_bbox_reg=False, fg_choice=None, bg_choice=None, assignment=None):
    """_sample_rois for one image.  `boxes`: the roidb's boxes (ground truth first, then proposals; unscaled image
    coordinates).  Returns the reference's blob dict with CUDA tensors: labels_int32, rois, bbox_targets,
    bbox_inside_weights, bbox_outside_weights (+ keep_inds, the rows of `boxes` that were sampled).
    assignment: (max_overlaps, box_to_gt_ind, max_classes) when the caller already has them (assign_rois)."""
    boxes = _boxes4(boxes, "boxes")
    dev = boxes.device
    if assignment is None:
        assignment = assign_rois(boxes, gt_boxes, gt_classes)
    max_overlaps, argmax, max_classes = assignment
    rois_per_image = int(batch_size_per_im)
    fg_rois_per_image = int(round(fg_fraction * rois_per_image))
    fg_inds, bg_inds = select_fg_bg(max_overlaps, fg_thresh, bg_thresh_hi, bg_thresh_lo)

    def pick(inds, size, choice):
        if inds.numel() == 0:
            return inds
        if choice is None:
            choice = torch.randperm(inds.numel(), device=dev)[:size]
        else:
            choice = torch.as_tensor(choice, device=dev, dtype=torch.long)
            if choice.numel() != size:
                raise ValueError("choice has %d entries, the minibatch takes %d" % (choice.numel(), size))
        return inds[choice.long()]

    n_fg = min(fg_rois_per_image, fg_inds.numel())
    fg_keep = pick(fg_inds, n_fg, fg_choice)
    n_bg = min(rois_per_image - n_fg, bg_inds.numel())
    bg_keep = pick(bg_inds, n_bg, bg_choice)
    keep = torch.cat([fg_keep, bg_keep]).to(torch.int32).contiguous()
    n = keep.numel()
    reg_classes = 2 if cls_agnostic_bbox_reg else int(num_classes)
    labels = torch.empty((n,), dtype=torch.int32, device=dev)
    rois = torch.empty((n, 5), dtype=torch.float32, device=dev)
    targets = torch.empty((n, 4 * reg_classes), dtype=torch.float32, device=dev)
    inside = torch.empty_like(targets)
    outside = torch.empty_like(targets)
    w = (ctypes.c_float * 4)(*[float(v) for v in bbox_reg_weights])
    G = int(gt_boxes.size(0)) if gt_boxes is not None else 0
    gt = _boxes4(gt_boxes, "gt_boxes") if G else None
    with torch.cuda.device(dev):
        _lib.check(_lib.load().b200_fast_rcnn_targets(boxes.data_ptr(), gt.data_ptr() if G else None, argmax.data_ptr(),
                                                      max_classes.data_ptr(), keep.data_ptr(), n, n_fg, w, reg_classes,
                                                      1 if cls_agnostic_bbox_reg else 0, float(im_scale), float(batch_idx),
                                                      labels.data_ptr(), rois.data_ptr(), targets.data_ptr(), inside.data_ptr(),
                                                      outside.data_ptr(), _stream()), "b200_fast_rcnn_targets")
    return dict(labels_int32=labels, rois=rois, bbox_targets=targets, bbox_inside_weights=inside, bbox_outside_weights=outside,
                keep_inds=keep)
