"""Seeded synthetic inputs for the hot path (BASELINE.json configs 1-3).

Pure numpy so the same tensors can be produced on the CPU build container, on the GPU box, and
inside the CPU oracle's tests.  Nothing here touches CUDA.
"""
import math

import numpy as np

#: BASELINE.json configs: (N, C, H, W), spatial_scale, R, pooled size, sampling_ratio
CFG1 = dict(shape=(1, 256, 50, 68), scale=1.0 / 16, rois=32, pooled=7, sampling_ratio=2)
CFG2 = dict(shape=(1, 256, 200, 272), scale=1.0 / 4, rois=512, pooled=7, sampling_ratio=2)
CFG3 = dict(boxes=6000, thresh=0.7)


def make_features(shape, seed=0):
    """fp32 NCHW standard-normal feature map."""
    rng = np.random.RandomState(seed)
    return rng.standard_normal(size=shape).astype(np.float32)


def make_rois(num_rois, shape, scale, seed=0, min_size=32.0, max_size=512.0):
    """(R,5) fp32 [batch_idx, x1, y1, x2, y2] in image pixels.

    centre ~ U(image), w,h ~ exp(U(ln min, ln max)) px independently, clipped to the image.
    """
    N, _, H, W = shape
    rng = np.random.RandomState(seed)
    IH, IW = H / scale, W / scale
    cx = rng.uniform(0, IW, num_rois)
    cy = rng.uniform(0, IH, num_rois)
    w = np.exp(rng.uniform(math.log(min_size), math.log(max_size), num_rois))
    h = np.exp(rng.uniform(math.log(min_size), math.log(max_size), num_rois))
    x1 = np.clip(cx - w / 2, 0, IW - 1)
    x2 = np.clip(cx + w / 2, 0, IW - 1)
    y1 = np.clip(cy - h / 2, 0, IH - 1)
    y2 = np.clip(cy + h / 2, 0, IH - 1)
    b = rng.randint(0, N, num_rois)
    return np.stack([b, x1, y1, x2, y2], axis=1).astype(np.float32)


def make_edge_rois(shape, scale):
    """Hand-picked RoIs for the edge cases the parity tests must cover: boxes partly / fully
    outside the map, zero-area and inverted boxes, boxes hugging the H-1 / W-1 clamp."""
    N, _, H, W = shape
    IH, IW = H / scale, W / scale
    r = [
        [0, 0, 0, IW - 1, IH - 1],                 # whole image
        [0, -50, -40, 30, 25],                     # partly outside (top-left)
        [0, IW - 20, IH - 20, IW + 60, IH + 60],   # partly outside (bottom-right)
        [0, -300, -300, -200, -200],               # fully outside (negative)
        [0, IW + 100, IH + 100, IW + 200, IH + 200],  # fully outside (positive)
        [0, 40, 40, 40, 40],                       # zero-area
        [0, 90, 80, 60, 50],                       # inverted
        [0, IW - 1.5, IH - 1.5, IW - 1, IH - 1],   # hugging the last row/col
        [0, 0.25, 0.25, 0.75, 0.75],               # sub-cell box
        [N - 1, 10.5, 20.25, 200.75, 150.125],     # last image of the batch
        [0, 3, IH / 2, IW - 3, IH / 2 + 1],        # very wide, very flat
        [0, IW / 2, 2, IW / 2 + 1, IH - 2],        # very tall, very thin
    ]
    return np.asarray(r, dtype=np.float32)


def make_nms_boxes(n, seed=0, img_w=1333.0, img_h=800.0, copies=10):
    """(n,5) fp32 [x1,y1,x2,y2,score]: n/copies seed boxes x `copies` jittered copies, clipped to the
    image; scores = unique U(0,1) sorted descending (nms_gpu assumes pre-sorted input)."""
    rng = np.random.RandomState(seed)
    n_seed = max(1, int(math.ceil(n / float(copies))))
    cx = rng.uniform(0, img_w, n_seed)
    cy = rng.uniform(0, img_h, n_seed)
    w = np.exp(rng.uniform(math.log(16.0), math.log(400.0), n_seed))
    h = np.exp(rng.uniform(math.log(16.0), math.log(400.0), n_seed))
    cx = np.repeat(cx, copies)[:n]
    cy = np.repeat(cy, copies)[:n]
    w = np.repeat(w, copies)[:n]
    h = np.repeat(h, copies)[:n]
    cx = cx + rng.normal(0, 0.1, n) * w
    cy = cy + rng.normal(0, 0.1, n) * h
    w = w * (1 + rng.normal(0, 0.1, n))
    h = h * (1 + rng.normal(0, 0.1, n))
    w = np.maximum(w, 1.0)
    h = np.maximum(h, 1.0)
    x1 = np.clip(cx - w / 2, 0, img_w - 1)
    x2 = np.clip(cx + w / 2, 0, img_w - 1)
    y1 = np.clip(cy - h / 2, 0, img_h - 1)
    y2 = np.clip(cy + h / 2, 0, img_h - 1)
    perm = rng.permutation(n)
    scores = np.sort(rng.uniform(0, 1, n))[::-1]
    boxes = np.stack([x1, y1, x2, y2], axis=1)[perm]
    return np.concatenate([boxes, scores[:, None]], axis=1).astype(np.float32)


def make_crop_grid(num_rois, out_h, out_w, seed=0, spread=1.2):
    """(R, out_h, out_w, 2) fp32 sampling grid in the reference's (y, x) order, values mostly in
    [-1, 1] with some samples outside (zero-padding path)."""
    rng = np.random.RandomState(seed)
    cy = rng.uniform(-0.8, 0.8, (num_rois, 1, 1))
    cx = rng.uniform(-0.8, 0.8, (num_rois, 1, 1))
    sy = rng.uniform(0.05, spread, (num_rois, 1, 1))
    sx = rng.uniform(0.05, spread, (num_rois, 1, 1))
    ly = np.linspace(-1, 1, out_h).reshape(1, out_h, 1)
    lx = np.linspace(-1, 1, out_w).reshape(1, 1, out_w)
    gy = cy + sy * ly + 0 * lx
    gx = cx + sx * lx + 0 * ly
    return np.stack([gy, gx], axis=-1).astype(np.float32)
