"""Shared, seeded test cases (inputs only -- expected values come from the oracle / golden files)."""
import numpy as np

from detectron.pytorch_b200 import synthetic as S

# name -> dict(shape, scale, P, sr, n_rois)   small enough for the CPU oracle to finish in < 1 s
ROI_CASES = {
    "cfg1_small": dict(shape=(2, 8, 50, 68), scale=1.0 / 16, P=7, sr=2, n_rois=32),
    "adaptive": dict(shape=(2, 8, 50, 68), scale=1.0 / 16, P=7, sr=0, n_rois=32),
    "p14": dict(shape=(1, 6, 50, 68), scale=1.0 / 16, P=14, sr=2, n_rois=24),
    "odd_hw": dict(shape=(2, 5, 25, 42), scale=1.0 / 32, P=7, sr=2, n_rois=24),
    "sr3": dict(shape=(1, 4, 40, 40), scale=1.0 / 8, P=5, sr=3, n_rois=16),
}

NMS_SIZES = (1, 2, 63, 64, 65, 128, 129, 1000, 2000, 6000, 12000)


def roi_case(name):
    c = ROI_CASES[name]
    feats = S.make_features(c["shape"], seed=0)
    rois = S.make_rois(c["n_rois"], c["shape"], c["scale"], seed=0)
    rois = np.concatenate([rois, S.make_edge_rois(c["shape"], c["scale"])]).astype(np.float32)
    R = rois.shape[0]
    dy = np.random.RandomState(1).standard_normal((R, c["shape"][1], c["P"], c["P"])).astype(np.float32)
    return c, feats, rois, dy


def crop_case(seed=0):
    shape = (2, 6, 30, 44)
    img = S.make_features(shape, seed=seed)
    grid = S.make_crop_grid(8, 7, 7, seed=seed)      # R = 8 -> 4 RoIs per image
    grid[0, 0, 0] = (-1.0, -1.0)                      # exact corners / borders
    grid[0, 0, 1] = (1.0, 1.0)
    grid[1, 3, 3] = (-1.5, 0.2)                       # outside
    grid[2, 2, 2] = (0.3, 1.7)
    go = np.random.RandomState(seed + 1).standard_normal((8, shape[1], 7, 7)).astype(np.float32)
    return img, grid.astype(np.float32), go


def nms_case(n, seed=0):
    return S.make_nms_boxes(n, seed=seed)


def nms_chain_case(n, shift=10.0, width=100.0):
    """Worst case of the block-parallel resolve: box i overlaps box i+1 above 0.7 but not box i+2, so the greedy
    result alternates keep / drop and every decision depends on the previous one (suppression chain of length n)."""
    import numpy as np
    x1 = np.arange(n, dtype=np.float32) * np.float32(shift)
    b = np.stack([x1, np.zeros(n, np.float32), x1 + np.float32(width), np.full(n, 50, np.float32),
                  np.linspace(1.0, 0.1, n).astype(np.float32)], axis=1)
    return b.astype(np.float32)


def nms_clustered_case(n, seed=0, copies=10):
    """Score-sorted proposals where near-duplicates are ADJACENT in the order (dense diagonal blocks)."""
    import numpy as np
    b = S.make_nms_boxes(n, seed=seed, copies=copies)
    rng = np.random.RandomState(seed + 1)
    seeds = b[rng.permutation(n)[: (n + copies - 1) // copies], :4]
    rep = np.repeat(seeds, copies, axis=0)[:n] + rng.normal(0, 2.0, (n, 4)).astype(np.float32)
    rep[:, 2:] = np.maximum(rep[:, 2:], rep[:, :2] + 1)
    return np.concatenate([rep, b[:, 4:5]], axis=1).astype(np.float32)
