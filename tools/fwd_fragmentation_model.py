"""CPU model of the tiled forward's fragmentation: how many sample evaluations (one evaluation = one
bilinear sample for 4 bins x 32 channels) a launch needs at BASELINE cfg2 as a function of the tile core size and of
the bin-group size, against the ideal R * bins * sr^2 / 4 * C/32.  Each (RoI, tile) piece evaluates the rectangle of
bins that have a sample in the tile, rounded up to whole groups; bins split over tiles are evaluated once per tile.
The current kernel (core 17 x 31, groups of 8) is predicted at 363 648 evaluations; ncu counts 355 k (profiles/r02m).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from detectron.pytorch_b200 import synthetic as S

cfg = S.CFG2
N, C, H, W = cfg["shape"]; P, sr = cfg["pooled"], cfg["sampling_ratio"]; scale = np.float32(cfg["scale"])
rois = S.make_rois(cfg["rois"], cfg["shape"], cfg["scale"]).astype(np.float32)


def lows_of(start, binsz, size):
    p = np.repeat(np.arange(P), sr).astype(np.float32); i = np.tile(np.arange(sr), P).astype(np.float32)
    base = (p.astype(np.float64) * np.float64(binsz) + np.float64(start)).astype(np.float32)
    v = np.maximum((base + (((i + np.float32(.5)) * binsz) / np.float32(sr)).astype(np.float32)).astype(np.float32), 0)
    return np.minimum(v.astype(np.int64), size - 1)


lows = []
for r in rois:
    sw, sh = r[1] * scale, r[2] * scale
    rw = max(np.float32(r[3] * scale - sw), np.float32(1)); rh = max(np.float32(r[4] * scale - sh), np.float32(1))
    lows.append((lows_of(sh, np.float32(rh / np.float32(P)), H), lows_of(sw, np.float32(rw / np.float32(P)), W)))


def evaluations(core_h, core_w, group):
    tiles_y = -(-H // core_h); ch = -(-H // tiles_y)
    slots = 0
    for yl, xl in lows:
        ty, tx = yl // ch, xl // core_w
        for a in np.unique(ty):
            sy = np.nonzero(ty == a)[0]; nby = sy[-1] // sr - sy[0] // sr + 1
            for b in np.unique(tx):
                sx = np.nonzero(tx == b)[0]; nbx = sx[-1] // sr - sx[0] // sr + 1
                slots += -(-(nby * nbx) // group) * group
    return slots * sr * sr / 4 * (C // 32)


ideal = len(lows) * P * P * sr * sr / 4 * (C // 32)
print("ideal %.0f" % ideal)
for ch, cw, g in [(17, 31, 8), (17, 31, 4), (34, 31, 8), (50, 31, 8), (67, 31, 8), (200, 31, 8), (17, 63, 8), (34, 63, 8), (200, 272, 8)]:
    c = evaluations(ch, cw, g)
    print("core %3d x %3d, groups of %d: %7.0f evaluations  (%.2f x ideal)" % (ch, cw, g, c, c / ideal))
