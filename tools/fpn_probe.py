"""time the reference flow of roi_feature_transform (per-level RoIAlign -> torch.cat -> restore gather)
against RoIAlignFPNFunction (indexed writes, no intermediate) on an FPN-shaped workload: 2 images 800x1333 (padded),
levels P2..P5 = 200x336 .. 25x42, C = 256, 1024 RoIs mapped to levels as lib/utils/fpn.py:11-28 does, 7x7, sr = 2."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from detectron.pytorch_b200 import synthetic as S
from detectron.pytorch_b200.modeling.roi_xfrom.roi_align.functions.roi_align import RoIAlignFunction
from detectron.pytorch_b200.modeling.roi_xfrom.roi_align.functions.roi_align_fpn import RoIAlignFPNFunction

dev = torch.device("cuda:0")
N, C, P, sr = 2, 256, 7, 2
shapes = [(N, C, 200, 336), (N, C, 100, 168), (N, C, 50, 84), (N, C, 25, 42)]
scales = [1 / 4., 1 / 8., 1 / 16., 1 / 32.]
R = 1024
rois = S.make_rois(R, shapes[0], scales[0], seed=0).astype(np.float32)
area = (rois[:, 3] - rois[:, 1] + 1) * (rois[:, 4] - rois[:, 2] + 1)
lvl = np.clip(np.floor(4 + np.log2(np.sqrt(area) / 224.0 + 1e-6)), 2, 5).astype(int)      # k0 = 4, s0 = 224
per_level = [rois[lvl == k] for k in (2, 3, 4, 5)]
order = np.concatenate([np.nonzero(lvl == k)[0] for k in (2, 3, 4, 5)])                       # shuffled -> original
restore = np.argsort(order).astype(np.int32)                                                   # original -> shuffled row
print("RoIs per level:", [len(r) for r in per_level])
feats = [torch.randn(sh, device=dev) for sh in shapes]
d_rois = [torch.from_numpy(r).to(dev) for r in per_level]
dy = torch.randn((R, C, P, P), device=dev)
restore_t = torch.from_numpy(restore.astype(np.int64)).to(dev)


def composed(train):
    F = [f.detach().requires_grad_(train) for f in feats]
    outs = [RoIAlignFunction(P, P, sc, sr)(f, r) for f, r, sc in zip(F, d_rois, scales) if r.size(0)]
    out = torch.cat(outs, dim=0)[restore_t]
    if train:
        out.backward(dy)
    return out


def fused(train):
    F = [f.detach().requires_grad_(train) for f in feats]
    out = RoIAlignFPNFunction(P, P, scales, sr)(F, d_rois, restore)
    if train:
        out.backward(dy)
    return out


a, b = composed(False), fused(False)
torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-6)
res = {}
for name, fn in (("per-level loop + cat + gather", composed), ("RoIAlignFPNFunction", fused)):
    for train in (False, True):
        for _ in range(5):
            fn(train)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        e0.record()
        for _ in range(n):
            fn(train)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        res["%s, %s" % (name, "fwd+bwd" if train else "fwd")] = ms
        print("%-34s %-8s %.3f ms  (%.2f M RoIs/s)" % (name, "fwd+bwd" if train else "fwd", ms, R / ms / 1e3))
print(json.dumps(res))
