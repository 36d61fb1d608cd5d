"""Phase costs of the quad-strip forward's prepass (run under ncu --metrics gpu__time_duration.sum): the kernel returns after
phase 1 (B200_STREAM_PHASES=1), after the grid barrier + scan (=2), or runs whole (=prepass)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from detectron.pytorch_b200 import _lib, ops, synthetic as S
shape, s, P, sr, n = (1, 256, 200, 272), 0.25, 7, 2, 512
f = torch.from_numpy(S.make_features(shape, seed=3)).cuda()
r = torch.from_numpy(S.make_rois(n, shape, s, seed=100).astype(np.float32)).cuda()
_lib.set_option("B200_ROI_ALIGN_PATH", "quad")
for ph in ("1", "2", "prepass"):
    _lib.set_option("B200_STREAM_PHASES", ph)
    for _ in range(6):
        ops.roi_align_forward(f, r, P, P, s, sr)
    torch.cuda.synchronize()
