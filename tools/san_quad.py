"""compute-sanitizer case for the quad-strip forward: small shapes (partial channel group, narrow strips, edge RoIs)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from detectron.pytorch_b200 import _lib, ops, synthetic as S
_lib.set_option("B200_ROI_ALIGN_PATH", "quad")
for shape, s, P, sr, n in [((2, 8, 50, 68), 1.0 / 16, 7, 2, 32), ((2, 40, 60, 100), 1.0 / 8, 7, 2, 100), ((1, 32, 30, 9), 1.0 / 32, 7, 2, 12),
                            ((1, 64, 50, 84), 1.0 / 16, 14, 2, 40), ((3, 32, 40, 68), 1.0 / 16, 7, 1, 50)]:
    f = torch.from_numpy(S.make_features(shape, seed=1)).cuda()
    r = torch.from_numpy(np.concatenate([S.make_rois(n, shape, s, seed=2), S.make_edge_rois(shape, s)]).astype(np.float32)).cuda()
    before = _lib.launch_count()
    out = ops.roi_align_forward(f, r, P, P, s, sr)
    torch.cuda.synchronize()
    print(shape, "launches", _lib.launch_count() - before, float(out.abs().sum()), flush=True)
print("sanitizer case done")
