"""Per-phase clock64 breakdown of the tiled RoIAlign forward at BASELINE cfg2 (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from detectron.pytorch_b200 import _lib, ops, synthetic as S
cfg = S.CFG2
f = torch.randn(cfg["shape"], device="cuda"); r = torch.from_numpy(S.make_rois(cfg["rois"], cfg["shape"], cfg["scale"])).cuda()
for _ in range(3):
    ops.roi_align_forward(f, r, 7, 7, cfg["scale"], 2)
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
lib = _lib.load(); lib.b200_roi_ops_debug_timing_buffer(buf.data_ptr())
n = 10
for _ in range(n):
    ops.roi_align_forward(f, r, 7, 7, cfg["scale"], 2)
torch.cuda.synchronize(); lib.b200_roi_ops_debug_timing_buffer(None)
t = buf.tolist(); w = max(t[3], 1)
print("per warp-item (cycles): staging %.0f  compute %.0f  end-wait %.0f   | warp-items/launch %d  RoI items/launch %d  8-bin groups/launch %d  max compute %d"
      % (t[0] / w, t[1] / w, t[2] / w, t[3] // n, t[4] // n, t[5] // n, t[6]))
print("cycles per 8-bin group (compute / groups): %.0f ; per RoI item: %.0f" % (t[1] / max(t[5], 1), t[1] / max(t[4], 1)))
