timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "roi_crop" 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python - <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from detectron.pytorch_b200 import _lib, ops, synthetic as S
shape, R = (2, 256, 50, 84), 512
grid = torch.from_numpy(S.make_crop_grid(R, 7, 7, seed=1).astype(np.float32)).cuda()
go = torch.randn((R, 256, 7, 7), device="cuda")
for path in (None, "generic"):
    _lib.set_option("B200_ROI_ALIGN_BWD_PATH", path)
    for _ in range(5): ops.roi_crop_backward(go, grid, shape)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): ops.roi_crop_backward(go, grid, shape)
    e1.record(); torch.cuda.synchronize()
    print("roi_crop_bwd 2x256x50x84 R=512, path %s: %.1f us" % (path or "vector", e0.elapsed_time(e1) / 50 * 1e3))
PY
