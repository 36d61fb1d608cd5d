"""What bounds bench.py's e2e step on this box?  Times the pipeline with its legs switched off one by one."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from detectron.pytorch_b200 import benchutil, synthetic as S
from detectron.pytorch_b200.modeling.roi_xfrom.roi_align.functions.roi_align import RoIAlignFunction

cfg = S.CFG2
device = torch.device("cuda:0")
shape = cfg["shape"]; R, C, P = cfg["rois"], cfg["shape"][1], cfg["pooled"]
rois = S.make_rois(R, shape, cfg["scale"], seed=1)
fn = RoIAlignFunction(P, P, cfg["scale"], cfg["sampling_ratio"])
for placement in ("default", "numa_local", "default", "numa_local"):
    ctx = benchutil.numa_local(0) if placement == "numa_local" else benchutil.numa_local.__new__(benchutil.numa_local)
    if placement == "default":
        ctx.cpus = None; ctx.prev = None
    with ctx:
        h_feat = torch.randn(shape).pin_memory(); h_rois = torch.from_numpy(rois).pin_memory()
        h_dy = torch.randn((R, C, P, P)).pin_memory()
        h_out = torch.empty((R, C, P, P)).pin_memory(); h_dx = torch.empty(shape).pin_memory()
    print(placement, benchutil.pcie_bandwidth(device, h_feat, h_dx))
    for legs in ((1, 0, 0), (0, 0, 1), (1, 0, 1), (0, 1, 0), (1, 1, 0), (1, 1, 1)):
        pipe = benchutil.E2EPipeline(fn, shape, R, C, P, device, h_feat, h_rois, h_dy, h_out, h_dx,
                                     h2d=bool(legs[0]), compute=bool(legs[1]), d2h=bool(legs[2]))
        pipe.run(3); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pipe.run(20); e1.record(); torch.cuda.synchronize()
        print("  legs h2d=%d compute=%d d2h=%d: %.3f ms/step  (allocator reserved %.0f MB)" % (legs + (e0.elapsed_time(e1) / 20, torch.cuda.memory_reserved() / 1e6)))
