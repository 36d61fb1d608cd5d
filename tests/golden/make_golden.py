"""Generate tests/golden/*.npz from the REFERENCE's own CUDA kernels (oracle/_ref/*.so, built by
`make -C oracle ref` from the unmodified sources under /root/reference).  Needs a GPU:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'     # then copy into tests/golden/

Inputs are the seeded cases of tests/cases.py, so only the reference OUTPUTS are stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gpu_ref as G      # noqa: E402
from tests import cases              # noqa: E402


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    assert torch.cuda.is_available() and G.available()
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
    for name in sorted(cases.ROI_CASES):
        c, f, r, dy = cases.roi_case(name)
        P, s, sr = c["P"], c["scale"], c["sr"]
        F, Rr, DY = t(f), t(r), t(dy)
        out = G.roi_align_forward(F, Rr, P, P, s, sr)
        dx = G.roi_align_backward(DY, Rr, c["shape"], P, P, s, sr)
        np.savez_compressed(os.path.join(out_dir, "roi_align_xfrom_%s.npz" % name), out=out.cpu().numpy(), dx=dx.cpu().numpy())
        lo = G.roi_align_legacy_forward(F, Rr, P, P, s)
        ldx = G.roi_align_legacy_backward(DY, Rr, c["shape"], P, P, s)
        po, am = G.roi_pool_forward(F, Rr, P, P, s)
        pdx = G.roi_pool_backward(DY, am, Rr, c["shape"], P, P, s)
        np.savez_compressed(os.path.join(out_dir, "legacy_pool_%s.npz" % name), legacy_out=lo.cpu().numpy(),
                            legacy_dx=ldx.cpu().numpy(), pool_out=po.cpu().numpy(), pool_argmax=am.cpu().numpy(),
                            pool_dx=pdx.cpu().numpy())
    img, grid, go = cases.crop_case()
    I, Gd, GO = t(img), t(grid), t(go)
    out = G.roi_crop_forward(I, Gd)
    gi, gg = G.roi_crop_backward(I, Gd, GO)
    np.savez_compressed(os.path.join(out_dir, "roi_crop.npz"), out=out.cpu().numpy(), grad_img=gi.cpu().numpy(),
                        grad_grid=gg.cpu().numpy())
    keeps = {}
    for n in cases.NMS_SIZES:
        keeps["keep_%d" % n] = G.nms_gpu(t(cases.nms_case(n)), 0.7).cpu().numpy().reshape(-1).astype(np.int32)
    np.savez_compressed(os.path.join(out_dir, "nms.npz"), **keeps)
    torch.cuda.synchronize()
    print("golden written to", out_dir, sorted(os.listdir(out_dir)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden"))
