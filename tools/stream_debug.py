"""First-contact check of the streaming-strip forward on a GPU box: a few shapes against the oracle, each in its own
subprocess under a timeout so that a protocol bug (deadlock -> watchdog trap, illegal address) cannot take the session down.

    python tools/stream_debug.py            # all cases
    python tools/stream_debug.py one NAME   # one case in this process
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    "tiny": ((1, 32, 16, 32), 1.0 / 16, 7, 2, 8, 32, 256),
    "small": ((1, 32, 50, 68), 1.0 / 16, 7, 2, 32, 32, 512),
    "c40_n2": ((2, 40, 60, 100), 1.0 / 8, 7, 2, 200, 32, 512),
    "wide": ((1, 32, 50, 336), 1.0 / 4, 7, 2, 150, 16, 1300),
    "tall": ((1, 32, 400, 64), 1.0 / 4, 7, 2, 100, 16, 1590),
    "odd_w": ((2, 48, 25, 42), 1.0 / 32, 7, 2, 60, 64, 900),
    "cfg2": ((1, 256, 200, 272), 1.0 / 4, 7, 2, 512, 32, 512),
}


def one(name):
    import numpy as np
    import torch
    from detectron.pytorch_b200 import _lib, synthetic as S
    from detectron.pytorch_b200.modeling.roi_xfrom.roi_align.functions.roi_align import RoIAlignFunction
    from oracle import cpu as O
    shape, s, P, sr, n, lo, hi = CASES[name]
    f = S.make_features(shape, seed=3)
    r = S.make_rois(n, shape, s, seed=4, min_size=lo, max_size=hi).astype(np.float32)
    ref = O.roi_align_forward(f, r, P, P, s, sr)
    _lib.set_option("B200_ROI_ALIGN_PATH", "stream")
    dbg = torch.zeros((148 * 17 * 8,), dtype=torch.int64).pin_memory()      # survives a trap: host memory
    _lib.load().b200_roi_ops_debug_timing_buffer(dbg.data_ptr())
    before = _lib.launch_count()
    try:
        out = RoIAlignFunction(P, P, s, sr)(torch.from_numpy(f).cuda(), torch.from_numpy(r).cuda())
        torch.cuda.synchronize()
    except Exception as exc:  # noqa: BLE001
        print("%s: FAILED %s" % (name, str(exc).splitlines()[0]))
        d = dbg.numpy().reshape(148, 17, 8).astype(np.uint64)
        shown = 0
        for blk in range(148):
            if not d[blk].any():
                continue
            if shown >= 6:
                break
            shown += 1
            for w in (0, 1, 15, 16):
                v = d[blk, w]
                print("   blk %3d warp %2d stage=%d L=(%d,%d) item(ya,yb,yhi,ne)=(%d,%d,%d,%d) w3=%016x w4=%016x dead=%016x" % (
                    blk, w, v[0], v[1] >> 32, v[1] & 0xffffffff, v[2] >> 48, (v[2] >> 32) & 0xffff, (v[2] >> 16) & 0xffff, v[2] & 0xffff,
                    v[3], v[4], v[6]))
        stages = d[:, :, 0]
        print("   stage histogram:", {int(k): int((stages == k).sum()) for k in np.unique(stages)}, " dead warps:", int((d[:, :, 6] != 0).sum()))
        return 3
    launches = _lib.launch_count() - before
    out = out.cpu().numpy()
    diff = np.abs(out - ref)
    bad = np.argwhere(diff > 1e-6)
    print("%s: launches=%d max|diff|=%.3g exact=%.5f bad=%d nan=%d" % (name, launches, diff.max(), (out == ref).mean(), len(bad),
                                                                        int(np.isnan(out).sum())))
    for b in bad[:8]:
        print("   bad at (r,c,ph,pw)=%s got %.6f want %.6f" % (tuple(b), out[tuple(b)], ref[tuple(b)]))
    if len(bad):
        rois_bad = np.unique(bad[:, 0]); ch_bad = np.unique(bad[:, 1])
        print("   bad rois:", rois_bad[:20], "n=", len(rois_bad), " bad channels:", ch_bad[:40], "n=", len(ch_bad))
    return 0 if len(bad) == 0 and launches == 3 else 1


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "one":
        sys.exit(one(sys.argv[2]))
    rc = 0
    for name in CASES:
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "one", name], timeout=180, stdout=subprocess.PIPE,
                               stderr=subprocess.STDOUT, text=True)
            print(p.stdout.strip()[-1500:]); print("   rc=%d" % p.returncode)
            rc |= p.returncode
        except subprocess.TimeoutExpired:
            print("%s: TIMEOUT" % name); rc |= 2
    sys.exit(rc)
