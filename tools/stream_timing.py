"""Per-warp timeline of the streaming forward at cfg2 (timing build: -DB200_STREAM_TIMING, see roi_align_stream.cu).

    B200_NVCC_EXTRA=-DB200_STREAM_TIMING B200_ROI_OPS_LIB=$PWD/detectron/pytorch_b200/libb200_roi_ops_tim.so python -m detectron.pytorch_b200.build
    B200_ROI_OPS_LIB=$PWD/detectron/pytorch_b200/libb200_roi_ops_tim.so python tools/stream_timing.py [out.json]

Prints, per CTA class, where the cycles of the main kernel go: start-up until the first fragment is resident, time spent
waiting on row barriers, time in the bin loop + flush, and the spread of CTA durations (the slowest CTA is the kernel time).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    from detectron.pytorch_b200 import _lib, synthetic as S
    from detectron.pytorch_b200.modeling.roi_xfrom.roi_align.functions.roi_align import RoIAlignFunction
    shape, s, P, sr, n = (1, 256, 200, 272), 0.25, 7, 2, 512
    f = torch.from_numpy(S.make_features(shape, seed=3)).cuda()
    r = torch.from_numpy(S.make_rois(n, shape, s, seed=100).astype(np.float32)).cuda()
    _lib.set_option("B200_ROI_ALIGN_PATH", "stream")
    CW, PW = 16, 4
    W = CW + PW
    tim = torch.zeros((148 * W * 8,), dtype=torch.int64, device="cuda")
    _lib.load().b200_roi_ops_debug_timing_buffer(tim.data_ptr())
    op = RoIAlignFunction(P, P, s, sr)
    for _ in range(3):
        op(f, r)
    torch.cuda.synchronize()
    tim.zero_()
    op(f, r)
    torch.cuda.synchronize()
    t = tim.cpu().numpy().reshape(148, W, 8).astype(np.int64)
    start, first, end, wait, nfr, comp = (t[:, :, k] for k in range(6))
    L = t[:, 0, 7]
    rows = (L & 0xffffffff) - (L >> 32)
    cta_start = start.min(axis=1)
    cta_end = end.max(axis=1)
    dur = cta_end - cta_start
    cons, prod = slice(0, CW), slice(CW, W)
    out = {
        "cta_cycles": {"min": int(dur.min()), "mean": float(dur.mean()), "max": int(dur.max())},
        "rows_per_piece": {"min": int(rows.min()), "mean": float(rows.mean()), "max": int(rows.max())},
        "consumer": {
            "startup_to_first_fragment": float((first[:, cons] - start[:, cons]).clip(min=0).mean()),
            "wait_cycles": float(wait[:, cons].mean()),
            "compute_cycles": float(comp[:, cons].mean()),
            "lifetime": float((end[:, cons] - start[:, cons]).mean()),
            "fragments_per_warp": float(nfr[:, cons].mean()),
            "finish_spread_within_cta": float((end[:, cons].max(axis=1) - end[:, cons].min(axis=1)).mean()),
        },
        "producer": {
            "startup_to_first_row": float((first[:, prod] - start[:, prod]).clip(min=0).mean()),
            "wait_cycles": float(wait[:, prod].mean()),
            "lifetime": float((end[:, prod] - start[:, prod]).mean()),
            "rows_per_warp": float(nfr[:, prod].mean()),
        },
    }
    order = np.argsort(dur)
    out["slowest_ctas"] = [{"cta": int(c), "cycles": int(dur[c]), "rows": int(rows[c]), "fragments": int(nfr[c, cons].sum()),
                            "cons_wait": float(wait[c, cons].mean()), "cons_comp": float(comp[c, cons].mean())} for c in order[-5:]]
    out["fastest_ctas"] = [{"cta": int(c), "cycles": int(dur[c]), "rows": int(rows[c]), "fragments": int(nfr[c, cons].sum()),
                            "cons_wait": float(wait[c, cons].mean()), "cons_comp": float(comp[c, cons].mean())} for c in order[:5]]
    # least squares: CTA cycles ~ a * rows + b * fragments + c
    A = np.stack([rows, nfr[:, cons].sum(axis=1), np.ones_like(rows)], axis=1).astype(np.float64)
    coef, *_ = np.linalg.lstsq(A, dur.astype(np.float64), rcond=None)
    out["fit_cycles"] = {"per_row": float(coef[0]), "per_fragment": float(coef[1]), "const": float(coef[2])}
    print(json.dumps(out, indent=1))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
