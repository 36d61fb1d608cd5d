"""Per-warp timeline of the quad-strip forward at cfg2 (timing build: -DB200_STRIP_TIMING, see roi_align_strip.cu).

    B200_NVCC_EXTRA=-DB200_STRIP_TIMING B200_ROI_OPS_LIB=$PWD/detectron/pytorch_b200/libvar_tim.so python -m detectron.pytorch_b200.build
    B200_ROI_OPS_LIB=$PWD/detectron/pytorch_b200/libvar_tim.so python tools/strip_timing.py [CW] [PW]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    from detectron.pytorch_b200 import _lib, ops, synthetic as S
    CW = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    PW = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    shape, s, P, sr, n = (1, 256, 200, 272), 0.25, 7, 2, 512
    feats = [torch.from_numpy(S.make_features(shape, seed=3 + i)).cuda() for i in range(3)]
    r = torch.from_numpy(S.make_rois(n, shape, s, seed=100).astype(np.float32)).cuda()
    _lib.set_option("B200_ROI_ALIGN_PATH", "quad")
    if len(sys.argv) > 3:
        _lib.set_option("B200_STREAM_STAGE", sys.argv[3])        # "async": cp.async producers instead of TMA + transposers
    W = CW + PW
    tim = torch.zeros((148 * W * 8,), dtype=torch.int64, device="cuda")
    _lib.load().b200_roi_ops_debug_timing_buffer(tim.data_ptr())
    for i in range(3):
        ops.roi_align_forward(feats[i], r, P, P, s, sr)
    torch.cuda.synchronize()
    tim.zero_()
    ops.roi_align_forward(feats[0], r, P, P, s, sr)
    torch.cuda.synchronize()
    t = tim.cpu().numpy().reshape(148, W, 8).astype(np.int64)
    start, first, end, wait, cnt, busy = (t[:, :, k] for k in range(6))
    t0 = start[start > 0].min()
    clk = 1.965  # GHz
    print("kernel span: %.1f us (first warp start -> last warp end)" % ((end.max() - t0) / 1e3))
    cta_end = (end.max(axis=1) - t0) / 1e3
    cta_start = (start.min(axis=1) - t0) / 1e3
    print("CTA start (us): min %.1f max %.1f | CTA end: min %.1f p10 %.1f median %.1f p90 %.1f max %.1f" % (
        cta_start.min(), cta_start.max(), cta_end.min(), np.percentile(cta_end, 10), np.median(cta_end), np.percentile(cta_end, 90), cta_end.max()))
    cons, prod = slice(0, CW), slice(CW, W)
    f_c = np.where(first[:, cons] > 0, first[:, cons] - start[:, cons], 0) / 1e3
    f_p = np.where(first[:, prod] > 0, first[:, prod] - start[:, prod], 0) / 1e3
    print("first fragment computed after (us, per consumer warp): median %.2f max %.2f | first row staged after: median %.2f max %.2f" % (
        np.median(f_c[f_c > 0]), f_c.max(), np.median(f_p[f_p > 0]), f_p.max()))
    dur_c = (end[:, cons] - start[:, cons]) / 1e3
    print("consumer warps: duration median %.1f us; waiting for rows %.1f us (%.0f %%); busy %.1f us; fragments / warp %.1f" % (
        np.median(dur_c), np.median(wait[:, cons]) / clk / 1e3, 100 * wait[:, cons].sum() / clk / 1e3 / dur_c.sum(),
        np.median(busy[:, cons]) / clk / 1e3, cnt[:, cons].mean()))
    dur_p = (end[:, prod] - start[:, prod]) / 1e3
    print("producer warps: duration median %.1f us; waiting for slots %.1f us; staging busy %.1f us; rows / warp %.1f; busy cycles / row %.0f" % (
        np.median(dur_p), np.median(wait[:, prod]) / clk / 1e3, np.median(busy[:, prod]) / clk / 1e3, cnt[:, prod].mean(),
        busy[:, prod].sum() / max(1, cnt[:, prod].sum())))
    for w in range(CW, W):
        print("  producer-side warp %d: duration %.1f us, wait %.1f us, busy %.1f us, rows %.1f, busy cycles / row %.0f" % (
            w - CW, np.median(dur_p[:, w - CW]), np.median(wait[:, w]) / clk / 1e3, np.median(busy[:, w]) / clk / 1e3, cnt[:, w].mean(),
            busy[:, w].sum() / max(1, cnt[:, w].sum())))
    worst = int(np.argmax(cta_end))
    print("slowest CTA %d: end %.1f us, rows %d, fragments %d, consumer wait %.1f us, producer wait %.1f us" % (
        worst, cta_end[worst], cnt[worst, prod].sum(), cnt[worst, cons].sum(), wait[worst, cons].mean() / clk / 1e3, wait[worst, prod].mean() / clk / 1e3))
    rows = cnt[:, prod].sum(axis=1); frags = cnt[:, cons].sum(axis=1)
    print("rows per CTA: min %d median %d max %d | fragments per CTA: min %d median %d max %d" % (rows.min(), np.median(rows), rows.max(), frags.min(), np.median(frags), frags.max()))
    c = np.corrcoef(np.stack([rows, frags, cta_end]))
    print("corr(end, rows) %.2f  corr(end, fragments) %.2f" % (c[2, 0], c[2, 1]))


if __name__ == "__main__":
    main()
