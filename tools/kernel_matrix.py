"""Every kernel the library ships, timed at the shapes the model uses, next to the reference's own CUDA kernel
(oracle/_ref, recompiled for sm_100a) on the same inputs -- VERDICT r01 item 4.

    python tools/kernel_matrix.py [--iters 50] > profiles/<tag>_kernel_matrix.json

Rows: RoIAlign (Caffe2-exact) fwd / bwd at FPN P2..P5 (2 images, 800x1333 padded to /32, C = 256) for the box head
(7x7, sr 2, 1000 RoIs per level -- the worst case of one level taking them all) and the mask head (14x14, 256 RoIs);
BASELINE cfg1 / cfg2; RoIPool, RoICrop and legacy RoIAlign fwd / bwd at the C4 shape (1024 x 50 x 84 would be the C4 model;
here C = 256 on P4-sized maps to keep one table) .  Timing: CUDA events around `iters` back-to-back calls on rotating input sets
larger than L2 in total; bytes = algorithmic bytes of SURVEY 8d (bwd: dY + whole dX; RoIPool fwd adds the argmax; RoICrop
adds the grid).  Prints one JSON object: {"rows": [...], "peak_gbs": ...}.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=40)
    args = ap.parse_args()
    import numpy as np
    import torch
    from detectron.pytorch_b200 import benchutil, ops
    from detectron.pytorch_b200 import synthetic as S
    from detectron.pytorch_b200.model.roi_align.functions.roi_align import RoIAlignFunction as LegacyFn
    from detectron.pytorch_b200.model.roi_crop.functions.roi_crop import RoICropFunction
    from detectron.pytorch_b200.model.roi_pooling.functions.roi_pool import RoIPoolFunction
    from oracle import gpu_ref as G          # baseline beside ours, never on the product path
    have_ref = G.available()
    dev = torch.device("cuda", 0)
    peak = 6572.5
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])

    def timed(fn, n):
        fn(0); fn(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3          # us

    rows = []

    def add(name, shape, R, P, us, byts, ref_us, extra=None):
        row = {"kernel": name, "shape": list(shape), "rois": R, "pooled": P, "us": round(us, 2), "algorithmic_bytes": int(byts),
               "gbs": round(byts / us / 1e3, 1), "frac_of_measured_hbm": round(byts / us / 1e3 / peak, 4),
               "reference_kernel_us": None if ref_us is None else round(ref_us, 2)}
        if extra:
            row.update(extra)
        rows.append(row)
        sys.stderr.write("%-34s %-22s R=%-5d P=%-2d %8.1f us %7.0f GB/s (%.3f)  ref %s\n" % (
            name, "x".join(map(str, shape)), R, P, us, row["gbs"], row["frac_of_measured_hbm"], "-" if ref_us is None else "%.1f us" % ref_us))

    def sets_for(shape, R, P, scale, nsets, seed):
        g = torch.Generator(device=dev); g.manual_seed(seed)
        feats = [torch.randn(shape, generator=g, device=dev) for _ in range(nsets)]
        rois_np = [S.make_rois(R, shape, scale, seed=seed + i) for i in range(nsets)]
        rois = [torch.from_numpy(r).to(dev) for r in rois_np]
        dys = [torch.randn((R, shape[1], P, P), generator=g, device=dev) for _ in range(nsets)]
        return feats, rois_np, rois, dys

    def roi_align_rows(tag, shape, scale, R, P, sr, seed):
        nbytes = 4 * (shape[0] * shape[1] * shape[2] * shape[3] + R * shape[1] * P * P)
        nsets = max(2, min(8, int(300e6 // nbytes) + 1))
        feats, rois_np, rois, dys = sets_for(shape, R, P, scale, nsets, seed)
        touched = benchutil.roi_align_touched_cells(rois_np[0], shape[0], shape[2], shape[3], P, P, scale, sr)
        b = benchutil.roi_align_bytes(shape, R, P, P, touched_cells=touched)
        n = args.iters
        us_f = timed(lambda i: ops.roi_align_forward(feats[i % nsets], rois[i % nsets], P, P, scale, sr), n)
        us_b = timed(lambda i: ops.roi_align_backward(dys[i % nsets], rois[i % nsets], shape, P, P, scale, sr), n)
        rf = rb = None
        if have_ref:
            rf = timed(lambda i: G.roi_align_forward(feats[i % nsets], rois[i % nsets], P, P, scale, sr), max(5, n // 4))
            rb = timed(lambda i: G.roi_align_backward(dys[i % nsets], rois[i % nsets], shape, P, P, scale, sr), max(5, n // 4))
        add("roi_align_fwd " + tag, shape, R, P, us_f, b["fwd"], rf, {"touched_cells": touched, "sampling_ratio": sr})
        add("roi_align_bwd " + tag, shape, R, P, us_b, b["bwd"], rb, {"sampling_ratio": sr})

    roi_align_rows("cfg1", (1, 256, 50, 68), 1.0 / 16, 32, 7, 2, 1)
    roi_align_rows("cfg2", (1, 256, 200, 272), 1.0 / 4, 512, 7, 2, 2)
    for lvl, (h, w) in zip((2, 3, 4, 5), ((200, 336), (100, 168), (50, 84), (25, 42))):
        roi_align_rows("fpn_p%d box" % lvl, (2, 256, h, w), 1.0 / (2 ** lvl), 1000, 7, 2, 10 + lvl)
    roi_align_rows("fpn_p2 mask", (2, 256, 200, 336), 1.0 / 4, 256, 14, 2, 20)
    roi_align_rows("fpn_p3 mask", (2, 256, 100, 168), 1.0 / 8, 256, 14, 2, 21)

    # ---- RoIPool / legacy RoIAlign / RoICrop at a C4-like shape
    shape, scale, R, P = (2, 256, 50, 84), 1.0 / 16, 512, 7
    nsets = 6
    feats, rois_np, rois, dys = sets_for(shape, R, P, scale, nsets, 30)
    whole = 4 * shape[0] * shape[1] * shape[2] * shape[3]
    out_b = 4 * R * shape[1] * P * P
    n = args.iters

    def pool_f(i):
        fn = RoIPoolFunction(P, P, scale)
        return fn, fn(feats[i % nsets], rois[i % nsets])
    us = timed(lambda i: pool_f(i), n)
    ref = timed(lambda i: G.roi_pool_forward(feats[i % nsets], rois[i % nsets], P, P, scale), max(5, n // 4)) if have_ref else None
    add("roi_pool_fwd", shape, R, P, us, whole + 2 * out_b, ref)
    Fg = [f.clone().requires_grad_(True) for f in feats[:2]]
    outs = []
    for j in range(2):
        outs.append(RoIPoolFunction(P, P, scale)(Fg[j], rois[j]))

    def pool_b(i):
        j = i % 2
        Fg[j].grad = None
        outs[j].backward(dys[j], retain_graph=True)
    us = timed(pool_b, n)
    ref = None
    if have_ref:
        _, arg = G.roi_pool_forward(feats[0], rois[0], P, P, scale)
        ref = timed(lambda i: G.roi_pool_backward(dys[0], arg, rois[0], shape, P, P, scale), max(5, n // 4))
    add("roi_pool_bwd (autograd call)", shape, R, P, us, whole + 2 * out_b, ref)

    us = timed(lambda i: LegacyFn(P, P, scale)(feats[i % nsets], rois[i % nsets]), n)
    ref = timed(lambda i: G.roi_align_legacy_forward(feats[i % nsets], rois[i % nsets], P, P, scale), max(5, n // 4)) if have_ref else None
    add("roi_align_legacy_fwd", shape, R, P, us, whole + out_b, ref)
    Fl = [f.clone().requires_grad_(True) for f in feats[:2]]
    louts = [LegacyFn(P, P, scale)(Fl[j], rois[j]) for j in range(2)]

    def leg_b(i):
        j = i % 2
        Fl[j].grad = None
        louts[j].backward(dys[j], retain_graph=True)
    us = timed(leg_b, n)
    ref = timed(lambda i: G.roi_align_legacy_backward(dys[i % nsets], rois[i % nsets], shape, P, P, scale), max(5, n // 4)) if have_ref else None
    add("roi_align_legacy_bwd (autograd call)", shape, R, P, us, whole + out_b, ref)

    grids = [torch.from_numpy(S.make_crop_grid(R, P, P, seed=40 + i).astype(np.float32)).to(dev) for i in range(nsets)]
    Rn = R                                                       # R rois over N images: R / N per image
    us = timed(lambda i: RoICropFunction()(feats[i % nsets], grids[i % nsets]), n)
    ref = timed(lambda i: G.roi_crop_forward(feats[i % nsets], grids[i % nsets]), max(5, n // 4)) if have_ref else None
    add("roi_crop_fwd", shape, Rn, P, us, whole + out_b + 4 * R * P * P * 2, ref)
    Fc = [f.clone().requires_grad_(True) for f in feats[:2]]
    couts = [RoICropFunction()(Fc[j], grids[j]) for j in range(2)]

    def crop_b(i):
        j = i % 2
        Fc[j].grad = None
        couts[j].backward(dys[j], retain_graph=True)
    us = timed(crop_b, n)
    ref = timed(lambda i: G.roi_crop_backward(feats[i % nsets], grids[i % nsets], dys[i % nsets]), max(5, n // 4)) if have_ref else None
    add("roi_crop_bwd (autograd call)", shape, Rn, P, us, whole + out_b + 4 * R * P * P * 2, ref)

    # ---- NMS at the proposal sizes
    for nb in (1000, 2000, 6000, 12000):
        boxes = [torch.from_numpy(S.make_nms_boxes(nb, seed=i)).to(dev) for i in range(4)]
        us = timed(lambda i: ops.nms_raw(boxes[i % 4], 0.7), n)
        kept = int(ops.nms_raw(boxes[0], 0.7)[1].item())
        rows.append({"kernel": "nms", "boxes": nb, "us": round(us, 2), "kept": kept, "boxes_per_s": round(nb / us * 1e6)})
        sys.stderr.write("nms %6d boxes %8.1f us kept %d\n" % (nb, us, kept))
    print(json.dumps({"peak_gbs": peak, "iters": args.iters, "timing": "CUDA events, back-to-back direct launches through the Python op layer, rotating input sets",
                      "rows": rows}))


if __name__ == "__main__":
    main()
