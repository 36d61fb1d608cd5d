"""SURVEY.md row X1 / BASELINE.json configs[3]: the UNMODIFIED reference model (Generalized_RCNN built from
configs/baselines/e2e_faster_rcnn_R-50-FPN_1x.yaml, random init) running its forward on a B200 with this package's ops
aliased in at the reference's own import paths -- and, as the checker, the same model with the reference's own CUDA
kernels (oracle/_ref, compiled unmodified for sm_100a) behind `RoIAlignFunction`.

    python tools/x1_cfg4.py [--images 2] [--iters 5]        # prints one JSON line

What is compared: every call of `Generalized_RCNN.roi_feature_transform` (model_builder.py:252-322) made by the heads
during the forward is re-run on the SAME inputs with the reference kernels; outputs must agree bit for bit (<= 1e-6 on
bins the streaming path splits).  What is timed: wall clock of `model(data, im_info)` (eval) with either op, after
warm-up, including the reference's host-side proposal code.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class RefRoIAlignFunction(object):
    """construct-then-call, like the reference's legacy Function (functions/roi_align.py:7-48), forward only"""

    def __init__(self, aligned_height, aligned_width, spatial_scale, sampling_ratio):
        self.args = (int(aligned_height), int(aligned_width), float(spatial_scale), int(sampling_ratio))

    def __call__(self, features, rois):
        from oracle import gpu_ref as G
        ah, aw, sc, sr = self.args
        return G.roi_align_forward(features.contiguous(), rois.contiguous(), ah, aw, sc, sr)


def run(cfg_name="e2e_faster_rcnn_R-50-FPN_1x.yaml", images=2, iters=5, hw=(800, 1344), seed=0):
    import numpy as np
    import torch
    from oracle import gpu_ref as G
    from oracle import refmodel
    import detectron.pytorch_b200 as pkg
    refmodel.setup()
    pkg.install_reference_aliases(nms=True)              # ops + utils.boxes.nms
    model = refmodel.build_model(cfg_name, seed=seed).cuda().eval()
    import modeling.model_builder as mb
    from detectron.pytorch_b200.modeling.roi_xfrom.roi_align.functions.roi_align import RoIAlignFunction as OurFn
    assert mb.RoIAlignFunction is OurFn, "the reference's import path must resolve to this package's op"

    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    data = torch.randn((images, 3) + tuple(hw), device="cuda", generator=g)
    im_info = torch.tensor([[800.0, 1333.0, 1.6]] * images)

    calls = []
    heads = [m for m in (getattr(model, "Box_Head", None), getattr(model, "Mask_Head", None), getattr(model, "Keypoint_Head", None))
             if m is not None and hasattr(m, "roi_xform")]
    orig = heads[0].roi_xform

    def spy(blobs_in, rpn_ret, **kw):
        out = orig(blobs_in, rpn_ret, **kw)
        calls.append((blobs_in, rpn_ret, kw, out))
        return out

    for h in heads:
        h.roi_xform = spy
    with torch.no_grad():
        ret = model(data, im_info)
    torch.cuda.synchronize()
    for h in heads:
        h.roi_xform = orig
    n_rois = int(ret["rois"].shape[0])

    # ---- parity of every roi_feature_transform call against the reference kernels on the same inputs
    parity = []
    mb.RoIAlignFunction = RefRoIAlignFunction
    try:
        with torch.no_grad():
            for blobs_in, rpn_ret, kw, out in calls:
                ref = orig(blobs_in, rpn_ret, **kw)
                parity.append({"shape": list(out.shape), "max_abs_diff": float((out - ref).abs().max()),
                               "frac_bit_equal": float((out == ref).float().mean()),
                               "levels": [int(len(rpn_ret[k])) for k in sorted(rpn_ret) if k.startswith(kw.get("blob_rois", "rois") + "_fpn")]})
    finally:
        mb.RoIAlignFunction = OurFn

    def timed(fn_cls):
        mb.RoIAlignFunction = fn_cls
        try:
            with torch.no_grad():
                model(data, im_info)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(iters):
                    model(data, im_info)
                torch.cuda.synchronize()
                return (time.perf_counter() - t0) / iters * 1e3
        finally:
            mb.RoIAlignFunction = OurFn

    ms_ours = timed(OurFn)
    ms_ref = timed(RefRoIAlignFunction) if G.available() else None
    return {"config": cfg_name, "images": images, "input": [images, 3] + list(hw), "rois": n_rois,
            "roi_feature_transform_calls": len(calls), "parity_vs_reference_kernels": parity,
            "forward_ms_b200_ops": ms_ours, "forward_ms_reference_kernels_sm100a": ms_ref,
            "note": "wall clock of the whole reference forward (backbone, FPN, RPN, host proposal code, heads); random-init weights"}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=2)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--config", default="e2e_faster_rcnn_R-50-FPN_1x.yaml")
    a = ap.parse_args()
    print(json.dumps(run(a.config, a.images, a.iters)))
