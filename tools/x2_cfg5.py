"""SURVEY.md row X2 / BASELINE.json configs[4]: one Mask R-CNN (e2e_mask_rcnn_R-50-FPN_1x) TRAINING step of the
UNMODIFIED reference model -- forward, losses, backward, SGD -- with this package's ops at the reference's import paths,
2 images per GPU, one process per GPU under torchrun; gradients are averaged with NCCL all-reduce (flat buckets) and
the loss dict with one small all-reduce (mean over ranks), replacing the reference's mynn.DataParallel
(tools/train_net_step.py:338,424-428; lib/nn/parallel/_functions.py:39,55; lib/utils/training_stats.py:82-95).

    python tools/x2_cfg5.py --steps 3                                  # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/x2_cfg5.py --steps 3

Synthetic data (no dataset on the box): random images written to a temp dir, 6 random ground-truth boxes with
rectangular polygon masks per image; the minibatch blobs are built by the reference's own roi_data code.
Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synthetic_roidb(n_images, seed, tmpdir, num_classes=81, hw=(600, 1000), n_gt=6):
    import cv2
    import numpy as np
    import scipy.sparse
    rng = np.random.RandomState(seed)
    roidb = []
    for i in range(n_images):
        H, W = hw
        im = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
        path = os.path.join(tmpdir, "im_%d_%d.png" % (seed, i))
        cv2.imwrite(path, im)
        cx = rng.uniform(0.2 * W, 0.8 * W, n_gt); cy = rng.uniform(0.2 * H, 0.8 * H, n_gt)
        bw = rng.uniform(40, 0.5 * W, n_gt); bh = rng.uniform(40, 0.5 * H, n_gt)
        x1 = np.clip(cx - bw / 2, 0, W - 2); x2 = np.clip(cx + bw / 2, x1 + 8, W - 1)
        y1 = np.clip(cy - bh / 2, 0, H - 2); y2 = np.clip(cy + bh / 2, y1 + 8, H - 1)
        boxes = np.stack([x1, y1, x2, y2], axis=1).astype(np.float32)
        classes = rng.randint(1, num_classes, n_gt).astype(np.int32)
        ov = np.zeros((n_gt, num_classes), dtype=np.float32)
        ov[np.arange(n_gt), classes] = 1.0
        segms = [[[float(a), float(b), float(c), float(b), float(c), float(d), float(a), float(d)]] for a, b, c, d in boxes]
        entry = {
            "id": seed * 1000 + i, "image": path, "flipped": False, "height": H, "width": W, "has_visible_keypoints": False,
            "boxes": boxes, "segms": segms, "seg_areas": ((x2 - x1 + 1) * (y2 - y1 + 1)).astype(np.float32),
            "gt_classes": classes, "gt_overlaps": scipy.sparse.csr_matrix(ov), "is_crowd": np.zeros((n_gt,), dtype=bool),
            "box_to_gt_ind_map": np.arange(n_gt, dtype=np.int32),
            "max_classes": classes.copy(), "max_overlaps": np.ones((n_gt,), dtype=np.float32),
        }
        roidb.append(entry)
    return roidb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--images", type=int, default=2, help="images per GPU")
    ap.add_argument("--config", default="e2e_mask_rcnn_R-50-FPN_1x.yaml")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--stub-ops", action="store_true", help="CPU dry run of the harness: torchvision RoIAlign instead of the product ops")
    args = ap.parse_args()

    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = args.device == "cuda"
    if use_cuda:
        torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl" if use_cuda else "gloo", **({"device_id": torch.device("cuda", local)} if use_cuda else {}))

    from oracle import refmodel
    if args.stub_ops:
        import types
        import torchvision.ops as tvo

        class StubFn(object):
            def __init__(self, h, w, s, sr): self.a = (h, w, s, sr)
            def __call__(self, f, r): return tvo.roi_align(f, r, (self.a[0], self.a[1]), self.a[2], self.a[3], aligned=False)
        for name in ("modeling.roi_xfrom", "modeling.roi_xfrom.roi_align", "modeling.roi_xfrom.roi_align.functions"):
            sys.modules.setdefault(name, types.ModuleType(name))
        m = types.ModuleType("modeling.roi_xfrom.roi_align.functions.roi_align"); m.RoIAlignFunction = StubFn
        sys.modules[m.__name__] = m
        for name, attr in (("model.roi_pooling.functions.roi_pool", "RoIPoolFunction"), ("model.roi_crop.functions.roi_crop", "RoICropFunction")):
            parts = name.split(".")
            for k in range(1, len(parts)):
                sys.modules.setdefault(".".join(parts[:k]), types.ModuleType(".".join(parts[:k])))
            mm = types.ModuleType(name); setattr(mm, attr, object); sys.modules[name] = mm
        refmodel.setup(use_b200_ops=False)
    else:
        refmodel.setup()
    model = refmodel.build_model(args.config, seed=0)
    from core.config import cfg
    cfg.TRAIN.IMS_PER_BATCH = args.images
    cfg.NUM_GPUS = world
    import roi_data.minibatch as minibatch
    import utils.blob as blob_utils
    import datasets.roidb as roidb_mod
    dev = torch.device("cuda", local) if use_cuda else torch.device("cpu")
    if not use_cuda:
        # the reference moves blobs with .cuda(device_id); on the CPU dry run those calls become no-ops
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.Tensor.get_device = lambda self: 0
    model.to(dev).train()
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=0.0025 * world, momentum=0.9, weight_decay=1e-4)

    tmpdir = tempfile.mkdtemp(prefix="x2_")
    np.random.seed(100 + rank)
    roidb = synthetic_roidb(args.images, seed=rank, tmpdir=tmpdir)
    roidb_mod.add_bbox_regression_targets(roidb)
    blobs, valid = minibatch.get_minibatch(roidb)
    assert valid
    inputs = {}
    for k, v in blobs.items():
        if k == "roidb":
            inputs[k] = [blob_utils.serialize([e]) for e in v]
        elif k == "im_info":
            inputs[k] = torch.from_numpy(np.asarray(v))
        else:
            inputs[k] = torch.from_numpy(np.asarray(v)).to(dev)

    def allreduce_grads():
        if world == 1:
            return 0.0
        t0 = time.perf_counter()
        bucket, size = [], 0
        def flush():
            nonlocal bucket, size
            if not bucket:
                return
            flat = torch.cat([g.reshape(-1) for g in bucket])
            dist.all_reduce(flat)
            flat.div_(world)
            off = 0
            for g in bucket:
                g.copy_(flat[off:off + g.numel()].view_as(g)); off += g.numel()
            bucket, size = [], 0
        for p in params:
            if p.grad is None:
                continue
            bucket.append(p.grad); size += p.grad.numel()
            if size >= (25 << 20) // 4:
                flush()
        flush()
        if use_cuda:
            torch.cuda.synchronize()
        return time.perf_counter() - t0

    def step():
        out = model(**inputs)
        losses = out["losses"]
        loss = sum(v.mean() for v in losses.values())
        opt.zero_grad()
        loss.backward()
        if use_cuda:
            torch.cuda.synchronize()
        t_ar = allreduce_grads()
        opt.step()
        keys = sorted(losses)
        vec = torch.stack([losses[k].detach().mean() for k in keys] + [loss.detach()])
        if world > 1:
            dist.all_reduce(vec); vec /= world                 # mean over ranks, like the reference's gather + mean(dim=0)
        return dict(zip(keys + ["total"], [float(x) for x in vec.cpu()])), t_ar

    for _ in range(args.warmup):
        step()
    if use_cuda:
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    t_ar = 0.0
    for _ in range(args.steps):
        ld, t = step(); t_ar += t
    if use_cuda:
        torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0, t_ar], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        n_params = sum(p.numel() for p in params)
        print(json.dumps({"row": "X2", "config": args.config, "n_gpus": world, "images_per_gpu": args.images, "steps": args.steps,
                          "step_ms": float(dt[0]) / args.steps * 1e3, "allreduce_ms_per_step": float(dt[1]) / args.steps * 1e3,
                          "grad_bytes": n_params * 4, "images_per_s": world * args.images * args.steps / float(dt[0]),
                          "losses": ld, "ops": "torchvision stub (CPU dry run)" if args.stub_ops else "detectron.pytorch_b200",
                          "reduction": "flat 25 MB buckets, NCCL all_reduce / world; loss dict: one all_reduce, mean over ranks"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
