"""time the device proposal layer against the numpy restatement of the reference's host path
(oracle/proposals.py; the reference itself also pays three blocking D2H copies before this work starts)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from detectron.pytorch_b200.modeling.generate_proposals import GenerateProposalsOp
from oracle import proposals as OP

rng = np.random.RandomState(0)
for name, (N, A, H, W, stride, pre, post) in {"FPN P2 train (2 img, 3x200x336, top 2000 -> 1000)": (2, 3, 200, 336, 4, 2000, 1000),
                                             "C4 test (1 img, 15x50x84, top 6000 -> 1000)": (1, 15, 50, 84, 16, 6000, 1000)}.items():
    anchors = rng.uniform(-1, 1, (A, 4)) * 40 + np.array([-30, -30, 30, 30])
    anchors = np.round(anchors * 2) / 2
    scores = ((rng.permutation(N * A * H * W).astype(np.float32) + 0.5) / (N * A * H * W)).reshape(N, A, H, W)
    deltas = (rng.standard_normal((N, 4 * A, H, W)) * 0.5).astype(np.float32)
    im_info = np.array([[H * stride, W * stride, 1.5]] * N, dtype=np.float32)
    mode = dict(RPN_PRE_NMS_TOP_N=pre, RPN_POST_NMS_TOP_N=post, RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=0)
    op = GenerateProposalsOp(anchors, 1.0 / stride, train=mode, test=mode)
    d_s, d_d, t_i = torch.from_numpy(scores).cuda(), torch.from_numpy(deltas).cuda(), torch.from_numpy(im_info)
    for _ in range(3):
        rois, _p = op(d_s, d_d, t_i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        rois, _p = op(d_s, d_d, t_i)
    torch.cuda.synchronize(); ms_dev = (time.perf_counter() - t0) / 20 * 1e3
    t0 = time.perf_counter()
    for _ in range(3):
        r2, _p2 = OP.generate_proposals(scores, deltas, im_info, anchors, float(stride), pre, post, 0.7, 0, nms="cython")
    ms_cpu = (time.perf_counter() - t0) / 3 * 1e3
    print("%-52s device %.2f ms (incl. the final D2H, wall)   host restatement %.1f ms   rois %d / %d" % (name, ms_dev, ms_cpu, len(rois), len(r2)))

# ---- all (level, image) problems of one FPN step: 5 levels x 2 images = 10 NMS problems in one batched launch pair,
#      one host read; against the same ten problems through five separate per-level calls
from detectron.pytorch_b200.modeling.generate_proposals import generate_proposals_batched
levels = [(2, 200, 336), (3, 100, 168), (4, 50, 84), (5, 25, 42), (6, 13, 21)]
N, A, pre, post = 2, 3, 2000, 1000
ops_l, probs_l, preds_l = [], [], []
for lvl, H, W in levels:
    stride = 2 ** lvl
    anchors = np.round((rng.uniform(-1, 1, (A, 4)) * 4 * stride + np.array([-3, -3, 3, 3]) * stride) * 2) / 2
    scores = ((rng.permutation(N * A * H * W).astype(np.float32) + 0.5) / (N * A * H * W)).reshape(N, A, H, W)
    deltas = (rng.standard_normal((N, 4 * A, H, W)) * 0.5).astype(np.float32)
    mode = dict(RPN_PRE_NMS_TOP_N=pre, RPN_POST_NMS_TOP_N=post, RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=0)
    ops_l.append(GenerateProposalsOp(anchors, 1.0 / stride, train=mode, test=mode, return_tensors=True))
    probs_l.append(torch.from_numpy(scores).cuda()); preds_l.append(torch.from_numpy(deltas).cuda())
t_i = torch.from_numpy(np.array([[800, 1333, 1.6]] * N, dtype=np.float32))


def fused():
    return generate_proposals_batched(ops_l, probs_l, preds_l, t_i)


def per_level():
    return [op(p, d, t_i) for op, p, d in zip(ops_l, probs_l, preds_l)]


for fn, label in ((fused, "one batched NMS over the 10 (level, image) problems, one host read"),
                  (per_level, "five per-level calls (each: one batched NMS over its 2 images)")):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        out = fn()
    torch.cuda.synchronize()
    print("FPN step, P2..P6 x 2 images, top 2000 -> NMS 0.7 -> 1000:  %.2f ms  (%s)  rois per level %s" % (
        (time.perf_counter() - t0) / 20 * 1e3, label, [int(o[0].shape[0]) for o in out]))
# share of the library top-k (torch.topk) in the fused call
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    for p in probs_l:
        for i in range(N):
            flat = p[i].permute(1, 2, 0).reshape(-1)
            if pre < flat.numel():
                torch.topk(flat, pre, largest=True, sorted=True)
            else:
                torch.sort(flat, descending=True, stable=True)
torch.cuda.synchronize()
print("   of which torch.topk / torch.sort on the 10 score maps: %.2f ms" % ((time.perf_counter() - t0) / 20 * 1e3))
