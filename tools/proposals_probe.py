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
