"""CPU-only: benchmark host logic -- bytes model vs the oracle's exact count, image sharding, and the
N>1 aggregation path (one process per rank, gloo, world_size 2)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from detectron.pytorch_b200 import benchutil
from detectron.pytorch_b200 import synthetic as S
from oracle import cpu as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_touched_cells_matches_oracle_and_survey():
    cfg = S.CFG2
    r = S.make_rois(cfg["rois"], cfg["shape"], cfg["scale"])
    n = benchutil.roi_align_touched_cells(r, 1, 200, 272, 7, 7, cfg["scale"], 2)
    assert n == O.roi_align_touched_cells(r, 1, 200, 272, 7, 7, cfg["scale"], 2) == 53927
    b = benchutil.roi_align_bytes(cfg["shape"], 512, 7, 7, touched_cells=n)
    assert b["fwd"] == 80921600 and b["bwd"] == 81405952
    c1 = S.CFG1
    r1 = S.make_rois(c1["rois"], c1["shape"], c1["scale"])
    for sr in (0, 2):
        assert benchutil.roi_align_touched_cells(r1, 1, 50, 68, 7, 7, c1["scale"], sr) == \
            O.roi_align_touched_cells(r1, 1, 50, 68, 7, 7, c1["scale"], sr)


def test_shard_images_balanced_and_complete():
    for n, w in ((16, 8), (16, 3), (5, 8), (1, 1)):
        shards = [benchutil.shard_images(n, r, w) for r in range(w)]
        assert sorted(sum(shards, [])) == list(range(n))
        assert max(map(len, shards)) - min(map(len, shards)) <= 1


def test_clock_summary_parsing():
    lines = ["0, 1965, 1965, 412.1, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
             "0, 1800, 1965, 612.1, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
             "0, 1900, 1965, 512.1, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active"]
    s = benchutil.summarize_clock_lines(lines)
    assert s["sm_mhz"] == 1900 and s["sm_max_mhz"] == 1965 and s["reasons"] == ["sw_power_cap"]


_WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from detectron.pytorch_b200 import benchutil
dist.init_process_group("gloo")
rank, world, local = benchutil.dist_env()
ms, units = benchutil.aggregate(10.0 * (rank + 1), 512 * 7, device=torch.device("cpu"))
imgs = benchutil.shard_images(16, rank, world)
if rank == 0:
    print(json.dumps({"ms": ms, "units": units, "world": world, "imgs": imgs}))
dist.destroy_process_group()
'''


def test_aggregate_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29641", str(script)],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["world"] == 2 and res["ms"] == 20.0 and res["units"] == 2 * 512 * 7    # max time, summed units
    assert res["imgs"] == list(range(8))


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "RoIs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True
