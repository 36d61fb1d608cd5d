#!/usr/bin/env bash
# full parity suite + bench + X2 on one GPU + kernel matrix (+ optional ncu)
set -u
TAG=${1:-r03f}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
echo "== pytest -m gpu (full)"; timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee "$OUT/status.txt"; tail -12 "$OUT/pytest_gpu.log"
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 600 python bench.py --steps 200 --warmup 10 > "$OUT/bench.json" 2> "$OUT/bench.err"; python - "$OUT/bench.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); k=d["kernels"]
print("value %.0f RoIs/s fwd %.4f (%.3f) bwd %.4f (%.3f) nms %.4f e2e %.0f roofline %s %.3f" % (d["value"], k["fwd"]["ms"], k["fwd"]["frac_of_measured"], k["bwd"]["ms"], k["bwd"]["frac_of_measured"], k["nms_6000"]["ms"], d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"]))
PY
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2>> "$OUT/bench.err"; cut -c1-300 "$OUT/bench_reference.json"
echo "== X2 one GPU"; [ "${3:-}" = "nox2" ] || timeout 900 python tools/x2_cfg5.py --steps 3 --warmup 1 > "$OUT/x2_n1.json" 2> "$OUT/x2_n1.err"; echo "x2 rc=$?" | tee -a "$OUT/status.txt"; cat "$OUT/x2_n1.json"; tail -2 "$OUT/x2_n1.err"
echo "== X1"; timeout 900 python tools/x1_cfg4.py --iters 5 > "$OUT/x1_cfg4.json" 2>/dev/null; cat "$OUT/x1_cfg4.json"
echo "== kernel matrix"; timeout 900 python tools/kernel_matrix.py --iters 30 > "$OUT/kernel_matrix.json" 2> "$OUT/kernel_matrix.log"; cat "$OUT/kernel_matrix.log"
echo "== fpn probe"; timeout 600 python tools/fpn_probe.py 2>&1 | tail -8
echo "== nms batched timing"; timeout 300 python - <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from detectron.pytorch_b200 import ops, synthetic as S
ten = [torch.from_numpy(S.make_nms_boxes(1000, seed=i)).cuda() for i in range(10)]
cat = torch.cat(ten); counts = [1000] * 10
for _ in range(3): ops.nms_batched_raw(cat, counts, 0.7)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): ops.nms_batched_raw(cat, counts, 0.7)
e1.record(); torch.cuda.synchronize(); print("10 x 1000 boxes, one batched call: %.1f us" % (e0.elapsed_time(e1) / 50 * 1e3))
e0.record()
for _ in range(50):
    for b in ten: ops.nms_raw(b, 0.7)
e1.record(); torch.cuda.synchronize(); print("10 x 1000 boxes, ten calls: %.1f us" % (e0.elapsed_time(e1) / 50 * 1e3))
PY
if [ "${2:-}" = "ncu" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$OUT/launches.csv" python bench.py --steps 4 --warmup 3 --no-graph --cpu-seconds 1 > "$OUT/ncu_launches.log" 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'roi_align_strip_fwd|strip_prep|roi_align_bwd_rows<' -s 9 -c 6 -o "$OUT/prof" -f python bench.py --steps 4 --warmup 3 --no-graph --cpu-seconds 1 > "$OUT/ncu_full.log" 2>&1
fi
cat "$OUT/status.txt"
