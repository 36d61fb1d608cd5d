#!/usr/bin/env bash
# One broad GPU session: bench A/B, full parity suite, kernel matrix, sanitizer passes, cfg-4 harness, ncu.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh <tag> [noncu]'
set -u
TAG=${1:-r03e}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.csv" 2>&1
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); k=d["kernels"]
    print("%-12s value %.0f RoIs/s  ms/step %.4f  fwd %.4f ms (%.3f)  bwd %.4f ms (%.3f)  e2e %.0f  launches/step %s" % (sys.argv[2], d["value"], d["ms_per_step"], k["fwd"]["ms"], k["fwd"]["frac_of_measured"], k["bwd"]["ms"], k["bwd"]["frac_of_measured"], d["e2e"]["value"], d.get("launches_per_step")))
except Exception as e: print(sys.argv[2], "bench parse failed", e)
PY
}
echo "== bench A/B" | tee "$OUT/status.txt"
timeout 600 python bench.py --steps 100 --warmup 5 --cpu-seconds 2 > "$OUT/bench.json" 2> "$OUT/bench.err"; summ "$OUT/bench.json" default
B200_STREAM_PHASES=prepass timeout 600 python bench.py --steps 100 --warmup 5 --cpu-seconds 1 > "$OUT/bench_prepass.json" 2>> "$OUT/bench.err"; summ "$OUT/bench_prepass.json" prepass-only
B200_ROI_ALIGN_PATH=tiled timeout 600 python bench.py --steps 100 --warmup 5 --cpu-seconds 1 > "$OUT/bench_tiled.json" 2>> "$OUT/bench.err"; summ "$OUT/bench_tiled.json" tiled-fwd
tail -3 "$OUT/bench.err"
echo "== pytest -m gpu (full)"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/status.txt"
tail -15 "$OUT/pytest_gpu.log"
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/status.txt"; tail -2 "$OUT/smoke.log"
echo "== cfg4 harness (X1)"
timeout 900 python tools/x1_cfg4.py --iters 5 > "$OUT/x1_cfg4.json" 2> "$OUT/x1_cfg4.err"; echo "x1 rc=$?" | tee -a "$OUT/status.txt"; cat "$OUT/x1_cfg4.json"; tail -3 "$OUT/x1_cfg4.err"
echo "== kernel matrix"
timeout 900 python tools/kernel_matrix.py --iters 30 > "$OUT/kernel_matrix.json" 2> "$OUT/kernel_matrix.log"; echo "matrix rc=$?" | tee -a "$OUT/status.txt"; cat "$OUT/kernel_matrix.log"
echo "== compute-sanitizer (memcheck + racecheck) on the new kernels, small shapes"
cat > /tmp/san_case.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from detectron.pytorch_b200 import _lib, ops, synthetic as S
from detectron.pytorch_b200.modeling.roi_xfrom.roi_align.functions.roi_align import RoIAlignFunction
shape, s = (2, 64, 40, 100), 1.0 / 8
f = torch.from_numpy(S.make_features(shape, seed=1)).cuda().requires_grad_(True)
r = torch.from_numpy(np.concatenate([S.make_rois(60, shape, s, seed=2), S.make_edge_rois(shape, s)]).astype(np.float32)).cuda()
_lib.set_option("B200_ROI_ALIGN_PATH", "stream")
out = RoIAlignFunction(7, 7, s, 2)(f, r)
out.backward(torch.ones_like(out))
b = torch.from_numpy(S.make_nms_boxes(700, seed=3)).cuda()
ops.nms_raw(b, 0.7)
ops.nms_batched_raw(torch.cat([b, b[:300]]), [700, 300], 0.7)
torch.cuda.synchronize()
print("sanitizer case done")
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_case.py > "$OUT/sanitizer_$tool.log" 2>&1; echo "sanitizer-$tool rc=$?" | tee -a "$OUT/status.txt"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer case done|Error|hazard" "$OUT/sanitizer_$tool.log" | head -8
done
if [ "${2:-}" != "noncu" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 4 --warmup 3 --no-graph --cpu-seconds 1 > "$OUT/ncu_launches.log" 2>&1; echo "ncu-list rc=$?" | tee -a "$OUT/status.txt"
python - "$OUT/launches.csv" <<'PY'
import csv,sys,collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=None; agg=collections.defaultdict(list)
for r in rows:
    if r[0]=="ID": hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r))
    try: agg[d["Kernel Name"][:60]].append(float(d["Metric Value"].replace(",","")))
    except Exception: pass
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print("%-62s n=%3d mean=%.1f us" % (k, len(v), sum(v)/len(v)/1e3))
PY
echo "== ncu full (stream fwd + bwd rows)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'roi_align_stream_fwd|roi_align_bwd_rows<' -s 8 -c 2 -o "$OUT/prof" -f \
    python bench.py --steps 4 --warmup 3 --no-graph --cpu-seconds 1 > "$OUT/ncu_full.log" 2>&1; echo "ncu-full rc=$?" | tee -a "$OUT/status.txt"
fi
cat "$OUT/status.txt"
