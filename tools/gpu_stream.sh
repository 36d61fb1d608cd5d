#!/usr/bin/env bash
# GPU session for the TMA streaming-strip forward: first-contact check (subprocess-isolated), parity subset, A/B bench, ncu.
#   gpurun --timeout 1500 -- 'bash tools/gpu_stream.sh <tag> [full]'
set -u
TAG=${1:-r03a}
MODE=${2:-quick}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.csv" 2>&1
echo "== stream_debug" | tee "$OUT/status.txt"
timeout 900 python tools/stream_debug.py > "$OUT/stream_debug.log" 2>&1; echo "stream_debug rc=$?" | tee -a "$OUT/status.txt"
cat "$OUT/stream_debug.log"
echo "== pytest stream subset"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "stream or cfg1_and_cfg2 or linearity or many_rois or fpn_equals" > "$OUT/pytest_stream.log" 2>&1; echo "pytest-stream rc=$?" | tee -a "$OUT/status.txt"
tail -15 "$OUT/pytest_stream.log"
echo "== bench (auto = stream)"
timeout 600 python bench.py --steps 100 --warmup 5 --cpu-seconds 2 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/status.txt"
python - "$OUT/bench.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); k=d["kernels"]
    print("value %.0f RoIs/s  ms/step %.4f  fwd %.4f ms (%.3f)  bwd %.4f ms (%.3f)  launches/step %s" % (d["value"], d["ms_per_step"], k["fwd"]["ms"], k["fwd"]["frac_of_measured"], k["bwd"]["ms"], k["bwd"]["frac_of_measured"], d.get("launches_per_step")))
except Exception as e: print("bench parse failed", e)
PY
tail -3 "$OUT/bench.err"
echo "== bench (tiled forward, A/B)"
B200_ROI_ALIGN_PATH=tiled timeout 600 python bench.py --steps 100 --warmup 5 --cpu-seconds 1 > "$OUT/bench_tiled.json" 2>> "$OUT/bench.err"
python - "$OUT/bench_tiled.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); k=d["kernels"]
    print("tiled: fwd %.4f ms  bwd %.4f ms" % (k["fwd"]["ms"], k["bwd"]["ms"]))
except Exception as e: print("bench parse failed", e)
PY
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
if [ "$MODE" = "full" ]; then
  echo "== ncu full (stream fwd)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'roi_align_stream_fwd' -s 4 -c 2 -o "$OUT/prof_stream" -f \
      python bench.py --steps 4 --warmup 3 --no-graph --cpu-seconds 1 > "$OUT/ncu_full.log" 2>&1; echo "ncu-full rc=$?" | tee -a "$OUT/status.txt"
fi
ls -la "$OUT"
