#!/usr/bin/env bash
# Focused GPU session for the quad-strip forward: A/B timing, its parity tests, bench, launch list, one full ncu capture.
#   gpurun --timeout 1200 -- 'bash tools/gpu_quad.sh <tag> [tests|notests] [ncu|noncu]'
set -u
TAG=${1:-r04a}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.csv" 2>&1
echo "== A/B" | tee "$OUT/status.txt"
timeout 300 python tools/fwd_ab.py --check --paths quad,stream --shapes cfg2,p2box,p2mask,p4box > "$OUT/fwd_ab.log" 2>&1; echo "ab rc=$?" | tee -a "$OUT/status.txt"
grep -v "^{" "$OUT/fwd_ab.log" | tail -12
if [ "${2:-tests}" = "tests" ]; then
  echo "== pytest (quad / stream / fpn subset)"
  timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "quad or stream or fpn or device_chain" > "$OUT/pytest_quad.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/status.txt"
  tail -8 "$OUT/pytest_quad.log"
fi
echo "== bench"
timeout 600 python bench.py --steps 100 --warmup 5 --cpu-seconds 2 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/status.txt"
python - "$OUT/bench.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); k=d["kernels"]
    print("value %.0f RoIs/s  ms/step %.4f  fwd %.4f ms (%.3f)  bwd %.4f ms (%.3f)  e2e %.0f  launches/step %s" % (d["value"], d["ms_per_step"], k["fwd"]["ms"], k["fwd"]["frac_of_measured"], k["bwd"]["ms"], k["bwd"]["frac_of_measured"], d["e2e"]["value"], d.get("launches_per_step")))
except Exception as e: print("bench parse failed", e)
PY
tail -3 "$OUT/bench.err"
if [ "${3:-ncu}" = "ncu" ]; then
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
echo "== ncu full (strip fwd + prep)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'roi_align_strip_fwd|strip_prep' -s 6 -c 2 -o "$OUT/prof" -f \
    python bench.py --steps 4 --warmup 3 --no-graph --cpu-seconds 1 > "$OUT/ncu_full.log" 2>&1; echo "ncu-full rc=$?" | tee -a "$OUT/status.txt"
fi
cat "$OUT/status.txt"
