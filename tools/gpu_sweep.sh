#!/usr/bin/env bash
# Sweep of quad-strip build variants (libvar_*.so built with B200_NVCC_EXTRA / B200_ROI_OPS_LIB) and run-time options.
#   gpurun --timeout 1200 -- 'bash tools/gpu_sweep.sh <tag>'
set -u
TAG=${1:-r04c}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
for lib in detectron/pytorch_b200/libvar_*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib" | tee -a "$OUT/sweep.log"
  B200_ROI_OPS_LIB=$PWD/$lib timeout 120 python tools/fwd_ab.py --paths quad --shapes cfg2,p2box --iters 100 2>&1 | grep -v "^{" | tee -a "$OUT/sweep.log"
done
echo "== options (default lib)" | tee -a "$OUT/sweep.log"
for rc in 0 2 5 9; do
  echo "rowcost $rc" | tee -a "$OUT/sweep.log"
  B200_STRIP_ROWCOST=$rc timeout 120 python tools/fwd_ab.py --paths quad --shapes cfg2 --iters 100 2>&1 | grep -v "^{" | tee -a "$OUT/sweep.log"
done
echo "pdl off" | tee -a "$OUT/sweep.log"
B200_STRIP_PDL=0 timeout 120 python tools/fwd_ab.py --paths quad --shapes cfg2 --iters 100 2>&1 | grep -v "^{" | tee -a "$OUT/sweep.log"
echo "== pytest (quad / stream / fpn subset)"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "nothing_selected_here" > "$OUT/pytest_quad.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/status.txt"
tail -8 "$OUT/pytest_quad.log"
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$OUT/launches.csv" \
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
echo "== ncu full (strip fwd)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'roi_align_strip_fwd' -s 3 -c 1 -o "$OUT/prof" -f \
    python bench.py --steps 4 --warmup 3 --no-graph --cpu-seconds 1 > "$OUT/ncu_full.log" 2>&1; echo "ncu-full rc=$?" | tee -a "$OUT/status.txt"
