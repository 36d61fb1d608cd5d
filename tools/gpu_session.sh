#!/usr/bin/env bash
# One GPU-box session: parity tests, bench (both arms), ncu launch list + one full capture.
#   gpurun --timeout 1800 -- 'bash tools/gpu_session.sh [tag] [quick]'
set -u
TAG=${1:-r01}
MODE=${2:-full}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/gpu.csv" 2>&1
echo "== smoke" | tee "$OUT/status.txt"
python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/status.txt"
echo "== pytest -m gpu" | tee -a "$OUT/status.txt"
python -m pytest tests -m gpu -q -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/status.txt"
tail -25 "$OUT/pytest_gpu.log"
echo "== bench"
python bench.py --steps 200 --warmup 10 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/status.txt"
cat "$OUT/bench.json"; tail -3 "$OUT/bench.err"
if [ "$MODE" = "full" ]; then
  B200_ROI_ALIGN_PATH=generic python bench.py --steps 100 --warmup 10 > "$OUT/bench_generic.json" 2>> "$OUT/bench.err"
  python bench.py --steps 200 --warmup 10 --no-graph > "$OUT/bench_nograph.json" 2>> "$OUT/bench.err"
  python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2>> "$OUT/bench.err"; echo "bench-ref rc=$?" | tee -a "$OUT/status.txt"
  cat "$OUT/bench_reference.json"
fi
echo "== phase probe (needs B200_NVCC_EXTRA=-DB200_FWD_TIMING build)"
echo "== ubench"; true
echo "== ncu nms"; ncu --set full --clock-control none --import-source on -k regex:nms_ -s 4 -c 2 -o "$OUT/prof_nms" -f python tools/nms_probe.py > "$OUT/ncu_nms.log" 2>&1; tail -2 "$OUT/ncu_nms.log"
echo "== ncu launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 4 --warmup 3 --no-graph --cpu-seconds 1 > "$OUT/ncu_launches.log" 2>&1; echo "ncu-list rc=$?" | tee -a "$OUT/status.txt"
echo "== ncu full (our kernels)"
ncu --set full --clock-control none --import-source on -k regex:'roi_align_tiled|roi_align_bwd_rows' -s 8 -c 5 -o "$OUT/prof" -f \
    python bench.py --steps 4 --warmup 3 --no-graph --cpu-seconds 1 > "$OUT/ncu_full.log" 2>&1; echo "ncu-full rc=$?" | tee -a "$OUT/status.txt"
ls -la "$OUT"
