#!/usr/bin/env bash
# One GPU-box session: parity tests, golden generation from the reference kernels, bench (both arms),
# ncu launch list + one full capture of the RoIAlign kernels.  Everything lands in gpurun_out/.
#   gpurun --timeout 1800 -- 'bash tools/gpu_session.sh [tag]'
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/gpu.csv" 2>&1
echo "== smoke" | tee "$OUT/status.txt"
python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/status.txt"
echo "== pytest -m gpu" | tee -a "$OUT/status.txt"
python -m pytest tests -m gpu -q -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/status.txt"
tail -5 "$OUT/pytest_gpu.log"
echo "== golden"
python tests/golden/make_golden.py "$OUT/golden" > "$OUT/golden.log" 2>&1; echo "golden rc=$?" | tee -a "$OUT/status.txt"
echo "== bench"
python bench.py --steps 200 --warmup 10 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/status.txt"
cat "$OUT/bench.json"
python bench.py --steps 200 --warmup 10 --no-graph > "$OUT/bench_nograph.json" 2>> "$OUT/bench.err"
python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2>> "$OUT/bench.err"; echo "bench-ref rc=$?" | tee -a "$OUT/status.txt"
cat "$OUT/bench_reference.json"
echo "== ncu launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 4 --warmup 3 --no-graph > "$OUT/ncu_launches.log" 2>&1; echo "ncu-list rc=$?" | tee -a "$OUT/status.txt"
echo "== ncu full (our roi_align kernels)"
ncu --set full --clock-control none --import-source on -k regex:roi_align -s 6 -c 4 -o "$OUT/prof_roi_align" -f \
    python bench.py --steps 4 --warmup 3 --no-graph > "$OUT/ncu_full.log" 2>&1; echo "ncu-full rc=$?" | tee -a "$OUT/status.txt"
ls -la "$OUT"
