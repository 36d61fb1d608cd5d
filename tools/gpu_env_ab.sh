#!/usr/bin/env bash
# A/B of an environment switch: tools/gpu_env_ab.sh <tag> VAR=value
set -u
OUT=gpurun_out/${1:-envab}; mkdir -p "$OUT"; KV=$2
summ() { python -c "import sys,json; d=json.loads(open('$1').read()); k=d['kernels']; print('$2', 'RoIs/s %.0f step %.1f us fwd %.1f bwd %.1f nms %.1f' % (d['value'], d['ms_per_step']*1e3, k['fwd']['ms']*1e3, k['bwd']['ms']*1e3, k['nms_6000']['ms']*1e3))"; }
env $KV timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "roi_align and not rows_path" > "$OUT/pytest_var.log" 2>&1; echo "variant pytest rc=$?"; tail -2 "$OUT/pytest_var.log"
env $KV python bench.py --steps 100 --warmup 5 --cpu-seconds 1 > "$OUT/bench_var.json" 2> "$OUT/bench.err"; summ "$OUT/bench_var.json" "$KV"
python bench.py --steps 100 --warmup 5 --cpu-seconds 1 > "$OUT/bench_base.json" 2>> "$OUT/bench.err"; summ "$OUT/bench_base.json" base
tail -2 "$OUT/bench.err"
