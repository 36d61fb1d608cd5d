#!/usr/bin/env bash
# RoIAlign-only GPU check: parity tests + bench with both backward variants.
set -u
OUT=gpurun_out/${1:-bwd}; mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "roi_align" > "$OUT/pytest.log" 2>&1; echo "pytest rc=$?"; tail -4 "$OUT/pytest.log"
summ() { python -c "import sys,json; d=json.loads(open('$1').read()); k=d['kernels']; print('$2', 'RoIs/s %.0f step %.1f us fwd %.1f bwd %.1f e2e %.0f (%.2f ms)' % (d['value'], d['ms_per_step']*1e3, k['fwd']['ms']*1e3, k['bwd']['ms']*1e3, d['e2e']['value'], d['e2e']['ms_per_step']))"; }
python bench.py --steps 100 --warmup 5 --cpu-seconds 1 > "$OUT/bench_cpl4.json" 2> "$OUT/bench.err"; summ "$OUT/bench_cpl4.json" rows-cpl4
B200_ROI_ALIGN_BWD_CPL=2 python bench.py --steps 100 --warmup 5 --cpu-seconds 1 > "$OUT/bench_cpl2.json" 2>> "$OUT/bench.err"; summ "$OUT/bench_cpl2.json" rows-cpl2
B200_ROI_ALIGN_BWD_PATH=nhwc python bench.py --steps 100 --warmup 5 --cpu-seconds 1 > "$OUT/bench_nhwc.json" 2>> "$OUT/bench.err"; summ "$OUT/bench_nhwc.json" nhwc
tail -3 "$OUT/bench.err"
