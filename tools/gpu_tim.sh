#!/bin/bash
# timing-build timeline of the streaming forward + flakiness check + full GPU tests + bench of the default build
OUT=gpurun_out/${1:-tim}; mkdir -p $OUT
TIM=$PWD/detectron/pytorch_b200/libb200_roi_ops_tim.so
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); k=d["kernels"]
    print(sys.argv[2], "value %.0f fwd %.4f bwd %.4f nms %.4f e2e %.0f" % (d["value"], k["fwd"]["ms"], k["bwd"]["ms"], k["nms_6000"]["ms"], d["e2e"]["value"]))
except Exception as e: print("parse failed", e)
PY
}
echo "== fpn repro"; bash tools/gpu_fpnrepro.sh ${2:-16}
echo "== pytest -m gpu (full)"
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
echo "== timeline"
B200_ROI_OPS_LIB=$TIM timeout 300 python tools/stream_timing.py $OUT/timeline.json 2>&1 | python -c "
import json,sys
d=json.load(sys.stdin)
for k in ['cta_cycles','rows_per_piece','consumer','producer','fit_cycles']: print(k, d[k])
"
echo "== bench"
timeout 600 python bench.py --steps 100 --warmup 5 --cpu-seconds 1 > $OUT/bench.json 2> $OUT/bench.err; summ $OUT/bench.json default
