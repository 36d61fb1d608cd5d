#!/usr/bin/env bash
OUT=gpurun_out/${1:-repro}; mkdir -p $OUT
echo "== plain P=14"; timeout 120 python tools/stream_repro.py 2 64 200 336 0.25 256 14 2 20 2>&1 | tail -2
echo "== plain tiny rois"; timeout 120 python tools/stream_repro.py 2 64 200 336 0.25 1000 7 2 5 4 112 2>&1 | tail -2
echo "== plain sparse tall"; timeout 120 python tools/stream_repro.py 1 32 600 64 0.25 12 7 2 7 8 64 2>&1 | tail -2
echo "== pytest subset"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "stream or reference_model or nms_batched or launchers or proposals" 2>&1 | tail -8
echo "== x1"; timeout 900 python tools/x1_cfg4.py --iters 5 2>/dev/null | tail -1
echo "== matrix"; timeout 900 python tools/kernel_matrix.py --iters 30 > $OUT/kernel_matrix.json 2> $OUT/kernel_matrix.log; cat $OUT/kernel_matrix.log
echo "== proposals probe"; timeout 600 python tools/proposals_probe.py 2>&1 | tail -6
echo "== bench"; timeout 600 python bench.py --steps 100 --warmup 5 --cpu-seconds 2 > $OUT/bench.json 2>/dev/null; python - $OUT/bench.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); k=d["kernels"]
print("value %.0f fwd %.4f bwd %.4f nms %.4f" % (d["value"], k["fwd"]["ms"], k["bwd"]["ms"], k["nms_6000"]["ms"]))
PY
