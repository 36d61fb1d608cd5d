#!/usr/bin/env bash
# N-GPU session: the cfg-5 train step (row X2) under torchrun + the headline bench exactly as the driver launches it
N=${1:-2}; TAG=${2:-r03g}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name --format=csv | head -10
echo "== X2, $N GPUs"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/x2_cfg5.py --steps 5 --warmup 2 > "$OUT/x2_n$N.json" 2> "$OUT/x2_n$N.err"; echo "x2 rc=$?"
grep '^{' "$OUT/x2_n$N.json" | tail -1; tail -3 "$OUT/x2_n$N.err"
echo "== bench, $N GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 100 --warmup 5 > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"; echo "bench rc=$?"
python - "$OUT/bench_n$N.json" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("n_gpus %d value %.0f RoIs/s ms/step %.4f e2e %.0f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e: print("parse failed", e)
PY
