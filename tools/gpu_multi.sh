#!/usr/bin/env bash
# Multi-GPU check of bench.py exactly as the driver launches it.   gpurun --gpus N -- 'bash tools/gpu_multi.sh N tag'
set -u
N=${1:-2}; TAG=${2:-multi}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpus.txt" 2>&1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 100 --warmup 5 > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"; echo "bench N=$N rc=$?" | tee "$OUT/status.txt"
tail -1 "$OUT/bench_n$N.json" | cut -c1-900; tail -3 "$OUT/bench_n$N.err"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > "$OUT/bench_ref_n$N.json" 2>> "$OUT/bench_n$N.err"; echo "bench-ref N=$N rc=$?" | tee -a "$OUT/status.txt"
tail -1 "$OUT/bench_ref_n$N.json" | cut -c1-300
