#!/usr/bin/env bash
# debug build of the streaming forward (progress markers in pinned host memory): first-contact cases only
set -u
TAG=${1:-dbg}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
for c in tiny small c40_n2 cfg2; do
  B200_ROI_OPS_LIB=$PWD/detectron/pytorch_b200/libb200_roi_ops_dbg.so timeout 120 python tools/stream_debug.py one $c 2>&1 | grep -v "^  File\|^    \|Traceback\|^Search\|^CUDA kernel\|^For debugging\|^Compile with" | tee -a "$OUT/dbg.log"
done
