#!/usr/bin/env bash
bash tools/gpu_dbg.sh "$1_dbg"
if grep -q "FAILED\|TIMEOUT" "gpurun_out/$1_dbg/dbg.log"; then echo "debug cases failed: skipping the full session"; exit 1; fi
bash tools/gpu_stream.sh "$1" "${2:-quick}"
