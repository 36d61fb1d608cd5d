#!/usr/bin/env bash
# NMS-only GPU check: parity tests for NMS + probe timing of the scan variants.
set -u
OUT=gpurun_out/${1:-nms}; mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k nms > "$OUT/pytest_nms.log" 2>&1; echo "pytest nms rc=$?"; tail -3 "$OUT/pytest_nms.log"
timeout 120 python tools/nms_probe.py | tee "$OUT/nms_probe.txt"
B200_NMS_SCAN=simple timeout 120 python tools/nms_probe.py | tee -a "$OUT/nms_probe.txt"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file "$OUT/nms_launches.csv" python tools/nms_probe.py > /dev/null 2>&1
grep -o 'b200::nms[a-z_]*\|"[0-9]*"$' "$OUT/nms_launches.csv" | paste - - | sort | uniq -c | sort -rn | head -8
