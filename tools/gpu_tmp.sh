timeout 300 python -m pytest tests/test_gpu_targets.py -m gpu -q -p no:cacheprovider -k "topk" 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "proposal or device_chain" 2>&1 | tail -5
echo "== proposals probe"; timeout 300 python tools/proposals_probe.py 2>&1 | tail -8
