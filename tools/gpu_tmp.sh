timeout 400 python -m pytest tests/test_gpu_targets.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "fpn or device_chain or quad" 2>&1 | tail -3
echo "== fwd A/B"; timeout 200 python tools/fwd_ab.py --paths quad --shapes cfg2,p2box --iters 100 2>&1 | grep -v "^{"
echo "== fpn probe"; timeout 300 python tools/fpn_probe.py 2>&1 | tail -6
