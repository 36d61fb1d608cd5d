mkdir -p gpurun_out/r04i
timeout 400 python -m pytest tests/test_gpu_targets.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "tolerance_is_derived or full_size" 2>&1 | tail -3
cp gpurun_out/parity_spread.json gpurun_out/r04i/ 2>/dev/null; cat gpurun_out/parity_spread.json
echo "== default (IF8)"; timeout 200 python tools/fwd_ab.py --paths quad --shapes cfg2,p2box --iters 100 2>&1 | grep -v "^{"
for lib in detectron/pytorch_b200/libvar_*.so; do echo "== $lib"; B200_ROI_OPS_LIB=$PWD/$lib timeout 120 python tools/fwd_ab.py --paths quad --shapes cfg2,p2box --iters 100 2>&1 | grep -v "^{"; done
