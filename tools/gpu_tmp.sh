mkdir -p gpurun_out/r04n
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r04n/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r04n/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/r04n/bench.json 2> gpurun_out/r04n/bench.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r04n/bench.json")); k=d["kernels"]
print("value %.0f RoIs/s fwd %.4f (%.3f) bwd %.4f (%.3f) nms %.4f e2e %.0f launches/step %s clocks %s" % (d["value"], k["fwd"]["ms"], k["fwd"]["frac_of_measured"], k["bwd"]["ms"], k["bwd"]["frac_of_measured"], k["nms_6000"]["ms"], d["e2e"]["value"], d.get("launches_per_step"), d["clocks"]))
PY
timeout 900 python tools/kernel_matrix.py --iters 30 > gpurun_out/r04n/kernel_matrix.json 2> gpurun_out/r04n/kernel_matrix.log; grep -E "crop|pool|legacy|cfg2" gpurun_out/r04n/kernel_matrix.log
