cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 800 python -m pytest tests/test_gpu_deep.py tests/test_gpu_main_experiment.py tests/test_trajectory.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/quick_bench.json 2>gpurun_out/quick_bench.err || tail -3 gpurun_out/quick_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/quick_bench.json"))
print("ms/step %.4f value %.1fM e2e %.1fM" % (d["ms_per_step"], d["value"] / 1e6, d["e2e"]["value"] / 1e6), "cfg4_fit", d.get("cfg4_fit", {}).get("ms_per_step"), d.get("cfg4_fit", {}).get("value"))
PY
