# new work of this session: dataset pipeline tests, single-pass hedge test, cfg4 records
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests/test_dataset.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_deep.py -m gpu -x -q -k "hedge" 2>&1 | tail -15
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/new_bench.json 2> gpurun_out/new_bench.err || tail -5 gpurun_out/new_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/new_bench.json"))
print("ms/step %.4f" % d["ms_per_step"], "e2e %.1fM" % (d["e2e"]["value"] / 1e6))
c = d["cfg4_fit"]; print("cfg4 fit ms", c.get("ms_per_step"), "hedge", c.get("hedge_fit"))
PY
