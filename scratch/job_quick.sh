cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests/test_gpu_fm.py -m gpu -x -q 2>&1 | tail -3
python scratch/timeline_sort.py 2>&1 | grep -v "field 1[0-2]\|field  [1-8] " | tail -16
python scratch/timeline_step.py 2>&1 | tail -5
python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extra > gpurun_out/quick_bench.json 2>/dev/null
python - <<PY
import json
d = json.load(open("gpurun_out/quick_bench.json"))
print("ms/step %.4f" % d["ms_per_step"], "value %.1fM" % (d["value"] / 1e6), "e2e %.1fM" % (d["e2e"]["value"] / 1e6),
      {k: round(v * 1e3, 1) for k, v in d["roofline"]["phase_ms"].items()}, "frac %.3f" % d["roofline"]["frac"])
PY
