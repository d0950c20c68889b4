cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests/test_gpu_fm.py -m gpu -x -q 2>&1 | tail -3
python scratch/timeline_sort.py 2>&1 | grep "flags\|field 38\|field  0" | tail -5
for sa in 1 0; do
FMB_SORT_AFTER=$sa python scratch/timeline_step.py 2>&1 | tail -5
FMB_SORT_AFTER=$sa python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extra > gpurun_out/quick_bench$sa.json 2>/dev/null
python - $sa <<PY
import json, sys
d = json.load(open("gpurun_out/quick_bench%s.json" % sys.argv[1]))
print("SORT_AFTER", sys.argv[1], "ms/step %.4f" % d["ms_per_step"], "value %.1fM" % (d["value"] / 1e6), "e2e %.1fM" % (d["e2e"]["value"] / 1e6),
      {k: round(v * 1e3, 1) for k, v in d["roofline"]["phase_ms"].items()}, "frac %.3f" % d["roofline"]["frac"])
PY
done
