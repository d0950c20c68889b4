cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests/test_gpu_fm.py -m gpu -x -q -k "sort_and_segments or large_batch_generic or sort_fields" 2>&1 | tail -3
for b in 131072 262144; do
python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline --no-extra > gpurun_out/quick_bench.json 2>gpurun_out/quick_bench.err || tail -3 gpurun_out/quick_bench.err
python - $b <<PY
import json, sys
d = json.load(open("gpurun_out/quick_bench.json"))
print("B", sys.argv[1], "ms/step %.4f" % d["ms_per_step"], "value %.1fM" % (d["value"] / 1e6), "e2e %.1fM" % (d["e2e"]["value"] / 1e6),
      {k: round(v * 1e3, 1) for k, v in d["roofline"]["phase_ms"].items()}, "frac %.3f" % d["roofline"]["frac"])
PY
done
