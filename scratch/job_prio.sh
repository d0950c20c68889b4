cd "${GRAFT_REPO_ROOT:-/root/repo}"
A="--no-cpu-baseline --no-extra"
for pr in 1 0 2; do
FMB_PRIO=$pr python bench.py --steps 200 --warmup 10 $A > gpurun_out/prio$pr.json 2>/dev/null
python - $pr <<PY
import json, sys
d = json.load(open("gpurun_out/prio%s.json" % sys.argv[1]))
print("FMB_PRIO", sys.argv[1], "ms/step %.4f" % d["ms_per_step"], "value %.1fM" % (d["value"] / 1e6), "e2e %.1fM" % (d["e2e"]["value"] / 1e6),
      {k: round(v * 1e3, 1) for k, v in d["roofline"]["phase_ms"].items()})
PY
done
