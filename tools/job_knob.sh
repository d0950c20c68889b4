# usage: bash tools/job_knob.sh VAR v1 v2 ...  -- bench.py --steps 200 for every value of an environment knob
cd "${GRAFT_REPO_ROOT:-/root/repo}"
var=$1; shift
for v in "$@"; do
for rep in 1 2; do
env $var=$v python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extra > gpurun_out/quick_bench.json 2>/dev/null
python - "$var=$v" <<PY
import json, sys
d = json.load(open("gpurun_out/quick_bench.json"))
print(sys.argv[1], "ms/step %.4f" % d["ms_per_step"], "value %.1fM" % (d["value"] / 1e6), "e2e %.1fM" % (d["e2e"]["value"] / 1e6))
PY
done
done
