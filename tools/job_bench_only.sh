cd "${GRAFT_REPO_ROOT:-/root/repo}"
for a in "--steps 20 --warmup 5" "--steps 20 --warmup 5" "--steps 200 --warmup 10" "--steps 20 --warmup 5"; do
python bench.py $a --no-cpu-baseline --no-extra > gpurun_out/quick_bench.json 2>gpurun_out/quick_bench.err || tail -3 gpurun_out/quick_bench.err
python - "$a" <<PY
import json, sys
d = json.load(open("gpurun_out/quick_bench.json"))
print(sys.argv[1], "ms/step %.4f" % d["ms_per_step"], "value %.1fM" % (d["value"] / 1e6), "e2e %.1fM" % (d["e2e"]["value"] / 1e6), "launches", d["gpu_launches"], "graphs", d["step_graphs_cached"])
PY
done
