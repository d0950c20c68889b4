# what the driver runs at round end on one GPU: smoke(), the reference arm, our arm (default flags)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/drv_ref.json 2> gpurun_out/drv_ref.err ) 2>&1 | grep real
tail -c 600 gpurun_out/drv_ref.json; echo
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/drv_ours.json 2> gpurun_out/drv_ours.err ) 2>&1 | grep real
python - <<PY
import json
d = json.loads(open("gpurun_out/drv_ours.json").read().strip().splitlines()[-1])
r = json.loads(open("gpurun_out/drv_ref.json").read().strip().splitlines()[-1])
print("ours: ms/step %.4f value %.1fM e2e %.1fM frac %.3f clocks %s" % (d["ms_per_step"], d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["roofline"]["frac"], d["clocks"]))
print("cpu_baseline", d.get("cpu_baseline")); print("cpu_baseline_reference", d.get("cpu_baseline_reference")); print("cfg4_fit", d.get("cfg4_fit"))
print("reference arm value %.3fM; e2e ratio %.1f" % (r["value"] / 1e6, d["e2e"]["value"] / r["value"]))
print(sorted(d.keys()))
PY
