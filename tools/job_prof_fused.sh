# one `ncu --set full` capture of the fused step kernel at B = 8 192 and at B = 65 536 (source page included), after
# plain runs of the same commands.  Outputs under gpurun_out/${TAG}_*.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r2f}
A="--no-cpu-baseline --no-extra"
python bench.py --steps 20 --warmup 5 $A > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || tail -5 gpurun_out/${TAG}_plain.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"fm_step_fused" -s 40 -c 1 -o gpurun_out/${TAG}_fused8k -f \
    python bench.py --steps 6 --warmup 3 $A > gpurun_out/${TAG}_ncu8k.log 2>&1
tail -1 gpurun_out/${TAG}_ncu8k.log
python bench.py --steps 20 --warmup 5 --batch 65536 $A > gpurun_out/${TAG}_b65536.json 2> gpurun_out/${TAG}_b65536.err || tail -5 gpurun_out/${TAG}_b65536.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"fm_step_fused" -s 20 -c 1 -o gpurun_out/${TAG}_fused64k -f \
    python bench.py --steps 6 --warmup 3 --batch 65536 $A > gpurun_out/${TAG}_ncu64k.log 2>&1
tail -1 gpurun_out/${TAG}_ncu64k.log
python - <<PY
import json
for f in ("${TAG}_plain", "${TAG}_b65536"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
    except Exception as e:
        print(f, "failed", e); continue
    print(f, "ms/step %.4f" % d["ms_per_step"], "value %.1fM" % (d["value"] / 1e6), "e2e %.1fM" % (d["e2e"]["value"] / 1e6),
          {k: round(v * 1e3, 1) for k, v in d["roofline"]["phase_ms"].items()}, "frac %.3f" % d["roofline"]["frac"])
PY
