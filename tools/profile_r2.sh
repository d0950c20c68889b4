# round-2 profiling pass of the single-GPU step (run under gpurun, one GPU): launch list, one full capture of the
# step's kernels, and the bench lines.  Outputs under gpurun_out/${TAG}_*.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r2p}
A="--no-cpu-baseline --no-extra"
K="fm_step_fused|fm_bwd_runs|sort_fields|sparse_fields|finish_step"
python bench.py --steps 20 --warmup 5 $A > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -s 150 -c 60 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 5 $A > gpurun_out/${TAG}_ncu.log 2>&1
python bench.py --steps 6 --warmup 3 $A > gpurun_out/${TAG}_plain2.json 2>/dev/null &&
ncu --set full --clock-control none --import-source on -k regex:"$K" \
    -s 100 -c 6 -o gpurun_out/${TAG}_prof -f python bench.py --steps 6 --warmup 3 $A > gpurun_out/${TAG}_ncu2.log 2>&1
tail -1 gpurun_out/${TAG}_ncu2.log
python bench.py --steps 200 --warmup 10 $A > gpurun_out/${TAG}_bench200.json 2>/dev/null
python bench.py --steps 20 --warmup 5 --batch 65536 $A > gpurun_out/${TAG}_bench_b65536.json 2>gpurun_out/${TAG}_b65536.err || tail -3 gpurun_out/${TAG}_b65536.err
python bench.py --steps 20 --warmup 5 $A --workload cfg4 > gpurun_out/${TAG}_bench_cfg4.json 2>gpurun_out/${TAG}_cfg4.err || tail -3 gpurun_out/${TAG}_cfg4.err
python - <<PY
import json
for f in ("${TAG}_plain", "${TAG}_bench200", "${TAG}_bench_b65536", "${TAG}_bench_cfg4"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
    except Exception as e:
        print(f, "failed", e); continue
    print(f, "ms/step %.4f" % d["ms_per_step"], "value %.1fM" % (d["value"] / 1e6), "e2e %.1fM" % (d["e2e"]["value"] / 1e6),
          {k: round(v * 1e3, 1) for k, v in d["roofline"]["phase_ms"].items()}, "frac %.3f" % d["roofline"]["frac"])
PY
