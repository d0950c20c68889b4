# round-2 profiling pass of the single-GPU step (run under gpurun, one GPU): launch list, one full capture of the
# step's kernels, and the large-batch bench line.  Outputs under gpurun_out/r2h_*.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
A="--no-cpu-baseline --no-extra"
python bench.py --steps 20 --warmup 5 $A > gpurun_out/r2h_plain.json 2> gpurun_out/r2h_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fm_|sort|finish|pos_flags" -s 150 -c 60 --csv \
    --log-file gpurun_out/r2h_launches.csv python bench.py --steps 20 --warmup 5 $A > gpurun_out/r2h_ncu.log 2>&1
python bench.py --steps 6 --warmup 3 $A > gpurun_out/r2h_plain2.json 2>/dev/null &&
ncu --set full --clock-control none --import-source on -k regex:"fm_step_fused|fm_bwd_runs|sort_fields|pos_flags|finish_step" \
    -s 100 -c 5 -o gpurun_out/r2h_prof -f python bench.py --steps 6 --warmup 3 $A > gpurun_out/r2h_ncu2.log 2>&1
tail -1 gpurun_out/r2h_ncu2.log
python bench.py --steps 20 --warmup 5 --batch 65536 $A > gpurun_out/r2h_bench_b65536.json 2>/dev/null
python bench.py --steps 20 --warmup 5 $A --workload cfg4 > gpurun_out/r2h_bench_cfg4.json 2>/dev/null
python - <<'PY'
import json
for f in ("r2h_plain", "r2h_bench_b65536", "r2h_bench_cfg4"):
    d = json.load(open("gpurun_out/%s.json" % f))
    print(f, "ms/step %.4f" % d["ms_per_step"], "value %.1fM" % (d["value"] / 1e6), "e2e %.1fM" % (d["e2e"]["value"] / 1e6),
          {k: round(v * 1e3, 1) for k, v in d["roofline"]["phase_ms"].items()}, "frac %.3f" % d["roofline"]["frac"])
PY
