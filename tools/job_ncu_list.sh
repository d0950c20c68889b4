# usage: bash tools/job_ncu_list.sh <tag>  -- bench line + ncu launch list (per-kernel durations, serialised, cold cache)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
tag=$1
A="--no-cpu-baseline --no-extra"
python bench.py --steps 20 --warmup 5 $A > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || tail -5 gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench.json"))
print("ms/step %.4f" % d["ms_per_step"], "value %.1fM" % (d["value"] / 1e6), "e2e %.1fM" % (d["e2e"]["value"] / 1e6),
      {k: round(v * 1e3, 1) for k, v in d["roofline"]["phase_ms"].items()}, "frac %.3f" % d["roofline"]["frac"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fm_|sort|finish|pos_flags" -s 150 -c 40 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 20 --warmup 5 $A > gpurun_out/${tag}_ncu.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${tag}_launches.csv")) if len(r) > 10]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    agg[r[ik][:60]].append(float(r[iv].replace(",", "")))
for k, v in agg.items():
    print("%-62s n=%3d mean %.1f us" % (k, len(v), sum(v) / len(v) / 1e3))
PY
