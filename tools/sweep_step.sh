cd "${GRAFT_REPO_ROOT:-/root/repo}"
for sb in 8; do for ps in 0 1; do
FMB_STEP_SB=$sb FMB_PRESORT_STANDALONE=$ps python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extra 2>/dev/null > /tmp/o.json
SB=$sb PS=$ps python - <<'PY'
import json, os
d = json.load(open("/tmp/o.json"))
print("SB", os.environ["SB"], "standalone", os.environ["PS"], "us/step %.1f" % (d["ms_per_step"] * 1e3),
      "fused %.1f" % (d["roofline"]["phase_ms"]["fm_step_fused"] * 1e3), "host %.1f" % (d["host_submit_ms_per_step"] * 1e3),
      "e2e %.1fM" % (d["e2e"]["value"] / 1e6))
PY
done; done
