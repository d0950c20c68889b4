"""bench.py's reference arm (`--impl reference`) needs no GPU: its JSON line must carry the driver's contract keys."""
import json
import os
import subprocess
import sys

from _util import ROOT


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, PYTHONPATH=ROOT)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "samples/s" and d["n_gpus"] == 1
    assert d["metric"].startswith("train samples/sec") and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("cfg5") and d["config"]["rows"] > 30_000_000 and d["config"]["batch_per_gpu"] == 8192
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # a non-zero rank of a torchrun launch exits 0 without output
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         env=dict(env, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2"), capture_output=True, text=True, timeout=300,
                         cwd=ROOT)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]
