# usage: bash tools/job_shard3.sh <N> <tag>  -- fused sharded step: equivalence with the NCCL step over a real process group
# (eager and graph), then bench lines of the fused and of the three-kernel path
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=$1; tag=$2
for g in "" "--graph"; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tests/sharded_pipeline_check.py --fused $g 2>&1 | grep -E "PIPELINE_CHECK_OK|Error|error|assert" | head -5
done
for ex in fused peers; do
FMB_SHARD_EXCHANGE=$ex timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 50 --warmup 10 > gpurun_out/${tag}_n${N}_${ex}.json 2> gpurun_out/${tag}_n${N}_${ex}.err || tail -5 gpurun_out/${tag}_n${N}_${ex}.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_n${N}_${ex}.json").read().strip().splitlines()[-1])
    print("N=$N $ex ms/step %.4f value %.1fM e2e %.1fM" % (d["ms_per_step"], d["value"] / 1e6, d["e2e"]["value"] / 1e6))
except Exception as e:
    print("N=$N $ex failed:", e)
PY
done
