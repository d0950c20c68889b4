# usage: bash tools/job_multi.sh <N> <tag>  -- multi-GPU bench lines (run list on / off for the owner-side run kernel)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=$1; tag=$2
for rl in 1 0; do
FMB_SHARD_RL=$rl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 50 --warmup 10 > gpurun_out/${tag}_n${N}_rl${rl}.json 2> gpurun_out/${tag}_n${N}_rl${rl}.err || tail -5 gpurun_out/${tag}_n${N}_rl${rl}.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_n${N}_rl${rl}.json").read().strip().splitlines()[-1])
    print("N=$N RL=$rl ms/step %.4f value %.1fM" % (d["ms_per_step"], d["value"] / 1e6), d.get("config", {}).get("exchange"), d.get("equivalence"))
except Exception as e:
    print("N=$N RL=$rl failed:", e)
PY
done
