"""whole-step GPU time (CUDA events behind a sleep kernel, eager) of the sharded step variants under torchrun."""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from fm_for_online_recommendation_b200 import sharded as sh
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sizes = bench.feature_sizes("cfg5"); B = 8192; k = 10
m = sh.ShardedFM(sizes, k, n=1e-4, seed=0)
host = bench.synth_batches(sizes, B, 4, 1234 + rank)
enc = [m.encode(Xi, Y) for Xi, Y in host]
st = torch.cuda.current_stream()
def timeit(fn, reps=30):
    tot = 0.0
    for it in range(reps + 5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        torch.cuda._sleep(4_000_000)
        e0.record(st); fn(it); e1.record(st)
        torch.cuda.synchronize()
        if it >= 5: tot += e0.elapsed_time(e1) * 1000
    t = torch.tensor([tot / reps], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t.item()), 1)
res = {}
res["unpipelined"] = timeit(lambda i: m.update_embedding(*enc[i % 4]))
m.prepare(enc[0][0])
res["pipelined"] = timeit(lambda i: m.update_embedding_pipelined(enc[i % 4][1], enc[(i + 1) % 4][0]))
# pipelined without preparing a next batch (what the step costs when the ids work is entirely off the path)
m.prepare(enc[0][0])
def no_next(i):
    m._slot = 0
    m.update_embedding_pipelined(enc[0][1], None)
res["pipelined_no_next"] = timeit(no_next)
if rank == 0: print(json.dumps({"world": world, "us": res}), flush=True)
torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush(); os._exit(0)
