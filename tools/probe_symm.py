import os, sys, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok_symm = False
try:
    import torch.distributed._symmetric_memory as sm
    t = sm.empty(1024, dtype=torch.float32, device=f"cuda:{local}")
    h = sm.rendezvous(t, dist.group.WORLD)
    t.fill_(float(rank + 1)); torch.cuda.synchronize(); dist.barrier()
    peer = h.get_buffer((rank + 1) % world, (1024,), torch.float32)
    v = float(peer[0].item())
    print(f"[{rank}] symm ok: peer value {v}, ptrs {[hex(p) for p in h.buffer_ptrs]}, signal_pad_size {h.signal_pad_size}", flush=True)
    # a kernel of this process/device storing straight into the peer's buffer
    peer.fill_(100.0 + rank); torch.cuda.synchronize(); dist.barrier()
    print(f"[{rank}] after peer wrote into my buffer: {float(t[0].item())} signal ptrs {[hex(p) for p in h.signal_pad_ptrs]}", flush=True)
    ok_symm = True
except Exception as e:
    print(f"[{rank}] symm FAILED: {type(e).__name__}: {str(e)[:300]}", flush=True)
try:
    x = torch.full((1024,), float(10 + rank), device="cuda")
    hd = x.untyped_storage()._share_cuda_()
    objs = [None] * world
    dist.all_gather_object(objs, hd)
    o = objs[(rank + 1) % world]
    st = torch.UntypedStorage._new_shared_cuda(*o)
    y = torch.empty(0, dtype=torch.float32, device=st.device).set_(st, 0, (1024,))
    torch.cuda.synchronize(); dist.barrier()
    print(f"[{rank}] ipc ok: peer value {float(y[0].item())} ptr {hex(y.data_ptr())} dev {y.device}", flush=True)
except Exception as e:
    print(f"[{rank}] ipc FAILED: {type(e).__name__}: {str(e)[:300]}", flush=True)
torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush(); os._exit(0)
