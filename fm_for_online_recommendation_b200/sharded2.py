"""Row-sharded multi-GPU FM training step, second design: O(B*F) work per rank (csrc/shard2.cu).

One process per GPU (`torch.distributed` for the rendezvous only).  Global row r of the packed table lives on rank
r % G at local row r // G; every rank feeds its own batch of B samples and one call on every rank performs the step
over the global batch of G*B samples.  Rows are gathered straight from the owners' shards over NVLink (peer-mapped
symmetric memory), every rank reduces the duplicates of its own batch and stores one partial gradient per distinct row
into the owner's inbox, the owner adds the partials in rank order and applies the update.  No NCCL call in the step.

`ShardedFM2` is not a reference class (the reference has no distributed code): it is the engine `bench.py --gpus N`
drives.  Summation order: oracle/fm_oracle.c `rank_B` (rank-partial); identical to the reference for G = 1.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, ptr
from .sharded import full_from_shards, local_rows_count, shard_from_full  # noqa: F401  (re-exported helpers)

CH_KEYS, CH_PUSH, CH_ROWS = 0, 1, 2


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _pp(ptrs):
    return (C.c_void_p * len(ptrs))(*[int(p) for p in ptrs])


class ShardedFM2:
    """FM (first + second order, bias) with row-sharded tables; `emulate` = list position of this rank inside a list
    of ShardedFM2 objects living in ONE process (tests: the ranks' phases are called in lock step, no flags)."""

    def __init__(self, feature_sizes, embedding_size, B, n=1e-4, b=0.99, update_mode=0, group=None, seed=0,
                 init="normal", world=None, rank=None, hot_max=4096):
        self._lib = _lib.require_cuda()
        self.group = group
        self.emulated = world is not None
        self.G = world if world is not None else dist.get_world_size(group)
        self.rank = rank if rank is not None else dist.get_rank(group)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.feature_sizes = list(feature_sizes)
        self.F, self.k, self.B = len(feature_sizes), embedding_size, B
        if self.k > 15:
            raise ValueError("the sharded step stores k + 1 <= 16 floats per inbox slot")
        self.rowp = self._lib.fmb_rowp(self.k)
        self.offsets_np = np.concatenate([[0], np.cumsum(feature_sizes)]).astype(np.int64)
        self.R = int(self.offsets_np[-1])
        self.R_local = local_rows_count(self.R, self.G, self.rank)
        self.field_off_dev = torch.from_numpy(self.offsets_np.astype(np.int32)).to(self.device)
        self.lr = float(np.float32(n))
        self.update_mode = update_mode
        G, N = self.G, B * self.F
        self.N = N
        slotw = self._lib.fmb_shard2_slot_floats()
        # "hot" fields (few rows, every batch hits every row): replicated on every rank, the owners keep the replicas
        # current; rows of the other fields reach a requester through its rowbox (pushed by the owners)
        hot_base, nh = [], 0
        for fs in feature_sizes:
            if fs <= hot_max:
                hot_base.append(nh)
                nh += int(fs)
            else:
                hot_base.append(-1)
        self.R_hot = nh
        self.hot_base_dev = torch.tensor(hot_base, dtype=torch.int32, device=self.device)
        # ---- buffers the peers read or write: one arena (symmetric memory when the ranks are processes)
        words = {"table": (self.R_local + 1) * self.rowp, "keys0": G * N, "keys1": G * N, "inbox": G * N * slotw,
                 "dl0": 2 * G * B, "dl1": 2 * G * B, "rowbox": N * slotw, "hot": max(nh, 1) * slotw, "flags": 64}
        # every rank must lay its arena out identically: the table size differs by at most one row between ranks
        words["table"] = (local_rows_count(self.R, G, 0) + 1) * self.rowp
        off, total = {}, 0
        for name, nw in words.items():
            off[name] = total
            total += (nw + 63) // 64 * 64
        if self.emulated:
            arena = torch.zeros(total, dtype=torch.int32, device=self.device)
            self._hdl = None
        else:
            import torch.distributed._symmetric_memory as symm
            grp = group if group is not None else dist.group.WORLD
            arena = symm.empty(total, dtype=torch.int32, device=self.device)
            arena.zero_()
            self._hdl = symm.rendezvous(arena, grp)
            torch.cuda.synchronize()
            dist.barrier(group=grp)
        self._arena, self._off, self._words = arena, off, words
        view = {name: arena[o:o + words[name]] for name, o in off.items()}
        self.table = view["table"].view(torch.float32).view(-1, self.rowp)[:self.R_local + 1]
        self.keys_all = [view["keys0"].view(G, N), view["keys1"].view(G, N)]
        self.inbox = view["inbox"].view(torch.float32)
        self.dl = [view["dl0"].view(torch.float32), view["dl1"].view(torch.float32)]
        self.flags = view["flags"]
        self.rowbox = view["rowbox"].view(torch.float32)
        self.hot = view["hot"].view(torch.float32)
        self._peer_ptrs = None
        if not self.emulated:
            self._bind([int(self._hdl.buffer_ptrs[r]) for r in range(G)])
        if init == "normal":  # N(0,1) like nn.Embedding; drawn on the device (synthetic weights)
            g = torch.Generator(device=self.device)
            g.manual_seed(seed * 1000 + self.rank)
            self.table[:self.R_local, :self.k + 1].normal_(generator=g)
        self.bias = torch.full((1,), float(np.float32(b)), device=self.device)
        if not self.emulated:
            self.sync_hot()
        # ---- private buffers
        self.skeys = [torch.empty(N, dtype=torch.int32, device=self.device) for _ in range(2)]
        self.perm = [torch.empty(N, dtype=torch.int32, device=self.device) for _ in range(2)]
        self.posflag = [torch.empty(N, dtype=torch.int32, device=self.device) for _ in range(2)]
        self.cnt = torch.zeros(self.R_local + 1, dtype=torch.int32, device=self.device)
        self.olist = torch.empty(G * N, dtype=torch.int32, device=self.device)    # owner phase: compact run-start list
        self.nlist = torch.zeros(2, dtype=torch.int32, device=self.device)
        self.ws_bytes = self._lib.fmb_bwd_workspace_bytes(N, self.k)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        self.epoch = torch.zeros(16, dtype=torch.int32, device=self.device)
        self.error = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._pre = torch.cuda.Stream()
        self._side = torch.cuda.Stream()
        self._slot = 0
        self.launches = 0

    def _bind(self, arena_ptrs):
        """peer pointer tables from the base address of every rank's arena"""
        o = self._off
        self._peer_ptrs = {name: _pp([p + 4 * o[name] for p in arena_ptrs])
                           for name in ("table", "keys0", "keys1", "inbox", "dl0", "dl1", "rowbox", "hot", "flags")}

    @staticmethod
    def bind_emulated(ranks):
        ptrs = [m._arena.data_ptr() for m in ranks]
        for m in ranks:
            m._bind(ptrs)

    # ---------------------------------------------------------------- parameters (tests)
    def load_full(self, V, w1, bias):
        t = np.zeros((self.R_local + 1, self.rowp), np.float32)
        t[:self.R_local, :self.k] = shard_from_full(np.asarray(V, np.float32), self.G, self.rank)
        t[:self.R_local, self.k] = shard_from_full(np.asarray(w1, np.float32), self.G, self.rank)
        self.table.copy_(torch.from_numpy(t))
        self.bias.fill_(float(np.asarray(bias).reshape(-1)[0]))

    def sync_hot(self):
        """(re)build every rank's replica of the hot-field rows from the owners' shards (collective: every rank calls it
        after initialising or loading parameters; emulated ranks: call it on every rank, then synchronize)"""
        check(self._lib.fmb_shard2_push_hot(ptr(self.table), self._peer_ptrs["hot"], ptr(self.hot_base_dev),
                                            ptr(self.field_off_dev), self.R_hot, self.G, self.rank, self.F, self.k,
                                            _stream()), "fmb_shard2_push_hot")
        if not self.emulated:
            torch.cuda.synchronize()
            dist.barrier(group=self.group if self.group is not None else dist.group.WORLD)

    def local_params(self):
        t = self.table[:self.R_local].cpu().numpy()
        return t[:, :self.k].copy(), t[:, self.k].copy()

    def encode(self, Xi_local, Y_local):
        a = np.asarray(Xi_local, dtype=np.int64).reshape(-1, self.F)
        ids = torch.from_numpy((a + self.offsets_np[:-1][None, :]).astype(np.int32)).to(self.device)
        y = torch.from_numpy(np.asarray(Y_local, dtype=np.float32).reshape(-1)).to(self.device)
        return ids.contiguous(), y

    # ---------------------------------------------------------------- phases
    def _signal(self, channel, mode):
        if self.emulated:
            return
        check(self._lib.fmb_shard_signal(self._peer_ptrs["flags"], ptr(self.flags), ptr(self.epoch), channel, self.G,
                                         self.rank, mode, ptr(self.error), _stream()), "fmb_shard_signal")
        self.launches += 1

    def _wait(self, channel):
        """(flags, epoch words, channel, error) of an in-kernel wait; nothing to wait for when the ranks are emulated"""
        if self.emulated:
            return None, None, -1, None
        return ptr(self.flags), ptr(self.epoch), channel, ptr(self.error)

    def phase_sort(self, ids, slot, push_after=None):
        """stable sort of MY batch's ids + position words (current stream); then -- once `push_after` (an event: the
        peers have finished with this key slot) has fired -- my sorted keys go to every rank"""
        lib, st = self._lib, _stream()
        check(lib.fmb_sort_fields(ptr(ids), self.B, self.F, ptr(self.field_off_dev), ptr(self.skeys[slot]),
                                  ptr(self.perm[slot]), st), "fmb_sort_fields")
        check(lib.fmb_pos_flags(ptr(self.skeys[slot]), ptr(self.perm[slot]), self.N, ptr(self.posflag[slot]), st),
              "fmb_pos_flags")
        if push_after is not None:
            torch.cuda.current_stream().wait_event(push_after)
        check(lib.fmb_shard2_push_keys(ptr(self.skeys[slot]), self.N, self.G, self.rank,
                                       self._peer_ptrs[f"keys{slot}"], st), "fmb_shard2_push_keys")
        self.launches += 3
        self._signal(CH_KEYS, 3)     # every rank's keys of this batch have landed here

    def phase_rows(self, slot):
        """row service: the rows I own that the other ranks' batches name go to their rowboxes (posted NVLink stores)"""
        check(self._lib.fmb_shard2_push_rows(ptr(self.keys_all[slot]), ptr(self.table), self._peer_ptrs["rowbox"],
                                             self._peer_ptrs["hot"], ptr(self.hot_base_dev), ptr(self.field_off_dev),
                                             self.G, self.rank, self.B, self.F, self.k, _stream()), "fmb_shard2_push_rows")
        self.launches += 1
        self._signal(CH_ROWS, 1)     # published; the forward kernel itself waits for every owner's ROWS epoch

    def phase_forward(self, ids, y, slot, loss_kind=0):
        """gather (all local), logits, loss, contributions: singles -> owners' inboxes, multis -> run kernel -> inboxes"""
        lib, st = self._lib, _stream()
        check(lib.fmb_shard2_fused(ptr(ids), None, ptr(y), ptr(self.posflag[slot]), self._peer_ptrs["table"],
                                   self._peer_ptrs["inbox"], self._peer_ptrs[f"dl{slot}"], ptr(self.rowbox), ptr(self.hot),
                                   ptr(self.hot_base_dev), ptr(self.field_off_dev), ptr(self.bias), self.G,
                                   self.rank, self.B, self.F, self.k, loss_kind, ptr(self.ws), self.ws_bytes,
                                   *self._wait(CH_ROWS), st), "fmb_shard2_fused")
        check(lib.fmb_shard2_runs(ptr(self.skeys[slot]), self.N, self.F, self.k, ptr(self.ws), self.ws_bytes,
                                  self._peer_ptrs["inbox"], self.G, self.rank, st), "fmb_shard2_runs")
        self.launches += 3
        self._signal(CH_PUSH, 1)     # published; the owner phase waits for every rank's PUSH epoch

    def phase_owner(self, slot):
        """owner side: rank-ordered add of the partials + row update; bias step and mean loss over the global batch
        (they need the deltas only: side stream, beside the row updates)"""
        lib = self._lib
        main = torch.cuda.current_stream()
        Bt = self.G * self.B
        loss = torch.empty((), device=self.device)
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            self._signal(CH_PUSH, 2)     # every rank's deltas have landed
            check(lib.fmb_finish_step(ptr(self.dl[slot]), ptr(self.dl[slot][Bt:]), Bt, ptr(self.bias), self.lr,
                                      self.update_mode, ptr(loss), _stream()), "fmb_finish_step")
        check(lib.fmb_shard2_owner_apply(ptr(self.keys_all[slot]), ptr(self.inbox), ptr(self.table), ptr(self.cnt),
                                         ptr(self.olist), ptr(self.nlist), slot, self._peer_ptrs["hot"],
                                         ptr(self.hot_base_dev), ptr(self.field_off_dev), self.G, self.rank, self.B, self.F,
                                         self.k, self.lr, self.update_mode, *self._wait(CH_PUSH), _stream()),
              "fmb_shard2_owner_apply")
        self.launches += 4
        main.wait_stream(self._side)
        return loss

    # ---------------------------------------------------------------- the step
    def prepare(self, ids):
        """sort (and publish the keys of) the first batch"""
        self._slot = 0
        main = torch.cuda.current_stream()
        self._pre.wait_stream(main)
        with torch.cuda.stream(self._pre):
            self.phase_sort(ids, 0)
        main.wait_stream(self._pre)

    def step(self, ids, y, ids_next, loss_kind=0):
        """train on (ids, y) -- the batch given as `ids_next` to the previous call (or to prepare()) -- while the next
        batch's ids are sorted and their keys exchanged on a side stream.  Returns the mean loss over the G*B samples.
        Order on the main stream: row service -> ROWS barrier -> forward + runs -> PUSH barrier -> owner updates."""
        p = self._slot
        main = torch.cuda.current_stream()
        self.phase_rows(p)
        if ids_next is not None:
            # the ROWS barrier means every rank has finished the previous step's owner phase, the last reader of key
            # slot 1 - p: from here on the next batch's keys may land there
            ev = torch.cuda.Event()
            ev.record(main)
            self._pre.wait_stream(main)
            with torch.cuda.stream(self._pre):
                self.phase_sort(ids_next, 1 - p, push_after=ev)
        self.phase_forward(ids, y, p, loss_kind)
        loss = self.phase_owner(p)
        main.wait_stream(self._pre)
        self._slot = 1 - p
        return loss

    def capture(self, ids, y, loss_kind=0):
        """CUDA graphs of the step for both buffer parities over static input buffers (ids/y seed them; two eager
        warm-up steps run first and do train)."""
        self._g_ids = [ids.clone(), ids.clone()]
        self._g_y = y.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self.prepare(self._g_ids[0])
            for _ in range(2):
                self.step(self._g_ids[self._slot], self._g_y, self._g_ids[1 - self._slot], loss_kind)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graphs, self._g_loss = [], []
        for parity in (0, 1):
            self._slot = parity
            g = torch.cuda.CUDAGraph()
            l0 = self.launches
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                loss = self.step(self._g_ids[parity], self._g_y, self._g_ids[1 - parity], loss_kind)
            self._graph_launches = self.launches - l0
            self.launches = l0
            self._graphs.append(g)
            self._g_loss.append(loss)
        self._slot = 0
        return self

    def step_graphed(self, y, ids_next):
        """step(...) through the captured graphs: the current batch's ids already sit in the static buffer of this
        parity (copied there as `ids_next` of the previous call)."""
        p = self._slot
        self._g_y.copy_(y, non_blocking=True)
        self._g_ids[1 - p].copy_(ids_next, non_blocking=True)
        self._graphs[p].replay()
        self.launches += self._graph_launches
        self._slot = 1 - p
        return self._g_loss[p]

    def check_exchange(self):
        v = int(self.error.item())
        if v:
            raise RuntimeError(f"peer-memory exchange: channel {v - 1} timed out waiting for a peer's epoch flag")


# ------------------------------------------------------------------ bench.py --gpus N (N > 1)
def bench_main(args, sizes, config):
    """weak scaling: every rank trains on its own batch of args.batch samples per step."""
    from bench import ClockSampler, measured_peaks, synth_batches  # noqa: WPS433 (bench.py is the caller)
    world, rank = dist.get_world_size(), dist.get_rank()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    B, F, k = args.batch, len(sizes), 10
    K, W = args.steps, max(args.warmup, 3)
    model = ShardedFM2(sizes, k, B, n=1e-4, seed=0)
    NB = 8
    host = synth_batches(sizes, B, NB, 1234 + rank)
    enc = [model.encode(Xi, Y) for Xi, Y in host]
    stream = torch.cuda.current_stream()
    use_graph = os.environ.get("FMB_NO_GRAPH", "0") != "1"
    sampler = ClockSampler(local)
    if use_graph:
        model.capture(*enc[0])
        model.prepare(enc[0][0])
        model._g_ids[0].copy_(enc[0][0])

        def step(i):
            return model.step_graphed(enc[i % NB][1], enc[(i + 1) % NB][0])
    else:
        model.prepare(enc[0][0])

        def step(i):
            return model.step(enc[i % NB][0], enc[i % NB][1], enc[(i + 1) % NB][0])
    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    l0 = model.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(K):
        step(W + i)
    ev1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = model.launches - l0
    # e2e: this rank's labels and the NEXT batch's ids come from pinned host memory every step, the loss goes back
    hosts = [(np.ascontiguousarray((Xi + model.offsets_np[:-1][None, :]).astype(np.int32)), Y) for Xi, Y in host]
    pin_i = torch.empty(B, F, dtype=torch.int32).pin_memory()
    pin_y = torch.empty(B, dtype=torch.float32).pin_memory()
    d_i = torch.empty(B, F, dtype=torch.int32, device="cuda")
    d_y = torch.empty(B, device="cuda")
    cur = [enc[(W + K) % NB][0]]

    def host_step(i):
        pin_i.numpy()[...] = hosts[(i + 1) % NB][0]
        pin_y.numpy()[...] = hosts[i % NB][1]
        d_i.copy_(pin_i, non_blocking=True)
        d_y.copy_(pin_y, non_blocking=True)
        if use_graph:
            out = model.step_graphed(d_y, d_i)
        else:
            out = model.step(cur[0], d_y, d_i)
            cur[0] = d_i.clone()
        return float(out.item())

    for i in range(W):
        host_step(W + K + i)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for i in range(K):
        host_step(2 * W + K + i)
    torch.cuda.synchronize()
    dist.barrier()
    e2e = torch.tensor([time.perf_counter() - t0], device="cuda")
    dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    clocks = sampler.stop()
    model.check_exchange()
    err = torch.tensor([int(model.error.item())], device="cuda")
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        value = world * B * K / (ms * 1e-3)
        step_bytes = B * (8 * F * (k + 1) + 8 * F + 8)
        step_gbps = step_bytes / (ms / K * 1e-3) / 1e9
        peaks, peak_src = measured_peaks()
        line = {
            "metric": "train samples/sec (fwd+bwd+update)", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config, parallelism=f"row-sharded tables over {world} GPUs (r % G): rows gathered from the "
                                               "owners' shards over NVLink peer memory, per-rank partial gradients "
                                               "stored into the owners' inboxes, epoch flags (no NCCL in the step); "
                                               "next batch's sort + key exchange overlapped",
                          global_batch=world * B, exchange_timeouts=int(err.item())),
            "clocks": clocks,
            "e2e": {"value": world * B * K / float(e2e.item()), "unit": "samples/s",
                    "h2d_bytes_per_step": 4 * B * F + 4 * B, "d2h_bytes_per_step": 4,
                    "api": "ShardedFM2.step" + ("_graphed" if use_graph else "") +
                           " (pinned host ids/y in, loss out, per rank)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "whole sharded step, per GPU (see DESIGN.md section 5)",
                         "achieved": step_gbps, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": step_gbps / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": step_bytes,
                         "note": "algorithmic bytes of one GPU's share of the step / step time; 7/8 of the row reads "
                                 "and partial-gradient writes cross NVLink, so the binding roofline is the link "
                                 "(770 GB/s measured per direction), not HBM; the per-kernel HBM roofline is the N=1 run's"},
        }
        print(json.dumps(line), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    # every rank has finished and rank 0 has printed; the symmetric-memory arena is still mapped by the peers, so the
    # process leaves without running destructors in an arbitrary order
    os._exit(0)
