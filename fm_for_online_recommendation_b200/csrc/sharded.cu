// sharded.cu -- kernels of the row-sharded multi-GPU FM step (BASELINE.json configs[4]: 33 M-row
// tables over 2/4/8 B200).
//
// Global row r lives on rank r % G at local row r / G.  Every rank holds its local batch of B
// samples; one step over the global batch of G*B samples is
//   1. all-gather of the transposed ids                      idsT_all [G][F][B]       (NCCL)
//   2. each OWNER sums, for every one of the G*B samples, the rows it owns in field order:
//      partial [G*B][PW] = [S | Q | first]                   shard_partial_forward_kernel
//      and stably sorts, per field, the entries it owns by local row
//                                                            shard_sort_fields_kernel
//   3. all-to-all of the pooled partials (2k+1 floats per sample and peer)            (NCCL)
//   4. each rank folds the G partials of its own samples in owner order, finishes logit, loss and
//      delta: ctx [B][CW] = [S | delta | loss]               shard_combine_kernel
//   5. all-gather of ctx                                                              (NCCL)
//   6. each owner runs the ordinary segmented backward + update over ITS sorted entries
//      (fm_backward.cu, keys >= key_limit are padding), and every rank applies the same bias step.
// The reduction order differs from the 1-GPU path only in step 4 (owner-major instead of
// field-major); oracle/fm_oracle.c restates it (orc_*_sharded) so multi-GPU runs are checked bit for
// bit as well.  No float atomics, no data-dependent host synchronisation.
//
// Peer-memory exchange (the *_peers entry points): with the three exchange buffers in symmetric memory (every
// rank maps every peer's copy), the producers of steps 1, 3 and 5 store their blocks straight into the
// consumers' buffers over NVLink -- the transpose writes its [F][B] slab into all G ranks' idsT_all, the partial
// forward writes block r into rank r's recv, the combine writes its ctx rows into all G ranks' ctx_all -- and a
// one-warp kernel (shard_signal_kernel) publishes a per-channel epoch to the peers' flag words (fence.sys +
// st.release.sys) and/or waits until all G peers have published theirs (ld.acquire.sys).  No NCCL call is left
// in the step; same values in the same places, so the arithmetic and the results are unchanged.
#include "fmb_common.cuh"
#include "smem_sort.cuh"
#include <algorithm>
#include <cstdlib>

namespace {

__device__ __forceinline__ bool owned_by(int32_t gid, int G, int glog, int me, int32_t& local) {
    if (glog >= 0) { local = gid >> glog; return (gid & (G - 1)) == me; }
    local = gid / G;
    return gid - local * G == me;
}

__global__ void transpose_ids_kernel(const int32_t* __restrict__ ids, int B, int F, int32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * F) return;
    const int f = (int)(i / B), b = (int)(i - (int64_t)f * B);
    out[i] = ids[(size_t)b * F + f];
}

constexpr int PF_U = 4;   // fields per round of the partial forward (39 Criteo fields = 3 rounds)

struct PartialParams {
    const int32_t* idsT_all;
    const float* table;
    int G, glog, me, B, F, k, rowp, kp4, cu, ql_log, PW;
    float* partial;
};

__global__ void __launch_bounds__(256) shard_partial_forward_kernel(PartialParams p) {
    const int q = threadIdx.x & ((1 << p.ql_log) - 1);
    const int64_t bg = (int64_t)blockIdx.x * (256 >> p.ql_log) + (threadIdx.x >> p.ql_log);
    if (bg >= (int64_t)p.G * p.B || q >= p.cu) return;
    const int r = (int)(bg / p.B), b = (int)(bg - (int64_t)r * p.B);
    const int32_t* col = p.idsT_all + (size_t)r * p.F * p.B + b;
    float S[4] = {0.f, 0.f, 0.f, 0.f}, Q[4] = {0.f, 0.f, 0.f, 0.f};
    float first = 0.f;
    // PF_U fields per round: all their id loads are in flight together, then the row loads of the owned ones (the
    // kernel is two dependent memory latencies per round and nothing else: 4 fields per round left it latency-bound)
    for (int f0 = 0; f0 < p.F; f0 += PF_U) {
        int32_t lr[PF_U];
        bool own[PF_U];
        float4 v[PF_U];
#pragma unroll
        for (int u = 0; u < PF_U; ++u) {
            own[u] = false;
            if (f0 + u < p.F) own[u] = owned_by(__ldg(col + (size_t)(f0 + u) * p.B), p.G, p.glog, p.me, lr[u]);
        }
#pragma unroll
        for (int u = 0; u < PF_U; ++u)
            if (own[u]) v[u] = *reinterpret_cast<const float4*>(p.table + (size_t)lr[u] * p.rowp + q * 4);
#pragma unroll
        for (int u = 0; u < PF_U; ++u) {
            if (!own[u]) continue;
            const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};  // x == 1 (all-ones feature values)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int j = q * 4 + t;
                if (j < p.k) { S[t] = __fadd_rn(S[t], e[t]); Q[t] = __fadd_rn(Q[t], __fmul_rn(e[t], e[t])); }
                else if (j == p.k) first = __fadd_rn(first, e[t]);
            }
        }
    }
    float* out = p.partial + (size_t)bg * p.PW;
    if (q * 4 < p.kp4) {
        *reinterpret_cast<float4*>(out + q * 4) = make_float4(S[0], S[1], S[2], S[3]);
        *reinterpret_cast<float4*>(out + p.kp4 + q * 4) = make_float4(Q[0], Q[1], Q[2], Q[3]);
    }
    if (q == p.k / 4) out[2 * p.kp4] = first;
}

__global__ void shard_combine_kernel(const float* __restrict__ recv, const float* __restrict__ bias,
                                     const float* __restrict__ y, int G, int me, int B, int k, int kp4, int PW, int CW,
                                     int loss_kind, float* __restrict__ ctx, float* __restrict__ z_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float sum_first = 0.f;
    for (int o = 0; o < G; ++o) sum_first = __fadd_rn(sum_first, recv[((size_t)o * B + b) * PW + 2 * kp4]);
    float* c = ctx + (size_t)b * CW;
    for (int j = 0; j < kp4; ++j) {   // S = sum over owners, owner order (each partial is field-ordered)
        float Sj = 0.f;
        for (int o = 0; o < G; ++o) Sj = __fadd_rn(Sj, recv[((size_t)o * B + b) * PW + j]);
        c[j] = Sj;
    }
    auto bi_of = [&](int j) {
        float Sj = 0.f, Qj = 0.f;
        for (int o = 0; o < G; ++o) {
            const float* r = recv + ((size_t)o * B + b) * PW;
            Sj = __fadd_rn(Sj, r[j]);
            Qj = __fadd_rn(Qj, r[kp4 + j]);
        }
        return __fmul_rn(__fsub_rn(__fmul_rn(Sj, Sj), Qj), 0.5f);
    };
    const float sum_bi = fmb::aten_row_sum_small(bi_of, k);
    const float z = __fadd_rn(__fadd_rn(sum_first, sum_bi), bias[0]);
    if (z_out) z_out[b] = z;
    const float yy = y[b];
    float lossv, d;   // sample b of rank `me` is element me*B + b of the global batch (torch.sigmoid is position-dependent)
    fmb::bce_logits_value_grad(loss_kind, z, yy, me * B + b, G * B, lossv, d);
    c[kp4] = d;
    c[kp4 + 1] = lossv;
    c[kp4 + 2] = z;
    c[kp4 + 3] = 0.f;
}

__global__ void shard_unpack_ctx_kernel(const float* __restrict__ ctx_all, int64_t n, int kp4, int CW,
                                        float* __restrict__ delta, float* __restrict__ lossv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    delta[i] = ctx_all[i * CW + kp4];
    lossv[i] = ctx_all[i * CW + kp4 + 1];
}

// per field: stable compaction of the entries this rank owns (global sample order), then the
// shared-memory LSD sort by local row.  Output segments are padded to `cap` with key = INT_MAX.
__global__ void __launch_bounds__(fmb::SS_THREADS) shard_sort_fields_kernel(
    const int32_t* __restrict__ idsT_all, int G, int glog, int me, int B, int F,
    const int32_t* __restrict__ field_off, int cap, int32_t* __restrict__ skeys, int32_t* __restrict__ perm,
    int32_t* __restrict__ counts, int32_t* __restrict__ overflow, int4* __restrict__ rl_entries,
    uint32_t* __restrict__ rl_segc, int rl_cap, uint32_t* __restrict__ posflag) {
    extern __shared__ __align__(16) unsigned char fs_smem[];
    uint32_t* kbuf0 = reinterpret_cast<uint32_t*>(fs_smem);
    uint32_t* kbuf1 = kbuf0 + cap;
    uint16_t* pbuf0 = reinterpret_cast<uint16_t*>(kbuf1 + cap);
    uint16_t* pbuf1 = pbuf0 + cap + (cap & 1);
    uint16_t* cnt = pbuf1 + cap + (cap & 1);
    __shared__ uint32_t tot[fmb::SS_RADIX];
    __shared__ int wtot[fmb::SS_WARPS];
    __shared__ int s_count;
    const int f = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t off = field_off[f];
    const uint32_t nrows = (uint32_t)(field_off[f + 1] - off);
    int32_t base;
    { int32_t l; owned_by(off, G, glog, me, l); base = l; }           // floor(off / G)
    const uint32_t nloc = nrows / (uint32_t)G + 2;                     // local rows of this field (upper bound)
    const int bits = 32 - __clz(nloc);
    const int passes = (bits + fmb::SS_RADIX_BITS - 1) / fmb::SS_RADIX_BITS;
    // stable compaction of the owned entries, two syncs in total: every warp owns a contiguous slice of the
    // global sample order (r-major, then b), counts its owned entries, the 32 counts are scanned, and the warp
    // then writes its entries behind those of the warps before it.
    const int64_t total = (int64_t)G * B;
    const int64_t per_warp = ((total + fmb::SS_WARPS - 1) / fmb::SS_WARPS + 31) / 32 * 32;
    const int64_t g_lo = min(total, (int64_t)warp * per_warp), g_hi = min(total, g_lo + per_warp);
    const int32_t* fbase = idsT_all + (size_t)f * B;
    const size_t rstride = (size_t)F * B;
    int mycount = 0;
    for (int64_t gb = g_lo; gb < g_hi; gb += 32) {   // warp-uniform trip count
        const int64_t g = gb + lane;
        bool own = false;
        if (g < g_hi) { const int r = (int)(g / B); int32_t l; own = owned_by(__ldg(fbase + r * rstride + (g - (int64_t)r * B)), G, glog, me, l); }
        mycount += __popc(__ballot_sync(0xffffffffu, own));
    }
    if (lane == 0) wtot[warp] = mycount;
    __syncthreads();
    int pre = 0, all = 0;
    for (int w = 0; w < fmb::SS_WARPS; ++w) { const int c = wtot[w]; if (w < warp) pre += c; all += c; }
    if (threadIdx.x == 0) s_count = all;
    for (int64_t gb = g_lo; gb < g_hi; gb += 32) {
        const int64_t g = gb + lane;
        bool own = false;
        int32_t local = 0;
        if (g < g_hi) { const int r = (int)(g / B); own = owned_by(__ldg(fbase + r * rstride + (g - (int64_t)r * B)), G, glog, me, local); }
        const unsigned bal = __ballot_sync(0xffffffffu, own);
        if (own) {
            const int o = pre + __popc(bal & ((1u << lane) - 1u));
            if (o < cap) { kbuf0[o] = (uint32_t)(local - base); pbuf0[o] = (uint16_t)g; }
        }
        pre += __popc(bal);
    }
    __syncthreads();
    int n = s_count;
    if (n > cap) { if (threadIdx.x == 0) atomicMax(overflow, n); n = cap; }
    if (threadIdx.x == 0) counts[f] = n;
    uint32_t* kc; uint16_t* pc;
    // Sparse field under the fused step's contract (posflag given: only the entries of rows hit MORE THAN ONCE need a
    // position, a sorted key and a run; the caller cleared posflag, so rows hit once read 0 = "single"): no sort of the n
    // owned entries -- a shared-memory hash table marks the entries whose row occurs twice, those (a few hundred of 8 192
    // at 1.27 M rows per field) are compacted in sample order and sorted alone.  Same idea as radix_sort.cu's
    // sparse_fields_kernel; the three radix passes over every entry were most of the owner sort's 60-100 us.
    int hslots = 1;
    while (hslots * 2 <= cap) hslots *= 2;
    const bool hashed = posflag != nullptr && n >= 64 && (unsigned long long)nloc >= 8ull * (unsigned)n && 4 * n <= 3 * hslots &&
                        cap <= fmb::SS_WARPS * fmb::SS_RADIX * 2;
    if (hashed) {
        uint32_t* tab = kbuf1;                                   // [hslots] entry index + 1, 0 = empty
        uint8_t* flag = reinterpret_cast<uint8_t*>(cnt);         // [cap] (the radix counters' 16 KB)
        for (int i = threadIdx.x; i < hslots; i += fmb::SS_THREADS) tab[i] = 0u;
        for (int i = threadIdx.x; i < n; i += fmb::SS_THREADS) flag[i] = 0;
        __syncthreads();
        const uint32_t hmask = (uint32_t)hslots - 1u;
        const int hshift = 32 - (31 - __clz(hslots));
        for (int i = threadIdx.x; i < n; i += fmb::SS_THREADS) {
            const uint32_t key = kbuf0[i];
            uint32_t sl = (key * 2654435761u) >> hshift;
            for (;;) {
                const uint32_t old = atomicCAS(&tab[sl], 0u, (uint32_t)i + 1u);
                if (old == 0u) break;
                if (kbuf0[old - 1] == key) { flag[old - 1] = 1; flag[i] = 1; break; }
                sl = (sl + 1u) & hmask;
            }
        }
        __syncthreads();
        // compaction of the marked entries in (global) sample order: thread t owns entries [t*per, (t+1)*per)
        const int per = (n + fmb::SS_THREADS - 1) / fmb::SS_THREADS;
        const int i0 = threadIdx.x * per;
        int mine = 0;
        for (int u = 0; u < per; ++u) if (i0 + u < n) mine += flag[i0 + u];
        int inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        int wpre = 0, nm = 0;
        for (int w = 0; w < fmb::SS_WARPS; ++w) { const int c = wtot[w]; if (w < warp) wpre += c; nm += c; }
        {
            int o = wpre + inc - mine;
            for (int u = 0; u < per; ++u)
                if (i0 + u < n && flag[i0 + u]) { kbuf1[o] = kbuf0[i0 + u]; pbuf1[o] = pbuf0[i0 + u]; ++o; }   // the table is dead: all marks are set
        }
        __syncthreads();
        fmb::smem_sort_passes(kbuf1, kbuf0, pbuf1, pbuf0, cnt, tot, nm, passes, &kc, &pc);
        n = nm;
        if (threadIdx.x == 0) counts[f] = n;
        const int nw = min(cap, n + 192);                        // the run kernel probes keys behind a run's end
        for (int i = threadIdx.x; i < nw; i += fmb::SS_THREADS) {
            skeys[(size_t)f * cap + i] = i < n ? (int32_t)kc[i] + base : 0x7fffffff;
            perm[(size_t)f * cap + i] = i < n ? (int32_t)pc[i] * F + f : 0;
        }
    } else {
        fmb::smem_sort_passes(kbuf0, kbuf1, pbuf0, pbuf1, cnt, tot, n, passes, &kc, &pc);
        for (int i = threadIdx.x; i < cap; i += fmb::SS_THREADS) {
            skeys[(size_t)f * cap + i] = i < n ? (int32_t)kc[i] + base : 0x7fffffff;
            perm[(size_t)f * cap + i] = i < n ? (int32_t)pc[i] * F + f : 0;
        }
    }
    // fused step (shard3.cu): per entry (field, global sample) its sorted position | multi-hit flag, field-major [F][G*B]
    if (posflag) {
        const size_t GB = (size_t)G * B;
        for (int i = threadIdx.x; i < n; i += fmb::SS_THREADS) {
            const uint32_t key = kc[i];
            const bool multi = (i > 0 && kc[i - 1] == key) || (i + 1 < n && kc[i + 1] == key);
            posflag[(size_t)f * GB + pc[i]] = (uint32_t)(f * cap + i) | (multi ? 0x80000000u : 0u);
        }
    }
    // run list for the backward's run kernel (one segment per field; rl_segc = [F short counts | F long counts])
    if (rl_entries)
        fmb::runlist_from_sorted<fmb::SS_THREADS>(kc, n, f * cap, base, rl_entries + (size_t)f * rl_cap, rl_cap, rl_segc + f,
                                                  rl_segc + F + f);
}


struct PeerPtrs { void* p[8]; };

// Exchange synchronisation state.  flags: uint32 [8 channels][8 ranks] in symmetric memory (peers write word
// [channel][their rank] of MY copy); sync_local: uint32 [16] in ordinary device memory = epoch[8] | block counter[8].
struct ExchSync {
    PeerPtrs peer_flags;
    uint32_t* flags_local;
    uint32_t* sync_local;
    int* error;
    int G, me;
};

// Called by EVERY thread of EVERY block at the end of a producer kernel: when the last block gets here, all the
// kernel's peer stores are ordered before the epoch it publishes to the G peers.
__device__ __forceinline__ void publish_epoch_last_block(const ExchSync& x, int channel) {
    __syncthreads();                 // the block's stores are ordered before thread 0's fence (barrier + cumulativity:
    if (threadIdx.x == 0) {          // the grid-barrier idiom of cooperative groups); one fence.sys per block, not per thread
        __threadfence_system();
        const unsigned nb = gridDim.x * gridDim.y * gridDim.z;
        const unsigned prev = atomicAdd(x.sync_local + 8 + channel, 1u);
        if (prev == nb - 1) {
            x.sync_local[8 + channel] = 0;
            __threadfence_system();
            const uint32_t e = x.sync_local[channel] + 1;
            x.sync_local[channel] = e;
            for (int r = 0; r < x.G; ++r) {   // one fence (above), then G relaxed system-scope stores
                uint32_t* f = static_cast<uint32_t*>(x.peer_flags.p[r]) + channel * 8 + x.me;
                asm volatile("st.relaxed.sys.global.u32 [%0], %1;\n" ::"l"(f), "r"(e) : "memory");
            }
        }
    }
}

// Called by every thread of a block at the start of a consumer kernel: returns once all G peers have published
// the epoch this rank's own producer of `channel` (earlier in the stream) has reached.
__device__ __forceinline__ void wait_epoch(const ExchSync& x, int channel) {
    if ((int)threadIdx.x < x.G) {
        const uint32_t e = x.sync_local[channel];
        const uint32_t* f = x.flags_local + channel * 8 + threadIdx.x;
        bool ok = false;
        for (long long spin = 0; spin < (1ll << 26) && !ok; ++spin) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(f) : "memory");
            ok = (int32_t)(v - e) >= 0;
        }
        if (!ok && x.error) *x.error = 1 + channel;
    }
    __syncthreads();
}

// ids [B,F] -> slab `me` of every rank's idsT_all [G][F][B].  A block stages 32 samples x F ids (one contiguous run of the
// sample-major input: coalesced reads) in shared memory and writes, per field, 32 consecutive ids (128 B) to each of the G
// destinations; the strided-read version took 56 us beside the step's kernels at B = 8 192, F = 39.
constexpr int TR_SB = 32;
__global__ void __launch_bounds__(256) transpose_ids_peers_kernel(const int32_t* __restrict__ ids, int B, int F, int G, int me, PeerPtrs dst,
                                                                  ExchSync x, int channel) {
    extern __shared__ int32_t tr_s[];                  // [TR_SB][F + 1]
    const int nb = (B + TR_SB - 1) / TR_SB;
    for (int blk = blockIdx.x; blk < nb; blk += gridDim.x) {
        const int b0 = blk * TR_SB, n = min(TR_SB, B - b0);
        for (int i = threadIdx.x; i < n * F; i += blockDim.x) {
            const int s = i / F, f = i - s * F;
            tr_s[s * (F + 1) + f] = __ldg(ids + (size_t)b0 * F + i);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < F * TR_SB; i += blockDim.x) {
            const int f = i / TR_SB, s = i - f * TR_SB;
            if (s < n) {
                const int32_t v = tr_s[s * (F + 1) + f];
                const size_t o = (size_t)me * F * B + (size_t)f * B + b0 + s;
                for (int r = 0; r < G; ++r) static_cast<int32_t*>(dst.p[r])[o] = v;
            }
        }
        __syncthreads();
    }
    if (channel >= 0) publish_epoch_last_block(x, channel);
}

// shard_partial_forward_kernel with block r of the result stored into rank r's recv [G][B][PW] at block `me`
__global__ void __launch_bounds__(256) shard_partial_forward_peers_kernel(PartialParams p, PeerPtrs dst, ExchSync x,
                                                                         int channel) {
    const int q = threadIdx.x & ((1 << p.ql_log) - 1);
    const int64_t bg = (int64_t)blockIdx.x * (256 >> p.ql_log) + (threadIdx.x >> p.ql_log);
    if (bg < (int64_t)p.G * p.B && q < p.cu) {
    const int r = (int)(bg / p.B), b = (int)(bg - (int64_t)r * p.B);
    const int32_t* col = p.idsT_all + (size_t)r * p.F * p.B + b;
    float S[4] = {0.f, 0.f, 0.f, 0.f}, Q[4] = {0.f, 0.f, 0.f, 0.f};
    float first = 0.f;
    // PF_U fields per round: all their id loads are in flight together, then the row loads of the owned ones (the
    // kernel is two dependent memory latencies per round and nothing else: 4 fields per round left it latency-bound)
    for (int f0 = 0; f0 < p.F; f0 += PF_U) {
        int32_t lr[PF_U];
        bool own[PF_U];
        float4 v[PF_U];
#pragma unroll
        for (int u = 0; u < PF_U; ++u) {
            own[u] = false;
            if (f0 + u < p.F) own[u] = owned_by(__ldg(col + (size_t)(f0 + u) * p.B), p.G, p.glog, p.me, lr[u]);
        }
#pragma unroll
        for (int u = 0; u < PF_U; ++u)
            if (own[u]) v[u] = *reinterpret_cast<const float4*>(p.table + (size_t)lr[u] * p.rowp + q * 4);
#pragma unroll
        for (int u = 0; u < PF_U; ++u) {
            if (!own[u]) continue;
            const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};  // x == 1 (all-ones feature values)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int j = q * 4 + t;
                if (j < p.k) { S[t] = __fadd_rn(S[t], e[t]); Q[t] = __fadd_rn(Q[t], __fmul_rn(e[t], e[t])); }
                else if (j == p.k) first = __fadd_rn(first, e[t]);
            }
        }
    }
    float* out = static_cast<float*>(dst.p[r]) + ((size_t)p.me * p.B + b) * p.PW;
    if (q * 4 < p.kp4) {
        *reinterpret_cast<float4*>(out + q * 4) = make_float4(S[0], S[1], S[2], S[3]);
        *reinterpret_cast<float4*>(out + p.kp4 + q * 4) = make_float4(Q[0], Q[1], Q[2], Q[3]);
    }
    if (q == p.k / 4) out[2 * p.kp4] = first;
    }
    if (channel >= 0) publish_epoch_last_block(x, channel);
}

// shard_combine_kernel fused with both of its exchanges: wait until the G owners' blocks of MY samples have
// landed in recv, stage 32 samples' rows in shared memory with coalesced loads, fold them in owner order (the
// arithmetic of shard_combine_kernel, term for term), store the ctx rows into every rank's ctx_all as 16-byte
// chunks, and publish the ctx epoch from the last block.
constexpr int CMB_SB = 32;        // samples per block
constexpr int CMB_THREADS = 128;
constexpr int CMB_MAXPW = 36;     // k <= 16 on this path (PW = 2*kp4 + 4)
__global__ void __launch_bounds__(CMB_THREADS) shard_combine_peers_kernel(
    const float* recv, const float* __restrict__ bias, const float* __restrict__ y, int G, int me, int B, int k, int kp4,
    int PW, int CW, int loss_kind, PeerPtrs dst_ctx_all, float* ctx_local, ExchSync x, int ch_wait, int ch_publish) {
    __shared__ float rows[8 * CMB_SB * (CMB_MAXPW + 1)];
    __shared__ __align__(16) float ctxs[CMB_SB * 20];
    const int b0 = blockIdx.x * CMB_SB;
    const int nb = min(CMB_SB, B - b0);
    const int pitch = PW + 1;
    if (ch_wait >= 0) wait_epoch(x, ch_wait);
    // recv[o][b0 .. b0+nb) is one contiguous run of nb*PW floats per owner
    const int per_owner4 = nb * PW / 4;
    for (int i = threadIdx.x; i < G * per_owner4; i += CMB_THREADS) {
        const int o = i / per_owner4, j4 = i - o * per_owner4;
        const float4 v = __ldcg(reinterpret_cast<const float4*>(recv + ((size_t)o * B + b0) * PW) + j4);
        const int e = j4 * 4, sb = e / PW, c = e - sb * PW;      // PW % 4 == 0: a chunk never straddles two samples
        float* d = rows + (o * CMB_SB + sb) * pitch + c;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();
    if ((int)threadIdx.x < nb) {
        const int sb = threadIdx.x, b = b0 + sb;
        const float* r0 = rows + sb * pitch;
        const int ostride = CMB_SB * pitch;
        float sum_first = 0.f;
        for (int o = 0; o < G; ++o) sum_first = __fadd_rn(sum_first, r0[o * ostride + 2 * kp4]);
        float* c = ctxs + sb * CW;
        for (int j = 0; j < kp4; ++j) {
            float Sj = 0.f;
            for (int o = 0; o < G; ++o) Sj = __fadd_rn(Sj, r0[o * ostride + j]);
            c[j] = Sj;
        }
        auto bi_of = [&](int j) {
            float Sj = 0.f, Qj = 0.f;
            for (int o = 0; o < G; ++o) {
                Sj = __fadd_rn(Sj, r0[o * ostride + j]);
                Qj = __fadd_rn(Qj, r0[o * ostride + kp4 + j]);
            }
            return __fmul_rn(__fsub_rn(__fmul_rn(Sj, Sj), Qj), 0.5f);
        };
        const float sum_bi = fmb::aten_row_sum_small(bi_of, k);
        const float z = __fadd_rn(__fadd_rn(sum_first, sum_bi), bias[0]);
        const float yy = y[b];
            float lossv, d;   // sample b of rank `me` is element me*B + b of the global batch (torch.sigmoid is position-dependent)
        fmb::bce_logits_value_grad(loss_kind, z, yy, me * B + b, G * B, lossv, d);
        c[kp4] = d;
        c[kp4 + 1] = lossv;
        c[kp4 + 2] = z;
        c[kp4 + 3] = 0.f;
    }
    __syncthreads();
    const int n4 = nb * CW / 4;
    for (int i = threadIdx.x; i < n4; i += CMB_THREADS) {
        const float4 v = reinterpret_cast<const float4*>(ctxs)[i];
        const size_t o4 = ((size_t)me * B + b0) * CW / 4 + i;
        for (int r = 0; r < G; ++r) static_cast<float4*>(dst_ctx_all.p[r])[o4] = v;
        if (ctx_local) reinterpret_cast<float4*>(ctx_local)[(size_t)b0 * CW / 4 + i] = v;
    }
    if (ch_publish >= 0) publish_epoch_last_block(x, ch_publish);
}

// Warp-per-sample partial forward (k <= 31).  A block stages the ids of 64 consecutive samples ([F][64], coalesced
// rows of the transposed ids) in shared memory; a warp then takes one sample at a time: lane f tests the ownership
// of field f (one ballot per 32 fields), the owned fields' rows are fetched with lane j reading component j (all
// rows of a round in flight together), and the sums run over the owned fields in field order -- the same chain of
// fp32 adds as shard_partial_forward_kernel, which spent its time in two dependent latencies per 4 fields.
constexpr int PW_SB = 64;          // samples per block
constexpr int PW_MAXF = 64;        // fields handled (two ballots)
constexpr int PW_INFLIGHT = 8;     // owned rows of one sample fetched together
constexpr int PW_NS = 4;           // samples interleaved per warp
__global__ void __launch_bounds__(256) shard_partial_forward_warp_kernel(PartialParams p, PeerPtrs dst, int to_peers,
                                                                        ExchSync x, int channel) {
    __shared__ int32_t sid[PW_MAXF * (PW_SB + 1)];   // pitch 65: lane f reads row f, bank f + sb
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t total = (int64_t)p.G * p.B;
    const int64_t bg0 = (int64_t)blockIdx.x * PW_SB;     // B % 64 == 0: a block never straddles two ranks' slabs
    const int r = (int)(bg0 / p.B), b0 = (int)(bg0 - (int64_t)r * p.B);
    if (bg0 < total) {
        const int32_t* slab = p.idsT_all + (size_t)r * p.F * p.B + b0;
        for (int i = threadIdx.x; i < p.F * PW_SB; i += 256) {
            const int f = i / PW_SB, sb = i - f * PW_SB;
            sid[f * (PW_SB + 1) + sb] = __ldg(slab + (size_t)f * p.B + sb);
        }
    }
    __syncthreads();
    if (bg0 < total) {
        // PW_NS samples per warp iteration, interleaved, so that their row latencies overlap
        for (int sb0 = warp * PW_NS; sb0 < PW_SB; sb0 += 8 * PW_NS) {
            int32_t l0[PW_NS], l1[PW_NS];
            unsigned m0[PW_NS], m1[PW_NS];
            float S[PW_NS], Q[PW_NS];   // lane j < k: components; lane k: the first-order weight (in S)
#pragma unroll
            for (int a = 0; a < PW_NS; ++a) {
                const int sb = sb0 + a;
                l0[a] = 0; l1[a] = 0;
                const bool o0 = lane < p.F && owned_by(sid[lane * (PW_SB + 1) + sb], p.G, p.glog, p.me, l0[a]);
                const bool o1 = 32 + lane < p.F && owned_by(sid[(32 + lane) * (PW_SB + 1) + sb], p.G, p.glog, p.me, l1[a]);
                m0[a] = __ballot_sync(0xffffffffu, o0);
                m1[a] = __ballot_sync(0xffffffffu, o1);
                S[a] = 0.f; Q[a] = 0.f;
            }
            bool more = true;
            while (more) {   // one round: up to PW_INFLIGHT owned rows of EACH sample in flight, then the ordered sums
                float e[PW_NS][PW_INFLIGHT];
                int cnt[PW_NS];
#pragma unroll
                for (int a = 0; a < PW_NS; ++a) {
                    cnt[a] = 0;
#pragma unroll
                    for (int u = 0; u < PW_INFLIGHT; ++u) {
                        e[a][u] = 0.f;
                        if (m0[a] | m1[a]) {   // warp-uniform
                            int32_t row;
                            if (m0[a]) { const int f = __ffs(m0[a]) - 1; m0[a] &= m0[a] - 1; row = __shfl_sync(0xffffffffu, l0[a], f); }
                            else { const int f = __ffs(m1[a]) - 1; m1[a] &= m1[a] - 1; row = __shfl_sync(0xffffffffu, l1[a], f); }
                            if (lane <= p.k) e[a][u] = p.table[(size_t)row * p.rowp + lane];
                            cnt[a] = u + 1;
                        }
                    }
                }
                more = false;
#pragma unroll
                for (int a = 0; a < PW_NS; ++a) {
#pragma unroll
                    for (int u = 0; u < PW_INFLIGHT; ++u)
                        if (u < cnt[a]) { S[a] = __fadd_rn(S[a], e[a][u]); Q[a] = __fadd_rn(Q[a], __fmul_rn(e[a][u], e[a][u])); }   // x == 1
                    more |= (m0[a] | m1[a]) != 0;
                }
            }
#pragma unroll
            for (int a = 0; a < PW_NS; ++a) {
                const int sb = sb0 + a;
                float* out = to_peers ? static_cast<float*>(dst.p[r]) + ((size_t)p.me * p.B + b0 + sb) * p.PW
                                      : p.partial + (size_t)(bg0 + sb) * p.PW;
                if (lane < p.kp4) {
                    out[lane] = lane < p.k ? S[a] : 0.f;
                    out[p.kp4 + lane] = lane < p.k ? Q[a] : 0.f;
                }
                if (lane == p.k) out[2 * p.kp4] = S[a];
            }
        }
    }
    if (channel >= 0) publish_epoch_last_block(x, channel);
}

// ---------------------------------------------------------------------------------------------------------------
// Tile version of the partial forward (round 2): one CTA = PW_SB samples of one source rank.  The owned entries of the
// tile (sample-major, field order -- the order of the per-sample sums) are compacted once with ballots and a block
// prefix, their rows are gathered into shared memory with 16-byte cp.async copies, all in flight together (the fused
// single-GPU kernel's gather), and 16 lanes per sample sum S, Q and the first-order weight in field order.  The
// warp-per-sample kernel above issues 16.8 M warp instructions for the same 319 k row reads (unrolled 4 x 8 predicated
// load slots per round whatever the sample owned): 35 us.  Bit-identical to it (tests/test_sharded.py with
// FMB_SHARD_TILE_PARTIAL=1), but not yet faster: see sp_use_tile.
// Rounds: the rows of at most `cap` entries are staged at a time (whole samples); with evenly spread ownership one
// round covers the tile, skewed ownership takes more rounds.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sp_cp_async16(void* smem, const void* gmem) {
    unsigned a = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(a), "l"(gmem));
}
__global__ void __launch_bounds__(256) shard_partial_forward_tile_kernel(PartialParams p, PeerPtrs dst, int to_peers,
                                                                        ExchSync x, int channel, int cap) {
    extern __shared__ __align__(16) unsigned char sp_sm[];
    const int F = p.F, rp = p.cu * 4;
    const int npairs = PW_SB * F, nblk = (npairs + 31) / 32;
    int32_t* sid = reinterpret_cast<int32_t*>(sp_sm);                       // [F][PW_SB + 1]
    int32_t* elist = sid + F * (PW_SB + 1);                                 // [PW_SB * F] local rows of the owned entries
    uint32_t* bal = reinterpret_cast<uint32_t*>(elist + PW_SB * F);         // [nblk] ownership ballots of 32 pairs
    uint32_t* boff = bal + nblk;                                            // [nblk + 1] exclusive prefix of their popcounts
    uint32_t* start = boff + nblk + 1;                                      // [PW_SB + 1] first entry of every sample
    float* rows_s = reinterpret_cast<float*>(start + PW_SB + 1 + ((nblk * 2 + PW_SB + 2) & 1 ? 1 : 0));
    rows_s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(rows_s) + 15) & ~(uintptr_t)15);   // [cap][rp]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t total = (int64_t)p.G * p.B;
    const int64_t bg0 = (int64_t)blockIdx.x * PW_SB;     // B % 64 == 0: a block never straddles two ranks' slabs
    const int r = (int)(bg0 / p.B), b0 = (int)(bg0 - (int64_t)r * p.B);
    if (bg0 < total) {
        const int32_t* slab = p.idsT_all + (size_t)r * F * p.B + b0;
        for (int i = threadIdx.x; i < F * PW_SB; i += 256) {
            const int f = i / PW_SB, sb = i - f * PW_SB;
            sid[f * (PW_SB + 1) + sb] = __ldg(slab + (size_t)f * p.B + sb);
        }
        __syncthreads();
        // ownership ballots of the pairs in sample-major order (pair i = sample i / F, field i % F)
        for (int bi = warp; bi < nblk; bi += 8) {
            const int i = bi * 32 + lane;
            bool own = false;
            if (i < npairs) { const int sb = i / F, f = i - sb * F; int32_t l; own = owned_by(sid[f * (PW_SB + 1) + sb], p.G, p.glog, p.me, l); }
            const unsigned m = __ballot_sync(0xffffffffu, own);
            if (lane == 0) bal[bi] = m;
        }
        __syncthreads();
        if (warp == 0) {      // exclusive prefix over the blocks' popcounts
            uint32_t carry = 0;
            for (int c0 = 0; c0 < nblk; c0 += 32) {
                const uint32_t v = c0 + lane < nblk ? __popc(bal[c0 + lane]) : 0u;
                uint32_t inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                if (c0 + lane < nblk) boff[c0 + lane] = carry + inc - v;
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) { boff[nblk] = carry; start[PW_SB] = carry; }
        }
        __syncthreads();
        for (int bi = warp; bi < nblk; bi += 8) {
            const int i = bi * 32 + lane;
            if (i < npairs) {
                const int sb = i / F, f = i - sb * F;
                const unsigned m = bal[bi];
                const uint32_t pos = boff[bi] + __popc(m & ((1u << lane) - 1u));
                if (f == 0) start[sb] = pos;
                if ((m >> lane) & 1u) { int32_t l; owned_by(sid[f * (PW_SB + 1) + sb], p.G, p.glog, p.me, l); elist[pos] = l; }
            }
        }
        __syncthreads();
        // rounds of whole samples whose entries fit the staging area
        int s_lo = 0;
        while (s_lo < PW_SB) {
            const uint32_t e_lo = start[s_lo];
            int s_hi = s_lo + 1;
            while (s_hi < PW_SB && start[s_hi + 1] - e_lo <= (uint32_t)cap) ++s_hi;
            const uint32_t e_hi = start[s_hi];
            {   // gather: 16 B per cp.async, every row of the round in flight at once
                const int q = threadIdx.x & ((1 << p.ql_log) - 1);
                const int estep = 256 >> p.ql_log;
                if (q < p.cu)
                    for (uint32_t e = e_lo + (threadIdx.x >> p.ql_log); e < e_hi; e += estep)
                        sp_cp_async16(rows_s + (size_t)(e - e_lo) * rp + q * 4, p.table + (size_t)elist[e] * p.rowp + q * 4);
            }
            asm volatile("cp.async.commit_group;\n" ::);
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            __syncthreads();
            // 16 lanes per sample: lane j < k sums component j (and its square), lane k the first-order weight; field order
            for (int sb = s_lo + (threadIdx.x >> 4); sb < s_hi; sb += 16) {
                const int j = threadIdx.x & 15;
                if (j <= p.k) {
                    float S = 0.f, Q = 0.f;
                    const float* rr = rows_s + (size_t)(start[sb] - e_lo) * rp + j;
                    const int n = (int)(start[sb + 1] - start[sb]);
                    for (int u = 0; u < n; ++u) { const float e = rr[(size_t)u * rp]; S = __fadd_rn(S, e); Q = __fadd_rn(Q, __fmul_rn(e, e)); }   // x == 1
                    float* out = to_peers ? static_cast<float*>(dst.p[r]) + ((size_t)p.me * p.B + b0 + sb) * p.PW
                                          : p.partial + (size_t)(bg0 + sb) * p.PW;
                    if (j < p.k) { out[j] = S; out[p.kp4 + j] = Q; }
                    else {
                        out[2 * p.kp4] = S;
                        if (j < p.kp4) { out[j] = 0.f; out[p.kp4 + j] = 0.f; }      // padding component
                    }
                } else if (j < p.kp4) {     // padding components
                    float* out = to_peers ? static_cast<float*>(dst.p[r]) + ((size_t)p.me * p.B + b0 + sb) * p.PW
                                          : p.partial + (size_t)(bg0 + sb) * p.PW;
                    out[j] = 0.f; out[p.kp4 + j] = 0.f;
                }
            }
            __syncthreads();
            s_lo = s_hi;
        }
    }
    if (channel >= 0) publish_epoch_last_block(x, channel);
}
static size_t sp_tile_smem(int F, int cu, int cap) {
    const int npairs = PW_SB * F, nblk = (npairs + 31) / 32;
    return (size_t)F * (PW_SB + 1) * 4 + (size_t)npairs * 4 + (size_t)(2 * nblk + 1 + PW_SB + 1 + 1) * 4 + 16 + (size_t)cap * cu * 16;
}
// staging capacity (entries): 1.5 x the tile's expected owned entries, at least one sample's worth, at most 1 024
static int sp_tile_cap(int G, int F) {
    int c = (PW_SB * F * 3 / 2 / G + 63) / 64 * 64;
    if (c < 128) c = 128;
    if (c > 1024) c = 1024;
    return c;
}
static bool sp_use_tile(int k, int F) {
    if (k > 15 || F > PW_MAXF) return false;
    // Experiment, off by default (FMB_SHARD_TILE_PARTIAL=1): measured 51 us against the warp kernel's 38 at G = 8 and
    // 31 against 33 at G = 2 (cold L2) -- 1 024 CTAs of 36 KB are 1.15 waves, and the tile's five barriers with two
    // runtime divisions per pair and a serial search for the round's end cost more than the instructions saved.
    static int forced = -2;
    if (forced == -2) { const char* e = getenv("FMB_SHARD_TILE_PARTIAL"); forced = e ? (e[0] != '0') : 0; }
    return forced != 0;
}

// mode bit 0: publish my next epoch of `channel` to every peer's flag word [channel][me];
// mode bit 1: wait until every peer's epoch of `channel` has reached mine.
__global__ void shard_signal_kernel(PeerPtrs peer_flags, uint32_t* flags_local, uint32_t* epoch_local, int channel, int G,
                                    int me, int mode, int* error) {
    const int lane = threadIdx.x;
    uint32_t e = epoch_local[channel];
    if (mode & 1) {
        e += 1;
        __threadfence_system();   // the producer kernels before this one in the stream: their peer stores first
        if (lane < G) {
            uint32_t* f = static_cast<uint32_t*>(peer_flags.p[lane]) + channel * 8 + me;
            asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(f), "r"(e) : "memory");
        }
        __syncwarp();
        if (lane == 0) epoch_local[channel] = e;
    }
    if (mode & 2) {
        if (lane < G) {
            const uint32_t* f = flags_local + channel * 8 + lane;
            bool ok = false;
            for (long long spin = 0; spin < (1ll << 26) && !ok; ++spin) {
                uint32_t v;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(f) : "memory");
                ok = (int32_t)(v - e) >= 0;
            }
            if (!ok && error) *error = 1 + channel;
        }
        __syncwarp();
        __threadfence_system();
    }
}

// warp-per-sample partial forward instead of the lane-per-chunk kernel (measured: 38 vs 55 us at
// G = 8, 33 vs 35 us at G = 2, cold L2, with four samples interleaved per warp); FMB_SHARD_WARP_PARTIAL=0/1 forces it
static bool fmb_shard_use_warp_partial(int G, int B, int F, int k) {
    if (!(k <= 31 && F <= PW_MAXF && B % PW_SB == 0)) return false;
    static int forced = -2;
    if (forced == -2) { const char* e = getenv("FMB_SHARD_WARP_PARTIAL"); forced = e ? (e[0] != '0') : -1; }
    return forced >= 0 ? forced != 0 : G >= 2;
}

static int ilog2_exact(int x) { int l = 0; while ((1 << l) < x) ++l; return (1 << l) == x ? l : -1; }
static int ilog2_ceil(int x) { int l = 0; while ((1 << l) < x) ++l; return l; }

}  // namespace

FMB_API int fmb_shard_pw(int k) { return 2 * fmb_round_up(k, 4) + 4; }   // floats per pooled partial
FMB_API int fmb_shard_cw(int k) { return fmb_round_up(k, 4) + 4; }       // floats per sample context
FMB_API int fmb_shard_sort_max_cap(void) { return fmb::SS_WARPS * fmb::SS_MAX_SLOTS * 32; }

// ids [B,F] -> out [F,B]
FMB_API int fmb_transpose_ids(const int32_t* ids, int B, int F, int32_t* out, cudaStream_t stream) {
    FMB_CHECK_ARG(ids && out && B > 0 && F > 0, "fmb_transpose_ids: bad arguments");
    const int64_t n = (int64_t)B * F;
    transpose_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(ids, B, F, out);
    FMB_CHECK_LAUNCH("transpose_ids_kernel");
    return FMB_OK;
}

// step 2a: pooled partials of the rows rank `me` owns, for all G*B samples (feature values all ones)
FMB_API int fmb_shard_partial_forward(const int32_t* idsT_all, const float* table_local, int G, int me, int B, int F,
                                      int k, float* partial, cudaStream_t stream) {
    FMB_CHECK_ARG(idsT_all && table_local && partial, "fmb_shard_partial_forward: null pointer");
    FMB_CHECK_ARG(G > 0 && me >= 0 && me < G && B > 0 && F > 0 && k > 0 && k <= 124, "fmb_shard_partial_forward: bad arguments");
    PartialParams p;
    p.idsT_all = idsT_all; p.table = table_local; p.G = G; p.glog = ilog2_exact(G); p.me = me; p.B = B; p.F = F;
    p.k = k; p.rowp = fmb_round_up(k + 1, 16); p.kp4 = fmb_round_up(k, 4); p.cu = (k + 1 + 3) / 4;
    p.ql_log = ilog2_ceil(p.cu); p.PW = fmb_shard_pw(k); p.partial = partial;
    const int spb = 256 >> p.ql_log;
    const int64_t n = (int64_t)G * B;
    if (fmb_shard_use_warp_partial(G, B, F, k)) {   // one CTA per 64 samples of one source rank
        PeerPtrs none = {};
        ExchSync nox = {};
        if (sp_use_tile(k, F)) {
            const int cap = sp_tile_cap(G, F);
            static bool attr = false;
            if (!attr) { cudaFuncSetAttribute(shard_partial_forward_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); attr = true; }
            shard_partial_forward_tile_kernel<<<(unsigned)(n / PW_SB), 256, sp_tile_smem(F, p.cu, cap), stream>>>(p, none, 0, nox, -1, cap);
        } else {
            shard_partial_forward_warp_kernel<<<(unsigned)(n / PW_SB), 256, 0, stream>>>(p, none, 0, nox, -1);
        }
    } else {
        shard_partial_forward_kernel<<<(unsigned)((n + spb - 1) / spb), 256, 0, stream>>>(p);
    }
    FMB_CHECK_LAUNCH("shard_partial_forward_kernel");
    return FMB_OK;
}

// step 4: recv [G][B][PW] (block o = partials owner o computed for MY samples) -> ctx [B][CW]
FMB_API int fmb_shard_combine(const float* recv, const float* bias, const float* y, int G, int me, int B, int k,
                              int loss_kind, float* ctx, float* z_out, cudaStream_t stream) {
    FMB_CHECK_ARG(recv && bias && y && ctx && G > 0 && me >= 0 && me < G && B > 0 && k > 0, "fmb_shard_combine: bad arguments");
    shard_combine_kernel<<<(B + 127) / 128, 128, 0, stream>>>(recv, bias, y, G, me, B, k, fmb_round_up(k, 4),
                                                            fmb_shard_pw(k), fmb_shard_cw(k), loss_kind, ctx, z_out);
    FMB_CHECK_LAUNCH("shard_combine_kernel");
    return FMB_OK;
}

FMB_API int fmb_shard_unpack_ctx(const float* ctx_all, int64_t n, int k, float* delta, float* lossv,
                                 cudaStream_t stream) {
    FMB_CHECK_ARG(ctx_all && delta && lossv && n > 0, "fmb_shard_unpack_ctx: bad arguments");
    shard_unpack_ctx_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(ctx_all, n, fmb_round_up(k, 4),
                                                                           fmb_shard_cw(k), delta, lossv);
    FMB_CHECK_LAUNCH("shard_unpack_ctx_kernel");
    return FMB_OK;
}

// step 2b: per-field sorted list of the entries rank `me` owns.  skeys/perm [F][cap] (padding key
// INT_MAX), counts [F], overflow [1] (max count seen when a field exceeded cap, else untouched).
// _rl: also the run list of every field's sorted owned entries (rl: nseg == F, seg_cap >= cap / 2 + 1; nullable), for
// fmb_fm_backward_update_rl
struct fmb_runlist_t { int32_t* entries; uint32_t* seg_count; int nseg, seg_cap; };   // include/fmb200.h
// _pf: also posflag [F][G*B] (nullable): sorted position | 0x80000000 when the row is hit more than once -- what
// fmb_shard3_step reads.  The buffer is cleared here first (memset on `stream`); fields with many more rows than owned entries
// then write only the words of their multi-hit entries and list only those in skeys / perm (hash pass instead of the sort)
FMB_API int fmb_shard_sort_fields_pf(const int32_t* idsT_all, int G, int me, int B, int F, const int32_t* field_off,
                                     int cap, int32_t* skeys, int32_t* perm, int32_t* counts, int32_t* overflow,
                                     const fmb_runlist_t* rl, uint32_t* posflag, cudaStream_t stream) {
    FMB_CHECK_ARG(idsT_all && field_off && skeys && perm && counts && overflow, "fmb_shard_sort_fields: null pointer");
    FMB_CHECK_ARG(!rl || (rl->entries && rl->seg_count && rl->nseg == F && rl->seg_cap >= cap / 2 + 1 && F <= 2048),
                  "fmb_shard_sort_fields: run list must have F segments of at least cap/2 + 1 entries");
    FMB_CHECK_ARG(cap > 0 && cap <= fmb_shard_sort_max_cap(), "fmb_shard_sort_fields: cap=%d out of range", cap);
    FMB_CHECK_ARG((int64_t)G * B <= 65536, "fmb_shard_sort_fields: G*B must be <= 65536 (16-bit sample payload)");
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(shard_sort_fields_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); attr = true; }
    if (posflag) {
        const cudaError_t me_ = cudaMemsetAsync(posflag, 0, (size_t)F * G * B * sizeof(uint32_t), stream);
        if (me_ != cudaSuccess) { fmb_set_error("fmb_shard_sort_fields: %s", cudaGetErrorString(me_)); return FMB_ERR_CUDA; }
    }
    shard_sort_fields_kernel<<<F, fmb::SS_THREADS, fmb::smem_sort_bytes(cap), stream>>>(
        idsT_all, G, ilog2_exact(G), me, B, F, field_off, cap, skeys, perm, counts, overflow,
        rl ? reinterpret_cast<int4*>(rl->entries) : nullptr, rl ? rl->seg_count : nullptr, rl ? rl->seg_cap : 0, posflag);
    FMB_CHECK_LAUNCH("shard_sort_fields_kernel");
    return FMB_OK;
}

FMB_API int fmb_shard_sort_fields_rl(const int32_t* idsT_all, int G, int me, int B, int F, const int32_t* field_off,
                                     int cap, int32_t* skeys, int32_t* perm, int32_t* counts, int32_t* overflow,
                                     const fmb_runlist_t* rl, cudaStream_t stream) {
    return fmb_shard_sort_fields_pf(idsT_all, G, me, B, F, field_off, cap, skeys, perm, counts, overflow, rl, nullptr, stream);
}

FMB_API int fmb_shard_sort_fields(const int32_t* idsT_all, int G, int me, int B, int F, const int32_t* field_off,
                                  int cap, int32_t* skeys, int32_t* perm, int32_t* counts, int32_t* overflow,
                                  cudaStream_t stream) {
    return fmb_shard_sort_fields_rl(idsT_all, G, me, B, F, field_off, cap, skeys, perm, counts, overflow, nullptr, stream);
}

// ---------------------------------------------------------------------------------------------- peer-memory exchange
// `dst`/`peer_flags`: HOST arrays of G device pointers, entry r = the address at which THIS device maps rank r's
// copy of the buffer (torch symmetric memory `buffer_ptrs`, or CUDA IPC mappings).  G <= 8.
static int fill_peers(PeerPtrs& pp, void* const* ptrs, int G, const char* who) {
    if (!ptrs || G < 1 || G > 8) { fmb_set_error("%s: needs 1..8 peer pointers", who); return FMB_ERR_ARG; }
    for (int r = 0; r < 8; ++r) pp.p[r] = r < G ? ptrs[r] : nullptr;
    for (int r = 0; r < G; ++r)
        if (!pp.p[r]) { fmb_set_error("%s: peer pointer %d is null", who, r); return FMB_ERR_ARG; }
    return FMB_OK;
}

static int fill_sync(ExchSync& x, void* const* flag_peers, uint32_t* flags_local, uint32_t* sync_local, int* error_dev,
                     int G, int me, bool needed, const char* who) {
    x.flags_local = flags_local; x.sync_local = sync_local; x.error = error_dev; x.G = G; x.me = me;
    for (int r = 0; r < 8; ++r) x.peer_flags.p[r] = nullptr;
    if (!needed) return FMB_OK;
    if (!flags_local || !sync_local) { fmb_set_error("%s: flag block / sync block missing", who); return FMB_ERR_ARG; }
    return fill_peers(x.peer_flags, flag_peers, G, who);
}

// step 1 without a collective: ids [B,F] -> slab `me` of every rank's idsT_all [G][F][B]; publish_channel >= 0:
// the last block publishes that channel's next epoch to the peers (see fmb_shard_signal), -1: no flag.
FMB_API int fmb_shard_transpose_ids_peers(const int32_t* ids, int B, int F, int G, int me, void* const* dst_idsT_all,
                                          void* const* flag_peers, uint32_t* flags_local, uint32_t* sync_local,
                                          int* error_dev, int publish_channel, cudaStream_t stream) {
    FMB_CHECK_ARG(ids && B > 0 && F > 0 && me >= 0 && me < G && publish_channel < 8, "fmb_shard_transpose_ids_peers: bad arguments");
    PeerPtrs pp;
    if (int rc = fill_peers(pp, dst_idsT_all, G, "fmb_shard_transpose_ids_peers")) return rc;
    ExchSync x;
    if (int rc = fill_sync(x, flag_peers, flags_local, sync_local, error_dev, G, me, publish_channel >= 0, "fmb_shard_transpose_ids_peers")) return rc;
    const int64_t n = (int64_t)B * F;
    const unsigned blocks = (unsigned)std::min<int64_t>((B + TR_SB - 1) / TR_SB, 296);   // grid-stride: few blocks, few fences
    (void)n;
    transpose_ids_peers_kernel<<<blocks, 256, (size_t)TR_SB * (F + 1) * 4, stream>>>(ids, B, F, G, me, pp, x, publish_channel);
    FMB_CHECK_LAUNCH("transpose_ids_peers_kernel");
    return FMB_OK;
}

// steps 2a + 3 without a collective: block r of the pooled partials goes straight into rank r's recv [G][B][PW]
FMB_API int fmb_shard_partial_forward_peers(const int32_t* idsT_all, const float* table_local, int G, int me, int B,
                                            int F, int k, void* const* dst_recv, void* const* flag_peers,
                                            uint32_t* flags_local, uint32_t* sync_local, int* error_dev,
                                            int publish_channel, cudaStream_t stream) {
    FMB_CHECK_ARG(idsT_all && table_local, "fmb_shard_partial_forward_peers: null pointer");
    FMB_CHECK_ARG(G > 0 && me >= 0 && me < G && B > 0 && F > 0 && k > 0 && k <= 124 && publish_channel < 8,
                  "fmb_shard_partial_forward_peers: bad arguments");
    PeerPtrs pp;
    if (int rc = fill_peers(pp, dst_recv, G, "fmb_shard_partial_forward_peers")) return rc;
    ExchSync x;
    if (int rc = fill_sync(x, flag_peers, flags_local, sync_local, error_dev, G, me, publish_channel >= 0, "fmb_shard_partial_forward_peers")) return rc;
    PartialParams p;
    p.idsT_all = idsT_all; p.table = table_local; p.G = G; p.glog = ilog2_exact(G); p.me = me; p.B = B; p.F = F;
    p.k = k; p.rowp = fmb_round_up(k + 1, 16); p.kp4 = fmb_round_up(k, 4); p.cu = (k + 1 + 3) / 4;
    p.ql_log = ilog2_ceil(p.cu); p.PW = fmb_shard_pw(k); p.partial = nullptr;
    const int spb = 256 >> p.ql_log;
    const int64_t n = (int64_t)G * B;
    if (fmb_shard_use_warp_partial(G, B, F, k) && sp_use_tile(k, F)) {
        const int cap = sp_tile_cap(G, F);
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(shard_partial_forward_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); attr = true; }
        shard_partial_forward_tile_kernel<<<(unsigned)(n / PW_SB), 256, sp_tile_smem(F, p.cu, cap), stream>>>(p, pp, 1, x, publish_channel, cap);
    } else if (fmb_shard_use_warp_partial(G, B, F, k))
        shard_partial_forward_warp_kernel<<<(unsigned)(n / PW_SB), 256, 0, stream>>>(p, pp, 1, x, publish_channel);
    else
        shard_partial_forward_peers_kernel<<<(unsigned)((n + spb - 1) / spb), 256, 0, stream>>>(p, pp, x, publish_channel);
    FMB_CHECK_LAUNCH("shard_partial_forward_peers_kernel");
    return FMB_OK;
}

// steps 4 + 5 without a collective: wait_channel >= 0: first wait until all G owners have published that channel
// (their blocks are in recv); fold; store the ctx rows into every rank's ctx_all (and ctx_local when given);
// publish_channel >= 0: publish it from the last block.  k <= 16.
FMB_API int fmb_shard_combine_peers(const float* recv, const float* bias, const float* y, int G, int me, int B, int k,
                                    int loss_kind, void* const* dst_ctx_all, float* ctx_local, void* const* flag_peers,
                                    uint32_t* flags_local, uint32_t* sync_local, int* error_dev, int wait_channel,
                                    int publish_channel, cudaStream_t stream) {
    FMB_CHECK_ARG(recv && bias && y && G > 0 && me >= 0 && me < G && B > 0 && k > 0, "fmb_shard_combine_peers: bad arguments");
    FMB_CHECK_ARG(fmb_shard_pw(k) <= CMB_MAXPW && wait_channel < 8 && publish_channel < 8, "fmb_shard_combine_peers: k=%d too large for this path (k <= 16)", k);
    PeerPtrs pp;
    if (int rc = fill_peers(pp, dst_ctx_all, G, "fmb_shard_combine_peers")) return rc;
    ExchSync x;
    if (int rc = fill_sync(x, flag_peers, flags_local, sync_local, error_dev, G, me, wait_channel >= 0 || publish_channel >= 0, "fmb_shard_combine_peers")) return rc;
    shard_combine_peers_kernel<<<(B + CMB_SB - 1) / CMB_SB, CMB_THREADS, 0, stream>>>(
        recv, bias, y, G, me, B, k, fmb_round_up(k, 4), fmb_shard_pw(k), fmb_shard_cw(k), loss_kind, pp, ctx_local, x,
        wait_channel, publish_channel);
    FMB_CHECK_LAUNCH("shard_combine_peers_kernel");
    return FMB_OK;
}

// Epoch flags of the exchange: flag words are uint32 [8 channels][8 ranks] in symmetric memory, `epoch_local` =
// the sync block, uint32 [16] in ordinary device memory (epoch[8] | block counters[8], zero-initialised).  mode 1 = publish (after the producer kernel, same stream), 2 = wait for
// all G peers (before the consumer kernel), 3 = both.  error_dev (nullable) receives 1 + channel on a time-out.
FMB_API int fmb_shard_signal(void* const* peer_flags, uint32_t* flags_local, uint32_t* epoch_local, int channel, int G,
                             int me, int mode, int* error_dev, cudaStream_t stream) {
    FMB_CHECK_ARG(flags_local && epoch_local && channel >= 0 && channel < 8 && me >= 0 && me < G && mode >= 1 && mode <= 3,
                  "fmb_shard_signal: bad arguments");
    PeerPtrs pp;
    if (int rc = fill_peers(pp, peer_flags, G, "fmb_shard_signal")) return rc;
    shard_signal_kernel<<<1, 32, 0, stream>>>(pp, flags_local, epoch_local, channel, G, me, mode, error_dev);
    FMB_CHECK_LAUNCH("shard_signal_kernel");
    return FMB_OK;
}
