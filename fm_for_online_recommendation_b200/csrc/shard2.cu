// shard2.cu -- row-sharded multi-GPU FM step with O(B*F) work per rank (BASELINE.json configs[4]).
//
// Row r of the packed table lives on rank r % G at local row r / G.  Every rank steps on ITS OWN batch of B samples;
// one call on every rank performs the step over the global batch of G*B samples (loss = mean over G*B, torch.sigmoid
// positions and the bias gradient over the concatenated batch).  The reference has no distributed code; this is the
// B200-native design for its Criteo-scale configuration, built on NVLink peer memory instead of collectives:
//
//   forward   the fused step kernel gathers rows straight from the owners' table shards (cp.async on peer-mapped
//             pointers: 64-byte reads over NVLink / NVSwitch), so logits are computed in the REFERENCE's field order --
//             bit-identical to the single-GPU step on the concatenated batch (the round-1 path folded owner partials)
//   backward  a rank reduces the duplicates of its own batch in sample order (the single-GPU run kernel) and stores one
//             partial gradient per distinct row into the owner's inbox, at the row's position in the rank's stable sort;
//             entries that are the only hit of their row in the rank's batch are stored by the fused kernel itself
//   exchange  epoch flags in peer memory (fmb_shard_signal): no NCCL call in the step
//   owner     scans the G sorted key lists (pushed one step ahead, with the sort), counts the ranks that hit each owned
//             row, adds their partials IN RANK ORDER and applies the row update
// The only change in arithmetic versus one GPU is that order ("rank-partial": oracle/fm_oracle.c rank_B); rows hit by a
// single rank -- all but the hot rows of the small fields and a few per cent of the others -- see no change at all.
// Work per rank per step: B*F entries forward and backward, G*B*F sorted KEYS scanned (10 MB at G = 8), <= B*F inbox
// slots applied -- nothing grows with the global batch except that key scan.
#include "fmb_common.cuh"

extern "C" size_t fmb_bwd_workspace_bytes(int64_t, int);

namespace {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

constexpr int SLOTW = 16;   // floats per inbox slot (64 bytes: k + 1 <= 16 on this path)

struct S2Params {
    float* tables[8];      // peer-mapped table shards
    float* inbox[8];       // peer-mapped inboxes [G sources][N slots][SLOTW]
    float* dl[8];          // peer-mapped [2][G*B]: delta | per-sample loss of the global batch
    const float* bias;
    int B, F, k, rowp, kp4, SB, cu, ql_log, jl_log, loss_kind;
    int G, me;
    float* Gst;            // local component-major staging of multi-hit entries (run kernel input)
    int64_t Npad, N;
};

//   ids [B,F] global row ids of MY batch;  xv [B,F] or NULL;  y [B];  posflag [B*F]: sorted position | multi-hit flag
__global__ void __launch_bounds__(256) shard2_fused_kernel(const int32_t* __restrict__ ids, const float* __restrict__ xv,
                                                           const float* __restrict__ y, const uint32_t* __restrict__ posflag,
                                                           S2Params p) {
    extern __shared__ __align__(16) float smem[];
    const int F = p.F, k = p.k, SB = p.SB, G = p.G;
    const int rp = p.cu * 4;
    float* rows_s = smem;                          // [SB][F][rp]
    float* x_s = rows_s + (size_t)SB * F * rp;     // [SB][F]
    float* bi_s = x_s + SB * F;                    // [SB][k]
    float* S_s = bi_s + SB * k;                    // [SB][kp4]
    float* d_s = S_s + SB * p.kp4;                 // [SB]
    int32_t* ids_s = reinterpret_cast<int32_t*>(d_s + SB);            // [SB][F]
    uint32_t* pos_s = reinterpret_cast<uint32_t*>(ids_s + SB * F);    // [SB][F]
    const int b0 = blockIdx.x * SB;
    const int nv = min(SB, p.B - b0);

    for (int e = threadIdx.x; e < nv * F; e += blockDim.x) {
        ids_s[e] = __ldg(ids + (size_t)b0 * F + e);
        x_s[e] = xv ? __ldg(xv + (size_t)b0 * F + e) : 1.0f;
        pos_s[e] = __ldg(posflag + (size_t)b0 * F + e);
    }
    __syncthreads();
    // gather: row r from its owner's shard (peer-mapped pointer: a 64-byte read over NVLink unless r % G == me)
    {
        const int q = threadIdx.x & ((1 << p.ql_log) - 1);
        const int estep = blockDim.x >> p.ql_log;
        if (q < p.cu)
            for (int ef = threadIdx.x >> p.ql_log; ef < nv * F; ef += estep) {
                const int r = ids_s[ef];
                const int o = r % G;
                cp_async16(rows_s + (size_t)ef * rp + q * 4, p.tables[o] + (size_t)(r / G) * p.rowp + q * 4);
            }
    }
    cp_async_wait_all();
    __syncthreads();
    {
        const int j = threadIdx.x & ((1 << p.jl_log) - 1);
        const int sstep = blockDim.x >> p.jl_log;
        if (j < p.kp4)
            for (int s = threadIdx.x >> p.jl_log; s < nv; s += sstep) {
                float Sj = 0.f;
                if (j < k) {
                    float Qj = 0.f;
                    const float* r = rows_s + (size_t)s * F * rp + j;
                    const float* xs = x_s + s * F;
#pragma unroll 4
                    for (int f = 0; f < F; ++f) {
                        const float e = __fmul_rn(r[(size_t)f * rp], xs[f]);
                        Sj = __fadd_rn(Sj, e);
                        Qj = __fadd_rn(Qj, __fmul_rn(e, e));
                    }
                    bi_s[s * k + j] = __fmul_rn(__fsub_rn(__fmul_rn(Sj, Sj), Qj), 0.5f);
                }
                S_s[s * p.kp4 + j] = Sj;
            }
    }
    __syncthreads();
    {
        const int s = threadIdx.x >> 3, l8 = threadIdx.x & 7;
        const int nv8 = (nv + 3) & ~3;
        if (s < nv8) {
            const int sc = min(s, nv - 1);
            const float* r = rows_s + (size_t)sc * F * rp + k;
            const float* xs = x_s + sc * F;
            const float sf = fmb::aten_row_sum_lanes8([&](int f) { return __fmul_rn(r[(size_t)f * rp], xs[f]); }, F, l8, 0xffffffffu);
            const float* bs = bi_s + sc * k;
            const float sb = fmb::aten_row_sum_lanes8([&](int j) { return bs[j]; }, k, l8, 0xffffffffu);
            if (l8 == 0 && s < nv) {
                const int b = b0 + s;
                const float z = __fadd_rn(__fadd_rn(sf, sb), __ldg(p.bias));
                float lv, d;
                // sample b of rank `me` is element me*B + b of the concatenated batch of G*B samples
                fmb::bce_logits_value_grad(p.loss_kind, z, y[b], p.me * p.B + b, G * p.B, lv, d);
                d_s[s] = d;
                const size_t gb = (size_t)p.me * p.B + b;
                for (int o = 0; o < G; ++o) {          // every rank needs every delta (bias gradient) and loss value
                    p.dl[o][gb] = d;
                    p.dl[o][(size_t)G * p.B + gb] = lv;
                }
            }
        }
    }
    __syncthreads();
    // contributions: a row hit once in MY batch goes straight to its owner's inbox (slot = my rank, sorted position),
    // the others are staged locally for the run kernel
    const int items = nv * F * p.cu;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int ef = it / p.cu, q = it - ef * p.cu;
        const int s = ef / F;
        const float x = x_s[ef], d = d_s[s];
        const uint32_t pf = pos_s[ef];
        const size_t pos = pf & 0x7fffffffu;
        const float* Ss = S_s + s * p.kp4;
        const float4 v4 = *reinterpret_cast<const float4*>(rows_s + (size_t)ef * rp + q * 4);
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
        float a[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int j = q * 4 + t;
            a[t] = 0.f;
            if (j < k) {
                const float ej = __fmul_rn(v[t], x);
                a[t] = __fmul_rn(__fsub_rn(__fmul_rn(d, Ss[j]), __fmul_rn(d, ej)), x);
            } else if (j == k) {
                a[t] = __fmul_rn(d, x);
            }
        }
        if (!(pf >> 31)) {
            const int o = ids_s[ef] % G;   // partial of a single entry = 0 + contribution
            *reinterpret_cast<float4*>(p.inbox[o] + ((size_t)p.me * p.N + pos) * SLOTW + q * 4) =
                make_float4(__fadd_rn(0.f, a[0]), __fadd_rn(0.f, a[1]), __fadd_rn(0.f, a[2]), __fadd_rn(0.f, a[3]));
        } else {
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (q * 4 + t <= k) p.Gst[(size_t)(q * 4 + t) * p.Npad + pos] = a[t];
        }
    }
}

// my sorted keys -> slab `me` of every rank's keys_all [G][N]
struct PeerI32 { int32_t* p[8]; };
__global__ void __launch_bounds__(256) push_keys_kernel(const int32_t* __restrict__ skeys, int64_t N, int G, int me,
                                                        PeerI32 dst) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const int32_t v = skeys[i];
    for (int o = 0; o < G; ++o) dst.p[o][(size_t)me * N + i] = v;
}

struct OwnerParams {
    const int32_t* keys_all;   // [G][N] every rank's sorted keys (per field: B keys, sorted by row id, ties in sample order)
    const float* inbox;        // [G][N][SLOTW] partial gradients at run-start positions
    float* table;              // my shard
    uint32_t* cnt;             // [R_local] ranks that hit each owned row this step (zero between steps)
    int64_t N;
    int B, F, k, rowp, G, me, mode;
    float lr, astep;
};

// is sorted position i of source s the first entry of a run of a row I own?  (key returned through *key)
__device__ __forceinline__ bool owned_run_start(const OwnerParams& p, int s, int64_t i, int32_t* key) {
    const int32_t kk = __ldg(p.keys_all + (size_t)s * p.N + i);
    if (kk % p.G != p.me) return false;
    const int64_t i0 = (i / p.B) * p.B;   // fields are sorted independently: a run never crosses a field boundary
    if (i > i0 && __ldg(p.keys_all + (size_t)s * p.N + i - 1) == kk) return false;
    *key = kk;
    return true;
}

__global__ void __launch_bounds__(256) owner_count_kernel(OwnerParams p) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (int64_t)p.G * p.N) return;
    const int s = (int)(idx / p.N);
    const int64_t i = idx - (int64_t)s * p.N;
    int32_t key;
    if (owned_run_start(p, s, i, &key)) atomicAdd(p.cnt + key / p.G, 1u);   // integer: order-independent
}

// first position of `key` in source s's sorted field segment [lo, lo+B), or -1
__device__ __forceinline__ int64_t find_key(const OwnerParams& p, int s, int64_t lo, int32_t key) {
    const int32_t* a = p.keys_all + (size_t)s * p.N + lo;
    int l = 0, h = p.B;
    while (l < h) { const int m = (l + h) >> 1; if (__ldg(a + m) < key) l = m + 1; else h = m; }
    return (l < p.B && __ldg(a + l) == key) ? lo + l : -1;
}

__global__ void __launch_bounds__(256) owner_apply_kernel(OwnerParams p) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (int64_t)p.G * p.N) return;
    const int s = (int)(idx / p.N);
    const int64_t i = idx - (int64_t)s * p.N;
    int32_t key;
    if (!owned_run_start(p, s, i, &key)) return;
    const int lrow = key / p.G;
    const uint32_t c = p.cnt[lrow];
    const int64_t lo = (i / p.B) * p.B;
    int64_t src_pos[8];
    int nsrc = 1;
    src_pos[0] = (int64_t)s * p.N + i;
    if (c > 1) {
        // several ranks hit this row: the lowest rank among them adds the partials in rank order
        for (int t = 0; t < s; ++t)
            if (find_key(p, t, lo, key) >= 0) return;
        for (int t = s + 1; t < p.G; ++t) {
            const int64_t q = find_key(p, t, lo, key);
            if (q >= 0) src_pos[nsrc++] = (int64_t)t * p.N + q;
        }
    }
    float* row = p.table + (size_t)lrow * p.rowp;
    const int cu = (p.k + 1 + 3) / 4;
    for (int q = 0; q < cu; ++q) {
        float4 g = __ldg(reinterpret_cast<const float4*>(p.inbox + (size_t)src_pos[0] * SLOTW + q * 4));
        for (int t = 1; t < nsrc; ++t) {
            const float4 h = __ldg(reinterpret_cast<const float4*>(p.inbox + (size_t)src_pos[t] * SLOTW + q * 4));
            g.x = __fadd_rn(g.x, h.x); g.y = __fadd_rn(g.y, h.y); g.z = __fadd_rn(g.z, h.z); g.w = __fadd_rn(g.w, h.w);
        }
        const float4 v = *reinterpret_cast<const float4*>(row + q * 4);
        float4 o = v;
        if (q * 4 + 0 <= p.k) o.x = fmb::apply_update_a(v.x, g.x, p.lr, p.astep, p.mode);
        if (q * 4 + 1 <= p.k) o.y = fmb::apply_update_a(v.y, g.y, p.lr, p.astep, p.mode);
        if (q * 4 + 2 <= p.k) o.z = fmb::apply_update_a(v.z, g.z, p.lr, p.astep, p.mode);
        if (q * 4 + 3 <= p.k) o.w = fmb::apply_update_a(v.w, g.w, p.lr, p.astep, p.mode);
        *reinterpret_cast<float4*>(row + q * 4) = o;
    }
    p.cnt[lrow] = 0;   // ready for the next step (only this thread touches the row's counter now)
}

static int ilog2_ceil(int x) { int l = 0; while ((1 << l) < x) ++l; return l; }

}  // namespace

FMB_API int fmb_shard2_slot_floats(void) { return SLOTW; }

// Forward + loss + contributions of MY batch (see file header).  tables/inbox/dl: arrays of G peer-mapped pointers
// (entry `me` = my own buffers).  posflag: fmb_pos_flags of MY ids.  ws: fmb_bwd_workspace_bytes(B*F, k) bytes, handed
// to fmb_shard2_runs afterwards.  dl[o]: [2][G*B] floats (delta | loss) of the step's parity.
FMB_API int fmb_shard2_fused(const int32_t* ids, const float* xv, const float* y, const uint32_t* posflag,
                             void* const* tables, void* const* inbox, void* const* dl, const float* bias, int G, int me,
                             int B, int F, int k, int loss_kind, void* ws, size_t ws_bytes, cudaStream_t stream) {
    FMB_CHECK_ARG(ids && y && posflag && tables && inbox && dl && bias && ws, "fmb_shard2_fused: null pointer");
    FMB_CHECK_ARG(G >= 1 && G <= 8 && me >= 0 && me < G, "fmb_shard2_fused: bad rank %d of %d", me, G);
    FMB_CHECK_ARG(B > 0 && F > 0 && F < 512 && k > 0 && k + 1 <= SLOTW, "fmb_shard2_fused: bad shape B=%d F=%d k=%d (k <= 15)", B, F, k);
    const int64_t N = (int64_t)B * F;
    if (ws_bytes < fmb_bwd_workspace_bytes(N, k)) { fmb_set_error("fmb_shard2_fused: workspace too small"); return FMB_ERR_WS; }
    S2Params p;
    for (int o = 0; o < 8; ++o) {
        p.tables[o] = o < G ? (float*)tables[o] : nullptr;
        p.inbox[o] = o < G ? (float*)inbox[o] : nullptr;
        p.dl[o] = o < G ? (float*)dl[o] : nullptr;
    }
    p.bias = bias; p.B = B; p.F = F; p.k = k; p.rowp = fmb_round_up(k + 1, 16); p.kp4 = fmb_round_up(k, 4);
    p.cu = (k + 1 + 3) / 4; p.ql_log = ilog2_ceil(p.cu); p.jl_log = ilog2_ceil(p.kp4);
    p.loss_kind = loss_kind; p.G = G; p.me = me;
    p.Gst = (float*)ws; p.Npad = (N + 3) / 4 * 4 + 64; p.N = N;
    int SB = 256 >> p.jl_log;
    if (SB < 4) SB = 4;
    if (SB > 32) SB = 32;
    auto bytes = [&](int sb) {
        return sizeof(float) * ((size_t)sb * F * p.cu * 4 + (size_t)3 * sb * F + (size_t)sb * k + (size_t)sb * p.kp4 + sb);
    };
    while (SB > 1 && bytes(SB) > 48 * 1024) SB >>= 1;
    p.SB = SB;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(shard2_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    shard2_fused_kernel<<<(B + SB - 1) / SB, 256, bytes(SB), stream>>>(ids, xv, y, posflag, p);
    FMB_CHECK_LAUNCH("shard2_fused_kernel");
    return FMB_OK;
}

// my sorted keys [N] -> slab `me` of every rank's keys_all (dst: G peer-mapped pointers to [G][N] int32)
FMB_API int fmb_shard2_push_keys(const int32_t* sorted_keys, int64_t N, int G, int me, void* const* dst,
                                 cudaStream_t stream) {
    FMB_CHECK_ARG(sorted_keys && dst && N > 0 && G >= 1 && G <= 8 && me >= 0 && me < G, "fmb_shard2_push_keys: bad arguments");
    PeerI32 d;
    for (int o = 0; o < 8; ++o) d.p[o] = o < G ? (int32_t*)dst[o] : nullptr;
    push_keys_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(sorted_keys, N, G, me, d);
    FMB_CHECK_LAUNCH("push_keys_kernel");
    return FMB_OK;
}

// Owner side: count the ranks that hit each owned row, then add their partials in rank order and update the row.
// keys_all [G][N], inbox [G][N][16] and cnt [R_local] (zero on entry, zero on return) are MY buffers.
FMB_API int fmb_shard2_owner_apply(const int32_t* keys_all, const float* inbox, float* table, uint32_t* cnt, int G, int me,
                                   int B, int F, int k, float lr, int mode, cudaStream_t stream) {
    FMB_CHECK_ARG(keys_all && inbox && table && cnt, "fmb_shard2_owner_apply: null pointer");
    FMB_CHECK_ARG(G >= 1 && G <= 8 && me >= 0 && me < G && B > 0 && F > 0 && k > 0 && k + 1 <= SLOTW, "fmb_shard2_owner_apply: bad arguments");
    FMB_CHECK_ARG(mode == 0 || mode == 1, "fmb_shard2_owner_apply: unknown update mode %d", mode);
    OwnerParams p;
    p.keys_all = keys_all; p.inbox = inbox; p.table = table; p.cnt = cnt; p.N = (int64_t)B * F;
    p.B = B; p.F = F; p.k = k; p.rowp = fmb_round_up(k + 1, 16); p.G = G; p.me = me; p.mode = mode;
    p.lr = lr; p.astep = -(lr / 0.1f);
    const unsigned grid = (unsigned)(((int64_t)G * p.N + 255) / 256);
    owner_count_kernel<<<grid, 256, 0, stream>>>(p);
    FMB_CHECK_LAUNCH("owner_count_kernel");
    owner_apply_kernel<<<grid, 256, 0, stream>>>(p);
    FMB_CHECK_LAUNCH("owner_apply_kernel");
    return FMB_OK;
}
