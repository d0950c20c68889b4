// fm_backward.cu -- sparse embedding gradient as a deterministic segmented reduce over the sorted
// (row id, entry) list, fused with the per-row parameter update.
//
// Replaces autograd + 2F dense `embedding_dense_backward` + the dense `Adam.step` over all rows
// (models/models_online_deep/fm_adam.py:56-69, deepfm_adam.py:91-117): only
// the rows a batch touches are read and written.  Per entry (b, f) of row r:
//     e_j   = V_r[j] * x                                     (deepfm_adam.py:60)
//     g_e_j = (g*S_b[j]) - (g*e_j)      g = delta_b (FM scalar path) and/or gvec_b[j] (MLP path)
//     grad V_r[j] += g_e_j * x ;  grad w_r += delta_b * x    summed IN SAMPLE ORDER per row
// then one update per row (fresh-Adam sign step or SGD).  No float atomics anywhere: each run of
// equal row ids is owned by exactly one CTA, which accumulates it left to right.
//
// Two kernels: fm_bwd_tile_kernel owns every run that starts inside its tile of TE sorted entries
// and ends inside a 2*TE window; runs that leave the window (hot rows of tiny fields) are pushed to
// a list and finished by fm_bwd_long_kernel, one CTA per run.
#include "fmb_common.cuh"

namespace {

struct BwdParams {
    const int32_t* skeys;
    const int32_t* perm;
    int64_t N;
    const float* xv;
    float* table;
    int F, k, rowp, kp4;
    const float* S;
    const float* gs;
    int use_fm2;
    const float* gvec;
    float lr;
    int mode;
    int32_t* long_list;
    int32_t* long_count;
    int TE;
};

// store the contributions of one (entry, chunk) item into the staging buffers
__device__ __forceinline__ void store_contrib(const BwdParams& p, bool two, float* bufA, float* bufB, int i, int q,
                                              float4 a, float4 c) {
    const int rowp = p.rowp;
    if (two) {
        *reinterpret_cast<float4*>(bufA + (size_t)i * rowp + q * 4) = a;
        *reinterpret_cast<float4*>(bufB + (size_t)i * rowp + q * 4) = c;
    } else if (p.gvec) {  // NFM: second-order comps from the MLP path, first-order comp from delta
        const int kq = p.k >> 2, kt = p.k & 3;
        float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
        if (q == kq) {
#pragma unroll
            for (int t = 0; t < 4; ++t) if (t >= kt) cv[t] = av[t];
        }
        *reinterpret_cast<float4*>(bufA + (size_t)i * rowp + q * 4) =
            (q > kq) ? a : make_float4(cv[0], cv[1], cv[2], cv[3]);
    } else {
        *reinterpret_cast<float4*>(bufA + (size_t)i * rowp + q * 4) = a;
    }
}

// everything one (entry, chunk) item needs from global memory, loaded up front so that a batch of
// items has all of its loads in flight together
struct ItemLoads {
    float4 v4, s4, g4;
    float x, d;
};
__device__ __forceinline__ void item_load(const BwdParams& p, int32_t key, int32_t e, int q, ItemLoads& L) {
    const int b = e / p.F;
    L.x = p.xv ? __ldg(p.xv + e) : 1.0f;
    L.d = __ldg(p.gs + b);
    L.v4 = *reinterpret_cast<const float4*>(p.table + (size_t)key * p.rowp + q * 4);
    L.s4 = make_float4(0.f, 0.f, 0.f, 0.f);
    L.g4 = L.s4;
    if (q * 4 < p.kp4) {
        L.s4 = __ldg(reinterpret_cast<const float4*>(p.S + (size_t)b * p.kp4 + q * 4));
        if (p.gvec) L.g4 = __ldg(reinterpret_cast<const float4*>(p.gvec + (size_t)b * p.kp4 + q * 4));
    }
}
__device__ __forceinline__ void item_compute(const BwdParams& p, int q, const ItemLoads& L, float4& outA,
                                             float4& outB) {
    const float v[4] = {L.v4.x, L.v4.y, L.v4.z, L.v4.w}, s[4] = {L.s4.x, L.s4.y, L.s4.z, L.s4.w},
                g[4] = {L.g4.x, L.g4.y, L.g4.z, L.g4.w};
    float a[4], c[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int j = q * 4 + t;
        a[t] = 0.f; c[t] = 0.f;
        if (j < p.k) {
            const float ej = __fmul_rn(v[t], L.x);
            if (p.use_fm2) a[t] = __fmul_rn(__fsub_rn(__fmul_rn(L.d, s[t]), __fmul_rn(L.d, ej)), L.x);
            if (p.gvec) c[t] = __fmul_rn(__fsub_rn(__fmul_rn(g[t], s[t]), __fmul_rn(g[t], ej)), L.x);
        } else if (j == p.k) {
            a[t] = __fmul_rn(L.d, L.x);
        }
    }
    outA = make_float4(a[0], a[1], a[2], a[3]);
    outB = make_float4(c[0], c[1], c[2], c[3]);
}

constexpr int ITEM_BATCH = 3;

__global__ void __launch_bounds__(256, 4) fm_bwd_tile_kernel(BwdParams p) {
    extern __shared__ __align__(16) float smem[];
    const int TE = p.TE, WN = 2 * TE, rowp = p.rowp, C = rowp >> 2;
    const bool two = p.use_fm2 && p.gvec;
    float* bufA = smem;                                       // [WN][rowp]
    float* bufB = bufA + (two ? (size_t)WN * rowp : 0);       // [WN][rowp] when both chains exist
    float* rowv = bufB + (size_t)WN * rowp;                   // [TE][rowp] old row values, per run start
    int32_t* keys_s = reinterpret_cast<int32_t*>(rowv + (size_t)TE * rowp);  // [WN+1]
    int32_t* rs_s = keys_s + (WN + 1);                        // run starts [WN+1]
    __shared__ int warp_cnt[8];
    __shared__ int carry_s, nruns_tile_s;

    const int64_t t0 = (int64_t)blockIdx.x * TE;
    const int W = (int)min((int64_t)WN, p.N - t0);
    const int TEe = min(TE, W);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (int i = threadIdx.x; i < WN; i += 256) keys_s[i + 1] = (i < W) ? __ldg(p.skeys + t0 + i) : -2;
    if (threadIdx.x == 0) {
        keys_s[0] = t0 > 0 ? __ldg(p.skeys + t0 - 1) : -1;
        carry_s = 0;
        nruns_tile_s = 0;
    }
    __syncthreads();

    // run starts of the window, in order (ballot scan, 256 positions per round)
    for (int base = 0; base < WN; base += 256) {
        const int i = base + threadIdx.x;
        const bool f = (i < W) && keys_s[i + 1] != keys_s[i];
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int wpre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const int c = warp_cnt[w]; if (w < warp) wpre += c; tot += c; }
        const int idx = carry_s + wpre + __popc(bal & ((1u << lane) - 1u));
        if (f) {
            rs_s[idx] = i;
            if (i < TEe) atomicMax(&nruns_tile_s, idx + 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s += tot;
        __syncthreads();
    }
    const int nruns_total = carry_s;
    int nruns = nruns_tile_s;  // runs that start inside the tile
    if (nruns == 0) return;
    if (threadIdx.x == 0) rs_s[nruns_total] = W;
    __syncthreads();
    // the last run of the tile is "long" when it reaches the end of a truncated window
    {
        const int last_end = rs_s[nruns];  // == W when no later start exists
        if (last_end == W && t0 + W < p.N) {
            if (threadIdx.x == 0) {
                const int slot = atomicAdd(p.long_count, 1);
                p.long_list[slot] = (int32_t)(t0 + rs_s[nruns - 1]);
            }
            nruns -= 1;
            if (nruns == 0) return;
        }
    }
    const int lo = rs_s[0], hi = rs_s[nruns];

    // contributions, one thread per (entry, 16-byte chunk); loads of ITEM_BATCH items are issued together
    const int nitems = (hi - lo) * C;
    for (int it0 = threadIdx.x; it0 < nitems; it0 += 256 * ITEM_BATCH) {
        int32_t pe[ITEM_BATCH];
        ItemLoads L[ITEM_BATCH];
#pragma unroll
        for (int u = 0; u < ITEM_BATCH; ++u) {
            const int it = it0 + u * 256;
            pe[u] = (it < nitems) ? __ldg(p.perm + t0 + lo + it / C) : 0;
        }
#pragma unroll
        for (int u = 0; u < ITEM_BATCH; ++u) {
            const int it = it0 + u * 256;
            if (it < nitems) item_load(p, keys_s[lo + it / C + 1], pe[u], it % C, L[u]);
        }
#pragma unroll
        for (int u = 0; u < ITEM_BATCH; ++u) {
            const int it = it0 + u * 256;
            if (it < nitems) {
                const int i = lo + it / C, q = it % C;
                float4 a, c;
                item_compute(p, q, L[u], a, c);
                store_contrib(p, two, bufA, bufB, i, q, a, c);
                if (i < TE && keys_s[i + 1] != keys_s[i])  // run start: keep the old row for the update
                    *reinterpret_cast<float4*>(rowv + (size_t)i * rowp + q * 4) = L[u].v4;
            }
        }
    }
    __syncthreads();

    // one thread per (run, component): left-to-right sum, then the row update
    const int kc = p.k + 1;
    for (int it = threadIdx.x; it < nruns * kc; it += 256) {
        const int r = it / kc, c = it - r * kc;
        const int s = rs_s[r], e = rs_s[r + 1];
        float acc = 0.f;
        for (int i = s; i < e; ++i) acc = __fadd_rn(acc, bufA[(size_t)i * rowp + c]);
        if (two && c < p.k) {
            float accB = 0.f;
            for (int i = s; i < e; ++i) accB = __fadd_rn(accB, bufB[(size_t)i * rowp + c]);
            acc = __fadd_rn(acc, accB);
        }
        p.table[(size_t)keys_s[s + 1] * rowp + c] = fmb::apply_update(rowv[(size_t)s * rowp + c], acc, p.lr, p.mode);
    }
}

// One CTA per long run.  Warps 1..7 produce the contributions of chunk c+1 into one half of a
// double buffer while warp 0 walks chunk c left to right (one lane per component / chain), so the
// serial fp32 chain -- the only part that cannot be parallelised without changing the rounding --
// runs back to back.  Falls back to produce-then-consume when the components do not fit one warp.
__global__ void __launch_bounds__(256) fm_bwd_long_kernel(BwdParams p) {
    extern __shared__ __align__(16) float smem[];
    const int rowp = p.rowp, C = rowp >> 2, CH = p.TE * 2;
    const bool two = p.use_fm2 && p.gvec;
    const int kc = p.k + 1;
    const int nacc = kc + (two ? p.k : 0);   // accumulator lanes: chain A comps, then chain B comps
    const bool piped = nacc <= 32;
    const size_t half = (size_t)CH * rowp * (two ? 2 : 1);
    __shared__ int64_t end_s;
    __shared__ int first_s[8];
    const int nlong = *p.long_count;
    const int warp = threadIdx.x >> 5;
    for (int li = blockIdx.x; li < nlong; li += gridDim.x) {
        const int64_t start = p.long_list[li];
        const int32_t key = __ldg(p.skeys + start);
        if (threadIdx.x == 0) end_s = -1;
        __syncthreads();
        for (int64_t base = start; end_s < 0; base += 256) {
            const int64_t pos = base + threadIdx.x;
            const bool mis = pos >= p.N || __ldg(p.skeys + pos) != key;
            const unsigned bal = __ballot_sync(0xffffffffu, mis);
            if ((threadIdx.x & 31) == 0) first_s[warp] = bal ? __ffs(bal) - 1 : -1;
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int w = 0; w < 8; ++w)
                    if (first_s[w] >= 0) { end_s = base + w * 32 + first_s[w]; break; }
            }
            __syncthreads();
        }
        const int64_t end = end_s;
        const int nchunks = (int)((end - start + CH - 1) / CH);
        float acc = 0.f;
        // accumulator thread -> (buffer, component)
        const int at = threadIdx.x;
        const bool is_acc = at < nacc;
        const bool accB = is_acc && at >= kc;
        const int acomp = accB ? at - kc : at;

        auto produce = [&](int c, int tid, int nthreads) {
            const int64_t cb = start + (int64_t)c * CH;
            const int n = (int)min((int64_t)CH, end - cb);
            float* bA = smem + (size_t)(c & 1) * half;
            float* bB = bA + (two ? (size_t)CH * rowp : 0);
            const int nitems = n * C;
            for (int it0 = tid; it0 < nitems; it0 += nthreads * ITEM_BATCH) {
                int32_t pe[ITEM_BATCH];
                ItemLoads L[ITEM_BATCH];
#pragma unroll
                for (int u = 0; u < ITEM_BATCH; ++u) {
                    const int it = it0 + u * nthreads;
                    pe[u] = (it < nitems) ? __ldg(p.perm + cb + it / C) : 0;
                }
#pragma unroll
                for (int u = 0; u < ITEM_BATCH; ++u) {
                    const int it = it0 + u * nthreads;
                    if (it < nitems) item_load(p, key, pe[u], it % C, L[u]);
                }
#pragma unroll
                for (int u = 0; u < ITEM_BATCH; ++u) {
                    const int it = it0 + u * nthreads;
                    if (it < nitems) {
                        float4 a, cc;
                        item_compute(p, it % C, L[u], a, cc);
                        store_contrib(p, two, bA, bB, it / C, it % C, a, cc);
                    }
                }
            }
        };
        auto consume = [&](int c) {
            const int64_t cb = start + (int64_t)c * CH;
            const int n = (int)min((int64_t)CH, end - cb);
            const float* b = smem + (size_t)(c & 1) * half + (accB ? (size_t)CH * rowp : 0) + acomp;
            int i = 0;
            for (; i + 8 <= n; i += 8) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = b[(size_t)(i + u) * rowp];
#pragma unroll
                for (int u = 0; u < 8; ++u) acc = __fadd_rn(acc, v[u]);
            }
            for (; i < n; ++i) acc = __fadd_rn(acc, b[(size_t)i * rowp]);
        };

        if (piped) {
            if (warp > 0) produce(0, threadIdx.x - 32, 224);
            __syncthreads();
            for (int c = 0; c < nchunks; ++c) {
                if (warp == 0) { if (is_acc) consume(c); }
                else if (c + 1 < nchunks) produce(c + 1, threadIdx.x - 32, 224);
                __syncthreads();
            }
        } else {
            for (int c = 0; c < nchunks; ++c) {
                produce(c, threadIdx.x, 256);
                __syncthreads();
                if (is_acc) consume(c);
                __syncthreads();
            }
        }
        // fold chain B into chain A and update the row
        float* fold = smem;  // reuse (all consumers are done: barrier above)
        if (accB) fold[acomp] = acc;
        __syncthreads();
        if (is_acc && !accB) {
            if (two && acomp < p.k) acc = __fadd_rn(acc, fold[acomp]);
            float* addr = p.table + (size_t)key * rowp + acomp;
            *addr = fmb::apply_update(*addr, acc, p.lr, p.mode);
        }
        __syncthreads();
    }
}

static int pick_te(int rowp, bool two) {
    int te = 256;
    while (te > 32 && (size_t)2 * te * rowp * 4 * (two ? 2 : 1) > 80 * 1024) te >>= 1;
    return te;
}
static size_t tile_smem(int te, int rowp, bool two) {
    return (size_t)2 * te * rowp * 4 * (two ? 2 : 1) + (size_t)te * rowp * 4 + (size_t)2 * (2 * te + 1) * 4 + 16;
}

}  // namespace

// workspace of fmb_fm_backward_update: the long-run list (<= N/TE + 1 entries) and its counter
FMB_API size_t fmb_bwd_workspace_bytes(int64_t N) { return ((size_t)(N / 32 + 2) * 4 + 255) / 256 * 256 + 256; }

// A6 sparse backward + update (see file header).
//   sorted_keys/perm [N]: output of fmb_sort_segment over ids[B*F]; xv [B*F] or NULL (ones)
//   table [R,rowp] updated in place; S [B,kp4]; gs [B]; gvec [B,kp4] or NULL
//   use_fm2: the scalar gs also flows through Sum_j bi (FM / DeepFM logit); 0 for NFM
//   mode 0: fresh-Adam sign step (reference), 1: SGD
FMB_API int fmb_fm_backward_update(const int32_t* sorted_keys, const int32_t* perm, int64_t N, const float* xv,
                                   float* table, int F, int k, const float* S, const float* gs, int use_fm2,
                                   const float* gvec, float lr, int mode, void* ws, size_t ws_bytes,
                                   cudaStream_t stream) {
    FMB_CHECK_ARG(sorted_keys && perm && table && S && gs && ws, "fmb_fm_backward_update: null pointer");
    FMB_CHECK_ARG(N > 0 && F > 0 && k > 0 && k < 256, "fmb_fm_backward_update: bad shape");
    FMB_CHECK_ARG(mode == 0 || mode == 1, "fmb_fm_backward_update: unknown update mode %d", mode);
    FMB_CHECK_ARG(use_fm2 || gvec, "fmb_fm_backward_update: neither gradient path enabled");
    if (ws_bytes < fmb_bwd_workspace_bytes(N)) { fmb_set_error("fmb_fm_backward_update: workspace too small"); return FMB_ERR_WS; }
    BwdParams p;
    p.skeys = sorted_keys; p.perm = perm; p.N = N; p.xv = xv; p.table = table;
    p.F = F; p.k = k; p.rowp = fmb_round_up(k + 1, 4); p.kp4 = fmb_round_up(k, 4);
    p.S = S; p.gs = gs; p.use_fm2 = use_fm2; p.gvec = gvec; p.lr = lr; p.mode = mode;
    p.long_count = (int32_t*)ws;
    p.long_list = (int32_t*)((char*)ws + 256);
    const bool two = use_fm2 && gvec;
    p.TE = pick_te(p.rowp, two);
    const size_t sm = tile_smem(p.TE, p.rowp, two);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(fm_bwd_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(fm_bwd_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr = true;
    }
    cudaMemsetAsync(p.long_count, 0, 4, stream);
    const int grid = (int)((N + p.TE - 1) / p.TE);
    fm_bwd_tile_kernel<<<grid, 256, sm, stream>>>(p);
    FMB_CHECK_LAUNCH("fm_bwd_tile_kernel");
    const size_t lsm = (size_t)2 * (2 * p.TE) * p.rowp * 4 * (two ? 2 : 1);  // double-buffered chunks of 2*TE
    fm_bwd_long_kernel<<<296, 256, lsm, stream>>>(p);
    FMB_CHECK_LAUNCH("fm_bwd_long_kernel");
    return FMB_OK;
}
