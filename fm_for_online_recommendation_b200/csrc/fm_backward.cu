// fm_backward.cu -- sparse embedding gradient as a deterministic segmented reduce over the sorted
// (row id, entry) list, fused with the per-row parameter update.
//
// Replaces autograd + 2F dense `embedding_dense_backward` + the dense `Adam.step` over all rows
// (models/models_online_deep/fm_adam.py:56-69, deepfm_adam.py:91-117): only
// the rows a batch touches are read and written.  Per entry (b, f) of row r:
//     e_j   = V_r[j] * x                                     (deepfm_adam.py:60)
//     g_e_j = (g*S_b[j]) - (g*e_j)      g = delta_b (FM scalar path) and/or gvec_b[j] (MLP path)
//     grad V_r[j] += g_e_j * x ;  grad w_r += delta_b * x    summed IN SAMPLE ORDER per row
// then one update per row (fresh-Adam sign step or SGD).  No float atomics anywhere.
//
// Two kernels over the sorted list:
//   fm_bwd_entry_kernel : one 4-lane group per entry (one lane per 16-byte chunk of the row).  A row
//       hit by a single entry of the batch (the common case in large fields) is updated right here
//       from registers; entries of rows hit several times write their contribution to a staging
//       buffer G (L2-resident).
//   fm_bwd_runs_kernel  : one warp per 32 sorted positions finds the runs (>= 2 entries) that START
//       there and sums each run left to right, one lane per component -- the first 32 entries straight
//       from G, longer runs through a 4-stage cp.async ring in shared memory so the serial fp32 chain
//       (the only part that cannot be parallelised without changing the rounding) never waits on L2.
#include "fmb_common.cuh"

namespace {

struct BwdParams {
    const int32_t* skeys;
    const int32_t* perm;
    int64_t N;
    const float* xv;
    float* table;
    int F, k, rowp, kp4;
    int cu;               // 16-byte chunks per row that hold data: ceil((k+1)/4)
    int ql_log;           // log2(lanes per entry), 2^ql_log >= cu
    unsigned long long fmagic;  // ceil(2^fshift / F): b = (e * fmagic) >> fshift, exact for e < N (2^fshift > N*F)
    int fshift;
    const float* S;
    const float* gs;
    int use_fm2;
    const float* gvec;
    float lr;
    int mode;
    int s_pitch;          // row pitch of S / gvec in floats (kp4 unless they live in a gathered context)
    int gs_stride;        // stride of gs in floats
    int32_t key_limit;    // keys >= key_limit are padding (sharded path) and are skipped
    float* G;             // [N][cu*4] staged contributions (chain A, or the only chain)
    float* G2;            // [N][cu*4] chain B when both gradient paths are live
    long long* dbg;       // optional per-run timing records (debug builds of bench only), else NULL
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fm_bwd_entry_kernel(BwdParams p) {
    const int q = threadIdx.x & ((1 << p.ql_log) - 1);
    const int64_t i = (int64_t)blockIdx.x * (256 >> p.ql_log) + (threadIdx.x >> p.ql_log);
    if (i >= p.N || q >= p.cu) return;
    const int32_t key = __ldg(p.skeys + i);
    if (key >= p.key_limit) return;
    const int32_t kprev = i > 0 ? __ldg(p.skeys + i - 1) : -1;
    const int32_t knext = i + 1 < p.N ? __ldg(p.skeys + i + 1) : -1;
    const int32_t e = __ldg(p.perm + i);
    const int b = (int)(((unsigned long long)(unsigned)e * p.fmagic) >> p.fshift);
    const float x = p.xv ? __ldg(p.xv + e) : 1.0f;
    const float d = __ldg(p.gs + (size_t)b * p.gs_stride);
    float* rowptr = p.table + (size_t)key * p.rowp + q * 4;
    const float4 v4 = *reinterpret_cast<const float4*>(rowptr);
    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = s4;
    if (q * 4 < p.kp4) {
        s4 = __ldg(reinterpret_cast<const float4*>(p.S + (size_t)b * p.s_pitch + q * 4));
        if (p.gvec) g4 = __ldg(reinterpret_cast<const float4*>(p.gvec + (size_t)b * p.s_pitch + q * 4));
    }
    const float v[4] = {v4.x, v4.y, v4.z, v4.w}, s[4] = {s4.x, s4.y, s4.z, s4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w};
    const bool two = p.use_fm2 && p.gvec;
    float a[4], c[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int j = q * 4 + t;
        a[t] = 0.f; c[t] = 0.f;
        if (j < p.k) {
            const float ej = __fmul_rn(v[t], x);
            if (p.use_fm2) a[t] = __fmul_rn(__fsub_rn(__fmul_rn(d, s[t]), __fmul_rn(d, ej)), x);
            if (p.gvec) c[t] = __fmul_rn(__fsub_rn(__fmul_rn(g[t], s[t]), __fmul_rn(g[t], ej)), x);
        } else if (j == p.k) {
            a[t] = __fmul_rn(d, x);
        }
    }
    if (key != kprev && key != knext) {
        // the only entry of its row: sum = 0 + contribution; update from registers
        float o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int j = q * 4 + t;
            if (j < p.k) {
                const float gr = two ? __fadd_rn(__fadd_rn(0.f, a[t]), __fadd_rn(0.f, c[t]))
                                     : __fadd_rn(0.f, p.gvec ? c[t] : a[t]);
                o[t] = fmb::apply_update(v[t], gr, p.lr, p.mode);
            } else if (j == p.k) {
                o[t] = fmb::apply_update(v[t], __fadd_rn(0.f, a[t]), p.lr, p.mode);
            } else {
                o[t] = v[t];
            }
        }
        *reinterpret_cast<float4*>(rowptr) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
        const int gp = p.cu * 4;
        if (two) {
            *reinterpret_cast<float4*>(p.G + (size_t)i * gp + q * 4) = make_float4(a[0], a[1], a[2], a[3]);
            *reinterpret_cast<float4*>(p.G2 + (size_t)i * gp + q * 4) = make_float4(c[0], c[1], c[2], c[3]);
        } else {
            float m[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) m[t] = (p.gvec && q * 4 + t < p.k) ? c[t] : a[t];
            *reinterpret_cast<float4*>(p.G + (size_t)i * gp + q * 4) = make_float4(m[0], m[1], m[2], m[3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
constexpr int RING_SE = 32;  // entries per ring stage (one key per lane)
constexpr int RING_NS = 8;   // stages

__global__ void __launch_bounds__(256, 3) fm_bwd_runs_kernel(BwdParams p, int warps_per_block, int warp_f, int accs_n) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t P0 = ((int64_t)blockIdx.x * warps_per_block + wib) * 32;
    if (P0 >= p.N) return;
    const bool two = p.use_fm2 && p.gvec;
    const int nbuf = two ? 2 : 1;
    const int gp = p.cu * 4, kc = p.k + 1;
    const int nv = kc + (two ? p.k : 0);          // virtual lanes: chain A comps, then chain B comps
    const int stage_f = RING_SE * gp * nbuf;
    float* ring = smem + (size_t)wib * warp_f;    // [NS][nbuf][SE][gp]
    int32_t* rkeys = reinterpret_cast<int32_t*>(ring + RING_NS * stage_f);  // [NS][SE]
    float* accs = reinterpret_cast<float*>(rkeys + RING_NS * RING_SE);      // [accs_n]

    // keys of this warp's 32 positions and of the 32 after them (one memory latency for both)
    const int64_t pos = P0 + lane;
    const int32_t k0 = pos < p.N ? __ldg(p.skeys + pos) : -2;
    const int32_t k1 = pos + 32 < p.N ? __ldg(p.skeys + pos + 32) : -2;
    int32_t prev = __shfl_up_sync(0xffffffffu, k0, 1);
    int32_t next = __shfl_down_sync(0xffffffffu, k0, 1);
    const int32_t k1_0 = __shfl_sync(0xffffffffu, k1, 0);
    if (lane == 0) prev = P0 > 0 ? __ldg(p.skeys + P0 - 1) : -1;
    if (lane == 31) next = k1_0;
    // runs (>= 2 entries) that START inside these 32 positions
    unsigned todo = __ballot_sync(0xffffffffu, pos < p.N && k0 >= 0 && k0 < p.key_limit && k0 != prev && k0 == next);

    // per-lane constants of the ring fill: lane -> (entry within a group of 32>>ql_log, 16-byte chunk)
    const int fl_q = lane & ((1 << p.ql_log) - 1);
    const int fl_e = lane >> p.ql_log;
    const int fl_iters = 1 << p.ql_log;                 // groups of (32 >> ql_log) entries per stage
    const int fl_estep = 32 >> p.ql_log;
    const bool fl_on = fl_q < p.cu;
    const int fl_off = fl_e * gp + fl_q * 4;            // float offset inside a stage / inside G

    while (todo) {
        const int bit = __ffs(todo) - 1;
        todo &= todo - 1;
        const int64_t s = P0 + bit;
        const int32_t key = __shfl_sync(0xffffffffu, k0, bit);
        long long t_start = 0, t_direct = 0, t_ring = 0;
        if (p.dbg) t_start = clock64();
        int run_len = 0;
        // leading matches among the first 32 entries of the run (keys are already in registers)
        const int t = bit + lane;
        const int32_t ka = __shfl_sync(0xffffffffu, k0, t & 31);
        const int32_t kb = __shfl_sync(0xffffffffu, k1, t & 31);
        const unsigned mm = __ballot_sync(0xffffffffu, (t < 32 ? ka : kb) == key);
        const int n0 = (mm == 0xffffffffu) ? 32 : __ffs(~mm) - 1;

        for (int v0 = 0; v0 < nv; v0 += 32) {  // one pass per group of 32 (buffer, component) lanes
            const int vl = v0 + lane;
            const bool active = vl < nv;
            const bool isB = vl >= kc;
            const int comp = isB ? vl - kc : vl;
            const float* src = (isB ? p.G2 : p.G) + comp;
            // the row's old value is needed only at the very end: fetch it now, off the critical path
            float pold = 0.f;
            if (v0 == 0 && lane < kc) pold = p.table[(size_t)key * p.rowp + lane];
            // long run: start streaming entries 32.. into the shared-memory ring right away.
            // A single warp executes this chain alone, so the fill path is kept to a handful of
            // instructions: pointers advance by constants, bounds are checked once per stage.
            int64_t fill = s + 32;
            const float* gsrc = p.G + (size_t)fill * gp + fl_off;
            const float* gsrc2 = two ? p.G2 + (size_t)fill * gp + fl_off : nullptr;
            auto issue = [&](int st) {
                float* dst = ring + (size_t)st * stage_f + fl_off;
                if (fill + RING_SE <= p.N) {
                    if (fl_on) {
                        for (int it = 0; it < fl_iters; ++it) {
                            cp_async16(dst + it * fl_estep * gp, gsrc + it * fl_estep * gp);
                            if (two) cp_async16(dst + (RING_SE + it * fl_estep) * gp, gsrc2 + it * fl_estep * gp);
                        }
                    }
                    cp_async4(rkeys + st * RING_SE + lane, p.skeys + fill + lane);
                } else {
                    for (int it = 0; it < fl_iters; ++it) {
                        if (fl_on && fill + fl_e + it * fl_estep < p.N) {
                            cp_async16(dst + it * fl_estep * gp, gsrc + it * fl_estep * gp);
                            if (two) cp_async16(dst + (RING_SE + it * fl_estep) * gp, gsrc2 + it * fl_estep * gp);
                        }
                    }
                    if (fill + lane < p.N) cp_async4(rkeys + st * RING_SE + lane, p.skeys + fill + lane);
                    else rkeys[st * RING_SE + lane] = -2;
                }
                cp_async_commit();
                fill += RING_SE;
                gsrc += RING_SE * gp;
                if (two) gsrc2 += RING_SE * gp;
            };
            if (n0 == 32) {
#pragma unroll
                for (int st = 0; st < RING_NS; ++st) issue(st);
            }
            // direct part: up to 32 entries straight from G, all loads in flight at once
            float acc = 0.f;
            {
                const float* sp = src + (size_t)s * gp;
                float tv[16], tw[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) tv[u] = (active && u < n0) ? __ldg(sp + u * gp) : 0.f;
#pragma unroll
                for (int u = 0; u < 16; ++u) tw[u] = (active && 16 + u < n0) ? __ldg(sp + (16 + u) * gp) : 0.f;
#pragma unroll
                for (int u = 0; u < 16; ++u)
                    if (u < n0) acc = __fadd_rn(acc, tv[u]);
#pragma unroll
                for (int u = 0; u < 16; ++u)
                    if (16 + u < n0) acc = __fadd_rn(acc, tw[u]);
            }
            if (p.dbg) t_direct = clock64();
            run_len = n0;
            if (n0 == 32) {
                int st = 0;
                while (true) {
                    cp_async_wait<RING_NS - 1>();
                    __syncwarp();
                    const unsigned m2 = __ballot_sync(0xffffffffu, rkeys[st * RING_SE + lane] == key);
                    const int n = (m2 == 0xffffffffu) ? 32 : __ffs(~m2) - 1;
                    if (active) {
                        const float* b = ring + (size_t)st * stage_f + (isB ? RING_SE * gp : 0) + comp;
#pragma unroll 1
                        for (int h = 0; h < 32; h += 16) {
                            float tv[16];
#pragma unroll
                            for (int u = 0; u < 16; ++u) tv[u] = b[(h + u) * gp];
#pragma unroll
                            for (int u = 0; u < 16; ++u)
                                if (h + u < n) acc = __fadd_rn(acc, tv[u]);
                        }
                    }
                    __syncwarp();
                    run_len += n;
                    if (n < 32) break;
                    issue(st);
                    st = (st + 1 == RING_NS) ? 0 : st + 1;
                }
                cp_async_wait<0>();
                __syncwarp();
            }
            if (p.dbg) t_ring = clock64();
            if (active) accs[vl] = acc;
            if (v0 == 0 && lane < kc) accs[accs_n + lane] = pold;
        }
        __syncwarp();
        // fold chain B into chain A and update the row (one lane per component)
        for (int c0 = 0; c0 < kc; c0 += 32) {
            const int c = c0 + lane;
            if (c < kc) {
                float gsum = accs[c];
                if (two && c < p.k) gsum = __fadd_rn(gsum, accs[kc + c]);
                float* addr = p.table + (size_t)key * p.rowp + c;
                const float old = c < 32 ? accs[accs_n + c] : *addr;
                *addr = fmb::apply_update(old, gsum, p.lr, p.mode);
            }
        }
        __syncwarp();
        if (p.dbg && lane == 0) {
            const unsigned long long slot = atomicAdd((unsigned long long*)p.dbg, 1ULL);
            if (slot < 4000) {
                long long* r = p.dbg + 8 + slot * 8;
                r[0] = run_len; r[1] = t_start; r[2] = t_direct - t_start; r[3] = t_ring - t_direct;
                r[4] = clock64() - t_ring; r[5] = 0; r[6] = 0; r[7] = 0;
            }
        }
    }
}

static int ilog2_ceil(int x) { int l = 0; while ((1 << l) < x) ++l; return l; }

}  // namespace

// workspace of fmb_fm_backward_update: the contribution staging buffers G and G2
FMB_API size_t fmb_bwd_workspace_bytes(int64_t N, int k) {
    const size_t gp = (size_t)((k + 1 + 3) / 4) * 4;
    return 2 * (((size_t)N * gp * 4 + 255) / 256 * 256) + 256;
}

static long long* g_runs_dbg = nullptr;
// debug hook (not in the public header): device buffer of 8 + 4000*8 int64 receiving per-run cycle counts
FMB_API void fmb_debug_set_runs_buffer(long long* dev) { g_runs_dbg = dev; }

// A6 sparse backward + update (see file header).
//   sorted_keys/perm [N]: output of fmb_sort_segment / fmb_sort_fields over ids[B*F]; xv [B*F] or NULL
//   table [R,rowp] updated in place; S [B,kp4]; gs [B]; gvec [B,kp4] or NULL
//   use_fm2: the scalar gs also flows through Sum_j bi (FM / DeepFM logit); 0 for NFM
//   mode 0: fresh-Adam sign step (reference), 1: SGD
// _ex adds: n_entries = 1 + the largest entry index stored in perm (B*F; the sharded path passes
//   world*B*F), s_pitch / gs_stride (S, gvec and gs may live inside a gathered per-sample context),
//   key_limit (sorted keys >= key_limit are padding and are skipped).
FMB_API int fmb_fm_backward_update_ex(const int32_t* sorted_keys, const int32_t* perm, int64_t N, int64_t n_entries,
                                      const float* xv, float* table, int F, int k, const float* S, int s_pitch,
                                      const float* gs, int gs_stride, int use_fm2, const float* gvec,
                                      int32_t key_limit, float lr, int mode, void* ws, size_t ws_bytes,
                                      cudaStream_t stream) {
    FMB_CHECK_ARG(sorted_keys && perm && table && S && gs && ws, "fmb_fm_backward_update: null pointer");
    FMB_CHECK_ARG(N > 0 && F > 0 && F < 512 && k > 0 && k <= 124, "fmb_fm_backward_update: bad shape");
    FMB_CHECK_ARG(mode == 0 || mode == 1, "fmb_fm_backward_update: unknown update mode %d", mode);
    FMB_CHECK_ARG(use_fm2 || gvec, "fmb_fm_backward_update: neither gradient path enabled");
    FMB_CHECK_ARG(n_entries > 0 && n_entries < ((int64_t)1 << 31), "fmb_fm_backward_update: n_entries out of range");
    if (ws_bytes < fmb_bwd_workspace_bytes(N, k)) { fmb_set_error("fmb_fm_backward_update: workspace too small"); return FMB_ERR_WS; }
    BwdParams p;
    p.skeys = sorted_keys; p.perm = perm; p.N = N; p.xv = xv; p.table = table;
    p.F = F; p.k = k; p.rowp = fmb_round_up(k + 1, 16); p.kp4 = fmb_round_up(k, 4);
    p.cu = (k + 1 + 3) / 4; p.ql_log = ilog2_ceil(p.cu);
    {
        int bn = 0, bf = 0;
        while (((int64_t)1 << bn) < n_entries) ++bn;
        while ((1 << bf) <= F) ++bf;
        p.fshift = bn + bf;  // 2^fshift > n_entries*F, and e*fmagic < 2^(2*bn+2) <= 2^64
        p.fmagic = ((1ULL << p.fshift) + (unsigned long long)F - 1) / (unsigned long long)F;
    }
    p.S = S; p.gs = gs; p.use_fm2 = use_fm2; p.gvec = gvec; p.lr = lr; p.mode = mode;
    p.s_pitch = s_pitch; p.gs_stride = gs_stride; p.key_limit = key_limit;
    p.dbg = g_runs_dbg;
    const size_t gbytes = ((size_t)N * p.cu * 16 + 255) / 256 * 256;
    p.G = (float*)ws;
    p.G2 = (float*)((char*)ws + gbytes);
    const bool two = use_fm2 && gvec;
    const int epb = 256 >> p.ql_log;
    fm_bwd_entry_kernel<<<(unsigned)((N + epb - 1) / epb), 256, 0, stream>>>(p);
    FMB_CHECK_LAUNCH("fm_bwd_entry_kernel");
    // runs kernel: per-warp shared memory = ring + keys + accumulators
    const int gp = p.cu * 4;
    const int nv = k + 1 + (two ? k : 0);
    const int accs_n = (nv + 3) / 4 * 4;
    const int warp_f = RING_NS * RING_SE * gp * (two ? 2 : 1) + RING_NS * RING_SE + accs_n + 32;
    int wpb = 8;
    while (wpb > 1 && (size_t)wpb * warp_f * 4 > 56 * 1024) wpb >>= 1;
    const size_t sm = (size_t)wpb * warp_f * 4;
    FMB_CHECK_ARG(sm <= 200 * 1024, "fmb_fm_backward_update: k too large for the run ring");
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(fm_bwd_runs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
    const int64_t nwarps = (N + 31) / 32;
    fm_bwd_runs_kernel<<<(unsigned)((nwarps + wpb - 1) / wpb), 32 * wpb, sm, stream>>>(p, wpb, warp_f, accs_n);
    FMB_CHECK_LAUNCH("fm_bwd_runs_kernel");
    return FMB_OK;
}

FMB_API int fmb_fm_backward_update(const int32_t* sorted_keys, const int32_t* perm, int64_t N, const float* xv,
                                   float* table, int F, int k, const float* S, const float* gs, int use_fm2,
                                   const float* gvec, float lr, int mode, void* ws, size_t ws_bytes,
                                   cudaStream_t stream) {
    return fmb_fm_backward_update_ex(sorted_keys, perm, N, N, xv, table, F, k, S, fmb_round_up(k, 4), gs, 1, use_fm2,
                                     gvec, 0x7fffffff, lr, mode, ws, ws_bytes, stream);
}
