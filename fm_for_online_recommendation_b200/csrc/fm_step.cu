// fm_step.cu -- the FM-only training step (forward_fm + BCE-with-logits + sparse backward + row update) with the
// rows read ONCE: one kernel gathers the rows of a tile of samples, computes the logits, the loss gradient and
// -- because the stable sort of the batch's row ids has already told it which entries are the only hit of their
// row -- applies the update of those rows from the copy it holds in shared memory.  Entries of rows hit several
// times write their contribution to the component-major staging buffer G at their SORTED position; the run
// kernel of fm_backward.cu then sums each run in sample order (the reference's embedding_dense_backward order)
// and updates those rows.
//
// Replaces, for FMAdam.update_embedding / fit and the update_embedding of the other four classes
// (models/models_online_deep/fm_adam.py:56-82, deepfm_adam.py:91-104, nfm_adam.py:90-103, deepfm_onn.py:156-169,
// nfm_onn.py:158-171), the round-1 sequence fm_forward_kernel -> fm_bwd_entry1_kernel (which read every row a
// second time: 31 MB of the step's 68 MB of DRAM traffic at cfg5) -> fm_bwd_runs_kernel.
//
// Safety of updating inside the forward kernel: a row that is hit by exactly one entry of the batch is read by
// exactly one CTA -- the one that updates it -- so no other sample's forward pass can observe the new value.
// Algorithmic bytes per sample (k = 10, F = 39): ids 4F + position/flag words 4F + rows read 4F(k+1) + rows
// written 4F(k+1) + delta/loss 8 = 3 752 B (SURVEY.md 8d's figure; the sorted-position words replace the
// sorted-key reads of the entry kernel).
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

// The four per-batch pointers are separate kernel arguments (not members of StepParams) so that a captured CUDA
// graph of the step can be re-pointed at another batch with cudaGraphExecKernelNodeSetParams (session.cu).
struct StepParams {
    float* table;              // [R,rowp]
    const float* bias;         // [1]
    int B, F, k, rowp, kp4, SB;
    int cu;                    // 16-byte chunks of a row that hold data: ceil((k+1)/4)
    int ql_log;                // log2 of lanes per (sample, field) in the gather phase (pow2 >= cu)
    int jl_log;                // log2 of lanes per sample in the reduce phase (pow2 >= kp4)
    int loss_kind, mode;
    float lr, astep;
    float* delta;              // [B] out
    float* lossv;              // [B] out
    float* G;                  // [k+1][Npad] staged contributions of multi-hit entries
    int64_t Npad;
};

//   ids [B,F] global row ids;  xv [B,F] or NULL (all ones);  y [B];
//   posflag [B*F] entry-major: sorted position of the entry | 0x80000000 if its row is hit more than once
__global__ void __launch_bounds__(256) fm_step_fused_kernel(const int32_t* __restrict__ ids, const float* __restrict__ xv,
                                                            const float* __restrict__ y, const uint32_t* __restrict__ posflag,
                                                            StepParams p) {
    extern __shared__ __align__(16) float smem[];
    const int F = p.F, k = p.k, SB = p.SB;
    const int rp = p.cu * 4;                       // shared-memory row pitch (floats)
    float* rows_s = smem;                          // [SB][F][rp]
    float* x_s = rows_s + (size_t)SB * F * rp;     // [SB][F]
    float* bi_s = x_s + SB * F;                    // [SB][k]
    float* S_s = bi_s + SB * k;                    // [SB][kp4]
    float* d_s = S_s + SB * p.kp4;                 // [SB]
    int32_t* ids_s = reinterpret_cast<int32_t*>(d_s + SB);            // [SB][F]
    uint32_t* pos_s = reinterpret_cast<uint32_t*>(ids_s + SB * F);    // [SB][F]
    const int b0 = blockIdx.x * SB;
    const int nv = min(SB, p.B - b0);

    // phase 0: the tile's row ids, values and sorted positions, coalesced
    for (int e = threadIdx.x; e < nv * F; e += blockDim.x) {
        ids_s[e] = __ldg(ids + (size_t)b0 * F + e);
        x_s[e] = xv ? __ldg(xv + (size_t)b0 * F + e) : 1.0f;
        pos_s[e] = __ldg(posflag + (size_t)b0 * F + e);
    }
    __syncthreads();
    // phase 1: gather rows, 16 B per cp.async; every row read of the tile is in flight at once
    {
        const int q = threadIdx.x & ((1 << p.ql_log) - 1);
        const int estep = blockDim.x >> p.ql_log;
        if (q < p.cu)
            for (int ef = threadIdx.x >> p.ql_log; ef < nv * F; ef += estep)
                cp_async16(rows_s + (size_t)ef * rp + q * 4, p.table + (size_t)ids_s[ef] * p.rowp + q * 4);
    }
    cp_async_wait_all();
    __syncthreads();

    // phase 2: one thread per (sample, component): S = sum_f e_f, Q = sum_f e_f^2, left to right (python sum() order)
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

    // phase 3: one thread per sample: logit in ATen's row-sum order, loss value, gradient on the logit
    if (threadIdx.x < nv) {
        const int s = threadIdx.x, b = b0 + s;
        const float* r = rows_s + (size_t)s * F * rp + k;
        const float* xs = x_s + s * F;
        const float sf = fmb::aten_row_sum_small([&](int f) { return __fmul_rn(r[(size_t)f * rp], xs[f]); }, F);
        const float* bs = bi_s + s * k;
        const float sb = fmb::aten_row_sum_small([&](int j) { return bs[j]; }, k);
        const float z = __fadd_rn(__fadd_rn(sf, sb), __ldg(p.bias));
        float lv, d;
        fmb::bce_logits_value_grad(p.loss_kind, z, y[b], b, p.B, lv, d);
        p.lossv[b] = lv;
        p.delta[b] = d;
        d_s[s] = d;
    }
    __syncthreads();

    // phase 4: one thread per (entry, 16-byte chunk): gradient contribution of the entry; single-hit rows are
    // updated here from the shared-memory copy, the others stage their contribution at their sorted position
    const int items = nv * F * p.cu;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int ef = it / p.cu, q = it - ef * p.cu;
        const int s = ef / F;
        const float x = x_s[ef], d = d_s[s];
        const uint32_t pf = pos_s[ef];
        const float4 v4 = *reinterpret_cast<const float4*>(rows_s + (size_t)ef * rp + q * 4);
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
        float a[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int j = q * 4 + t;
            a[t] = 0.f;
            if (j < k) {
                const float ej = __fmul_rn(v[t], x);
                a[t] = __fmul_rn(__fsub_rn(__fmul_rn(d, S_s[s * p.kp4 + j]), __fmul_rn(d, ej)), x);
            } else if (j == k) {
                a[t] = __fmul_rn(d, x);
            }
        }
        if (!(pf >> 31)) {
            float o[4];
#pragma unroll
            for (int t = 0; t < 4; ++t)   // sum over the row's single entry = 0 + contribution
                o[t] = (q * 4 + t <= k) ? fmb::apply_update_a(v[t], __fadd_rn(0.f, a[t]), p.lr, p.astep, p.mode) : v[t];
            *reinterpret_cast<float4*>(p.table + (size_t)ids_s[ef] * p.rowp + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
            const size_t pos = pf & 0x7fffffffu;
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (q * 4 + t <= k) p.G[(size_t)(q * 4 + t) * p.Npad + pos] = a[t];
        }
    }
}

// posflag[perm[i]] = i | (row of sorted position i is hit more than once ? 0x80000000 : 0)
__global__ void __launch_bounds__(256) pos_flags_kernel(const int32_t* __restrict__ skeys, const int32_t* __restrict__ perm,
                                                        int64_t N, uint32_t* __restrict__ posflag) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const int32_t key = __ldg(skeys + i);
    const int32_t kprev = i > 0 ? __ldg(skeys + i - 1) : -1;
    const int32_t knext = i + 1 < N ? __ldg(skeys + i + 1) : -1;
    const bool multi = key == kprev || key == knext;
    posflag[__ldg(perm + i)] = (uint32_t)i | (multi ? 0x80000000u : 0u);
}

static int ilog2_ceil(int x) { int l = 0; while ((1 << l) < x) ++l; return l; }

}  // namespace

// host-side handle of the fused kernel, for graph-node identification in session.cu (arguments 0..3 are ids, xv, y, posflag)
FMB_API const void* fmb_fused_kernel_fn(void) { return (const void*)fm_step_fused_kernel; }
FMB_API const void* fmb_pos_flags_kernel_fn(void) { return (const void*)pos_flags_kernel; }

// per-entry sorted position + multi-hit flag from the stable sort's output (fmb_sort_fields / fmb_sort_segment)
FMB_API int fmb_pos_flags(const int32_t* sorted_keys, const int32_t* perm, int64_t N, uint32_t* posflag,
                          cudaStream_t stream) {
    FMB_CHECK_ARG(sorted_keys && perm && posflag && N > 0 && N < ((int64_t)1 << 31), "fmb_pos_flags: bad arguments");
    pos_flags_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(sorted_keys, perm, N, posflag);
    FMB_CHECK_LAUNCH("pos_flags_kernel");
    return FMB_OK;
}

// Forward + loss + single-hit row updates + staging of multi-hit contributions (see file header).
//   posflag [B*F]: output of fmb_pos_flags for THIS batch's ids;  G: workspace of fmb_bwd_workspace_bytes(B*F, k)
//   bytes, handed to fmb_fm_backward_runs afterwards;  delta/lossv [B]: per-sample gradient and loss (inputs of
//   fmb_finish_step).  mode as in fmb_fm_backward_update.
FMB_API int fmb_fm_step_fused(const int32_t* ids, const float* xv, const float* y, float* table, const float* bias,
                              const uint32_t* posflag, int B, int F, int k, int loss_kind, float lr, int mode,
                              float* delta, float* lossv, void* ws, size_t ws_bytes, cudaStream_t stream) {
    FMB_CHECK_ARG(ids && y && table && bias && posflag && delta && lossv && ws, "fmb_fm_step_fused: null pointer");
    FMB_CHECK_ARG(B > 0 && F > 0 && F < 512 && k > 0 && k <= 124, "fmb_fm_step_fused: bad shape B=%d F=%d k=%d", B, F, k);
    FMB_CHECK_ARG(loss_kind == 0 || loss_kind == 1, "fmb_fm_step_fused: unknown loss kind %d", loss_kind);
    FMB_CHECK_ARG(mode == 0 || mode == 1, "fmb_fm_step_fused: unknown update mode %d", mode);
    const int64_t N = (int64_t)B * F;
    if (ws_bytes < fmb_bwd_workspace_bytes(N, k)) { fmb_set_error("fmb_fm_step_fused: workspace too small"); return FMB_ERR_WS; }
    StepParams p;
    p.table = table; p.bias = bias;
    p.B = B; p.F = F; p.k = k; p.rowp = fmb_round_up(k + 1, 16); p.kp4 = fmb_round_up(k, 4);
    p.cu = (k + 1 + 3) / 4; p.ql_log = ilog2_ceil(p.cu); p.jl_log = ilog2_ceil(p.kp4);
    p.loss_kind = loss_kind; p.mode = mode; p.lr = lr; p.astep = -(lr / 0.1f);
    p.delta = delta; p.lossv = lossv;
    p.G = (float*)ws; p.Npad = (N + 3) / 4 * 4 + 64;
    int SB = 256 >> p.jl_log;
    if (SB < 4) SB = 4;
    if (SB > 32) SB = 32;
    auto bytes = [&](int sb) {
        return sizeof(float) * ((size_t)sb * F * p.cu * 4 + (size_t)3 * sb * F + (size_t)sb * k + (size_t)sb * p.kp4 + sb);
    };
    while (SB > 1 && bytes(SB) > 48 * 1024) SB >>= 1;
    FMB_CHECK_ARG(bytes(SB) <= 200 * 1024, "fmb_fm_step_fused: F*k too large for one sample tile");
    p.SB = SB;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(fm_step_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    fm_step_fused_kernel<<<(B + SB - 1) / SB, 256, bytes(SB), stream>>>(ids, xv, y, posflag, p);
    FMB_CHECK_LAUNCH("fm_step_fused_kernel");
    return FMB_OK;
}
