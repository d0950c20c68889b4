// dense_ops.cu -- small dense pieces of the training step: loss/delta, ATen-order reductions,
// dense parameter updates (bias, MLP), and the fused end-of-step kernel.
// Reference: F.binary_cross_entropy_with_logits + loss.backward() + torch.optim.Adam.step in
// models/models_online_deep/fm_adam.py:60-68.
#include "fmb_common.cuh"

namespace {

__global__ void loss_delta_kernel(int kind, const float* __restrict__ z, const float* __restrict__ y, int B,
                                  float* __restrict__ delta, float* __restrict__ lossv) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float yy = y[b];
    float in = z[b], pr = 0.f;
    if (kind == 1) { pr = fmb::sigmoidf_p(in); in = pr; }
    const float ls = __fsub_rn(fminf(in, 0.f), fmb::log1pf_p(fmb::expf_p(-fabsf(in))));
    if (lossv) lossv[b] = __fsub_rn(__fmul_rn(__fsub_rn(1.0f, yy), in), ls);
    float d = __fdiv_rn(__fsub_rn(fmb::sigmoidf_p(in), yy), (float)B);
    if (kind == 1) d = __fmul_rn(__fmul_rn(d, __fsub_rn(1.0f, pr)), pr);
    delta[b] = d;
}

__global__ void sum_aten_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
    const float s = fmb::aten_row_sum_warp(x, n);
    if (threadIdx.x == 0) out[0] = s;
}

__global__ void update_dense_kernel(float* __restrict__ p, const float* __restrict__ g, int64_t n, float lr,
                                    int mode) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = fmb::apply_update(p[i], g[i], lr, mode);
}

// ATen-order sum of x[0:n] by one warp, with the data staged through shared memory in chunks of
// FIN_CHUNK floats by the whole CTA (coalesced, one latency per chunk) so the serial chain reads
// shared memory instead of L2.  Chunk boundaries are multiples of 32*16, so the cascade state
// simply carries over; bit-identical to fmb::aten_row_sum_warp.
constexpr int FIN_CHUNK = 4096;  // 128 rows of 32: a multiple of every cascade step <= 128
struct AtenAcc {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
};

__global__ void __launch_bounds__(1024) finish_step_kernel(const float* __restrict__ delta,
                                                           const float* __restrict__ lossv, int B, float* bias,
                                                           float lr, int mode, float* loss_out) {
    __shared__ float buf[2][FIN_CHUNK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* src[2] = {bias ? delta : nullptr, (loss_out && lossv) ? lossv : nullptr};
    const int64_t n = B;
    const int64_t vec_size = n >> 3, size_ilp = vec_size >> 2;   // rows of 32 floats handled by the cascade
    int lp = 0;
    while (((int64_t)1 << lp) < size_ilp) ++lp;
    lp >>= 2;
    const int level_power = lp > 4 ? lp : 4;
    const int64_t level_step = (int64_t)1 << level_power, level_mask = level_step - 1;
    AtenAcc acc;
    const bool small = n < 8;
    for (int64_t base = 0; base < n; base += FIN_CHUNK) {
        const int m = (int)min((int64_t)FIN_CHUNK, n - base);
        for (int i = threadIdx.x; i < m; i += blockDim.x) {
            if (src[0]) buf[0][i] = src[0][base + i];
            if (src[1]) buf[1][i] = src[1][base + i];
        }
        __syncthreads();
        if (warp < 2 && src[warp] && !small) {
            const float* b = buf[warp];
            // rows [base/32, ...) that lie fully inside both the chunk and the cascade range
            const int64_t r0 = base >> 5;
            const int64_t r1 = min(size_ilp, (base + m) >> 5);
            const int64_t full_limit = (size_ilp / level_step) * level_step;
            const int64_t rf = min(r1, full_limit);
            int64_t i = r0;
            while (i + level_step <= rf) {  // chunk boundaries are multiples of level_step rows
                for (int64_t j = 0; j < level_step; j += 16, i += 16) {  // level_step is a multiple of 16
                    float v[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) v[u] = b[((i + u) << 5) - base + lane];
#pragma unroll
                    for (int u = 0; u < 16; ++u) acc.a0 = __fadd_rn(acc.a0, v[u]);
                }
                acc.a1 = __fadd_rn(acc.a1, acc.a0); acc.a0 = 0.f;
                if ((i & (level_mask << level_power)) == 0) {
                    acc.a2 = __fadd_rn(acc.a2, acc.a1); acc.a1 = 0.f;
                    if ((i & (level_mask << (2 * level_power))) == 0) { acc.a3 = __fadd_rn(acc.a3, acc.a2); acc.a2 = 0.f; }
                }
            }
            for (; i < r1; ++i) acc.a0 = __fadd_rn(acc.a0, b[(i << 5) - base + lane]);
        }
        __syncthreads();
    }
    if (warp < 2 && src[warp]) {
        float total;
        if (small) {
            total = fmb::aten_row_sum_warp(src[warp], n);
        } else {
            float a = __fadd_rn(__fadd_rn(__fadd_rn(acc.a0, acc.a1), acc.a2), acc.a3);
            const float* x = src[warp];
            if (lane < 8)
                for (int64_t v = size_ilp << 2; v < vec_size; ++v) a = __fadd_rn(a, x[v * 8 + lane]);
            const float t1 = __shfl_down_sync(0xffffffffu, a, 8);
            const float t2 = __shfl_down_sync(0xffffffffu, a, 16);
            const float t3 = __shfl_down_sync(0xffffffffu, a, 24);
            const float folded = __fadd_rn(__fadd_rn(__fadd_rn(a, t1), t2), t3);
            float fin = 0.f;
            for (int64_t q = vec_size << 3; q < n; ++q) fin = __fadd_rn(fin, x[q]);
#pragma unroll
            for (int l = 0; l < 8; ++l) fin = __fadd_rn(fin, __shfl_sync(0xffffffffu, folded, l));
            total = fin;
        }
        if (lane == 0) {
            if (warp == 0) bias[0] = fmb::apply_update(bias[0], total, lr, mode);
            else loss_out[0] = __fdiv_rn(total, (float)B);
        }
    }
}

}  // namespace

// BCEWithLogits value and gradient per sample.  kind 0: loss(z); kind 1: loss(sigmoid(z)).
// delta[B] = dLoss/dz (mean reduction folded in); lossv[B] (nullable) = per-sample loss.
FMB_API int fmb_loss_delta(int kind, const float* z, const float* y, int B, float* delta, float* lossv,
                           cudaStream_t stream) {
    FMB_CHECK_ARG(z && y && delta && B > 0, "fmb_loss_delta: bad arguments");
    FMB_CHECK_ARG(kind == 0 || kind == 1, "fmb_loss_delta: unknown loss kind %d", kind);
    loss_delta_kernel<<<(B + 255) / 256, 256, 0, stream>>>(kind, z, y, B, delta, lossv);
    FMB_CHECK_LAUNCH("loss_delta_kernel");
    return FMB_OK;
}

// out[0] = sum(x[0:n]) in ATen's CPU order (what torch.sum / .mean() / the bias gradient use)
FMB_API int fmb_sum_aten(const float* x, int64_t n, float* out, cudaStream_t stream) {
    FMB_CHECK_ARG(x && out && n > 0, "fmb_sum_aten: bad arguments");
    sum_aten_kernel<<<1, 32, 0, stream>>>(x, n, out);
    FMB_CHECK_LAUNCH("sum_aten_kernel");
    return FMB_OK;
}

// p[i] <- update(p[i], g[i]) for dense parameters (bias, MLP weights); mode as in fmb_fm_backward_update
FMB_API int fmb_update_dense(float* p, const float* g, int64_t n, float lr, int mode, cudaStream_t stream) {
    FMB_CHECK_ARG(p && g && n > 0, "fmb_update_dense: bad arguments");
    FMB_CHECK_ARG(mode == 0 || mode == 1, "fmb_update_dense: unknown update mode %d", mode);
    update_dense_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, g, n, lr, mode);
    FMB_CHECK_LAUNCH("update_dense_kernel");
    return FMB_OK;
}

// end of a training step: bias update from sum(delta) (bias nullable) and mean loss (loss_out nullable)
FMB_API int fmb_finish_step(const float* delta, const float* lossv, int B, float* bias, float lr, int mode,
                            float* loss_out, cudaStream_t stream) {
    FMB_CHECK_ARG(delta && B > 0, "fmb_finish_step: bad arguments");
    finish_step_kernel<<<1, 1024, 0, stream>>>(delta, lossv, B, bias, lr, mode, loss_out);
    FMB_CHECK_LAUNCH("finish_step_kernel");
    return FMB_OK;
}
