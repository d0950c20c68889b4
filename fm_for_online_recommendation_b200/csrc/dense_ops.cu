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

// warp 0: bias -= step(sum(delta));  warp 1: loss_out = sum(lossv) / B
__global__ void __launch_bounds__(64) finish_step_kernel(const float* __restrict__ delta,
                                                         const float* __restrict__ lossv, int B, float* bias,
                                                         float lr, int mode, float* loss_out) {
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        if (bias) {
            const float g = fmb::aten_row_sum_warp(delta, B);
            if (threadIdx.x == 0) bias[0] = fmb::apply_update(bias[0], g, lr, mode);
        }
    } else if (loss_out && lossv) {
        const float s = fmb::aten_row_sum_warp(lossv, B);
        if ((threadIdx.x & 31) == 0) loss_out[0] = __fdiv_rn(s, (float)B);
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
    finish_step_kernel<<<1, 64, 0, stream>>>(delta, lossv, B, bias, lr, mode, loss_out);
    FMB_CHECK_LAUNCH("finish_step_kernel");
    return FMB_OK;
}
