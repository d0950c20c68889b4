// heads.cu -- logit assembly, ONN per-layer heads, predict thresholds and the hedge-backprop
// bookkeeping (models/models_online_deep/deepfm_adam.py:88, nfm_adam.py:79-87,
// deepfm_onn.py:88-102,109-154,171-175).
#include "fmb_common.cuh"

namespace {

__device__ __forceinline__ float base_of(int nfm, const float* z_fm, const float* sum_first, const float* bias, int b) {
    // NFM: fm_first = sum(first_order, 1) + bias (nfm_adam.py:79); DeepFM: fm_part = forward_fm (deepfm_adam.py:80)
    return nfm ? __fadd_rn(sum_first[b], bias[0]) : z_fm[b];
}

__global__ void combine_logit_kernel(int nfm, const float* z_fm, const float* sum_first, const float* bias,
                                     const float* head, int B, float* z) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) z[b] = __fadd_rn(base_of(nfm, z_fm, sum_first, bias, b), head[b]);
}

__global__ void onn_heads_kernel(int nfm, const float* z_fm, const float* sum_first, const float* bias,
                                 const float* head, int L, int B, float* p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L * B) return;
    const int b = i % B;
    p[i] = fmb::sigmoid_at(__fadd_rn(base_of(nfm, z_fm, sum_first, bias, b), head[i]), b, B);   // one torch.sigmoid per layer on [B]
}

__global__ void predict_kernel(const float* z, int n, uint8_t* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fmb::sigmoid_at(z[i], i, n) > 0.5f;
}

// nn.BCELoss value and d(loss)/d(pre-sigmoid logit) of one head (deepfm_onn.py:117-120,127)
__global__ void hedge_head_grad_kernel(const float* p, const float* y, int B, float* gtop, float* lossv) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float pr = p[b], yy = y[b];
    float l1 = fmaxf(fmb::log1pf_p(-pr), -100.f);
    float l0 = fmaxf(fmb::logf_p(pr), -100.f);
    lossv[b] = __fsub_rn(__fmul_rn(__fsub_rn(yy, 1.0f), l1), __fmul_rn(yy, l0));
    const float invB = __fdiv_rn(1.0f, (float)B);
    const float den = fmaxf(__fmul_rn(__fsub_rn(1.0f, pr), pr), 1e-12f);
    const float dp = __fdiv_rn(__fmul_rn(invB, __fsub_rn(pr, yy)), den);
    gtop[b] = __fmul_rn(__fmul_rn(dp, __fsub_rn(1.0f, pr)), pr);
}

// acc[t] (+)= alpha[i] * g[t] for the parameters of layers 0..i (deepfm_onn.py:132-139)
__global__ void hedge_accumulate_kernel(float* acc, const float* g, const float* alpha, int i, int64_t lo_i,
                                        int64_t hi_i) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= hi_i) return;
    const float term = __fmul_rn(alpha[i], g[t]);
    acc[t] = (t >= lo_i) ? term : __fadd_rn(acc[t], term);  // layer i itself is seen for the first time
}

// W -= n * acc (deepfm_onn.py:143-145): elementwise over the whole tower (325 200 parameters at cfg4: a grid, not one CTA)
__global__ void hedge_apply_w_kernel(float* mlp, const float* acc, int64_t n, float lr) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) mlp[t] = __fsub_rn(mlp[t], __fmul_rn(lr, acc[t]));
}
// the alpha update (deepfm_onn.py:147-154). One thread.
__global__ void hedge_alpha_kernel(float* alpha, const float* loss_sum, int L, int B, float hb, float hs) {
    if (threadIdx.x == 0) {
        const float floorv = __fdiv_rn(hs, (float)L);
        for (int i = 0; i < L; ++i) {
            const float loss = __fdiv_rn(loss_sum[i], (float)B);
            const float a = __fmul_rn(alpha[i], fmb::powf_p(hb, loss));
            alpha[i] = fmaxf(a, floorv);
        }
        const float zt = fmb::aten_row_sum_small([&](int j) { return alpha[j]; }, L);
        for (int i = 0; i < L; ++i) alpha[i] = __fdiv_rn(alpha[i], zt);
    }
}

}  // namespace

// z[b] = base[b] + head[b]; base = z_fm (DeepFM) or sum_first + bias (NFM, nfm=1)
FMB_API int fmb_combine_logit(int nfm, const float* z_fm, const float* sum_first, const float* bias,
                              const float* head, int B, float* z, cudaStream_t stream) {
    FMB_CHECK_ARG(z_fm && sum_first && bias && head && z && B > 0, "fmb_combine_logit: bad arguments");
    combine_logit_kernel<<<(B + 255) / 256, 256, 0, stream>>>(nfm, z_fm, sum_first, bias, head, B, z);
    FMB_CHECK_LAUNCH("combine_logit_kernel");
    return FMB_OK;
}

// p[l,b] = sigmoid(base[b] + head[l,b])   (deepfm_onn.py:95-99)
FMB_API int fmb_onn_heads(int nfm, const float* z_fm, const float* sum_first, const float* bias, const float* head,
                          int L, int B, float* p, cudaStream_t stream) {
    FMB_CHECK_ARG(z_fm && sum_first && bias && head && p && B > 0 && L > 0, "fmb_onn_heads: bad arguments");
    onn_heads_kernel<<<(L * B + 255) / 256, 256, 0, stream>>>(nfm, z_fm, sum_first, bias, head, L, B, p);
    FMB_CHECK_LAUNCH("onn_heads_kernel");
    return FMB_OK;
}

// out[i] = sigmoid(z[i]) > 0.5   (fm_adam.py:87-88; deepfm_onn.py:174-175 applies it to p_last)
FMB_API int fmb_predict(const float* z, int n, uint8_t* out, cudaStream_t stream) {
    FMB_CHECK_ARG(z && out && n > 0, "fmb_predict: bad arguments");
    predict_kernel<<<(n + 255) / 256, 256, 0, stream>>>(z, n, out);
    FMB_CHECK_LAUNCH("predict_kernel");
    return FMB_OK;
}

FMB_API int fmb_hedge_head_grad(const float* p, const float* y, int B, float* gtop, float* lossv,
                                cudaStream_t stream) {
    FMB_CHECK_ARG(p && y && gtop && lossv && B > 0, "fmb_hedge_head_grad: bad arguments");
    hedge_head_grad_kernel<<<(B + 255) / 256, 256, 0, stream>>>(p, y, B, gtop, lossv);
    FMB_CHECK_LAUNCH("hedge_head_grad_kernel");
    return FMB_OK;
}

static size_t w_off(int k, int H, int l) { return l == 0 ? 0 : (size_t)H * k + H + (size_t)(l - 1) * ((size_t)H * H + H); }

FMB_API int fmb_hedge_accumulate(float* acc, const float* gmlp, const float* alpha, int i, int k, int L, int H,
                                 cudaStream_t stream) {
    FMB_CHECK_ARG(acc && gmlp && alpha && i >= 0 && i < L, "fmb_hedge_accumulate: bad arguments");
    const int64_t lo = (int64_t)w_off(k, H, i), hi = (int64_t)w_off(k, H, i + 1);
    hedge_accumulate_kernel<<<(unsigned)((hi + 255) / 256), 256, 0, stream>>>(acc, gmlp, alpha, i, lo, hi);
    FMB_CHECK_LAUNCH("hedge_accumulate_kernel");
    return FMB_OK;
}

FMB_API int fmb_hedge_apply(float* mlp, const float* acc, float lr, float* alpha, const float* loss_sum, int B,
                            int k, int L, int H, float hb, float hs, cudaStream_t stream) {
    FMB_CHECK_ARG(mlp && acc && alpha && loss_sum && L > 0 && L < 512, "fmb_hedge_apply: bad arguments");
    const int64_t n = (int64_t)w_off(k, H, L);
    hedge_apply_w_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(mlp, acc, n, lr);
    hedge_alpha_kernel<<<1, 32, 0, stream>>>(alpha, loss_sum, L, B, hb, hs);
    FMB_CHECK_LAUNCH("hedge_apply_kernel");
    return FMB_OK;
}
