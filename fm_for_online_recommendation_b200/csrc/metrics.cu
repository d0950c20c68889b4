// metrics.cu -- evaluation metrics on the device: the running curves of utils/metric_manager.py:7-29,
// the confusion counts of run_experiment (models/models_online_deep/fm_adam.py:101-116) for a batch of predictions, and
// an exact ROC AUC.  Everything a parity check needs (AUC / RMSE to 4 decimals) without a host round trip per sample.
#include "fmb_common.cuh"

namespace {

// regression_metric (metric_manager.py:7-15): out[0] = inf, out[i+1] = (1/(i+1)) * sum_{j<=i} (pred_j - real_j)^2 with
// the sum accumulated sequentially in fp64, like the reference's Python loop.  One warp: lanes stage 32 squared errors,
// lane 0 carries the chain.
__global__ void regression_metric_kernel(const double* __restrict__ pred, const double* __restrict__ real, int64_t n,
                                         double* __restrict__ out) {
    const int lane = threadIdx.x;
    if (lane == 0) out[0] = __longlong_as_double(0x7ff0000000000000ll);
    double val = 0.0;
    for (int64_t base = 0; base < n; base += 32) {
        const int64_t i = base + lane;
        double sq = 0.0;
        if (i < n) { const double d = __dsub_rn(pred[i], real[i]); sq = __dmul_rn(d, d); }
        double mine = 0.0;
        for (int l = 0; l < 32; ++l) {
            const double s = __shfl_sync(0xffffffffu, sq, l);
            if (base + l < n) {
                val = __dadd_rn(val, s);
                if (l == lane) mine = val;
            }
        }
        if (i < n) out[i + 1] = __dmul_rn(__ddiv_rn(1.0, (double)(i + 1)), mine);
    }
}

// classfication_metric (metric_manager.py:18-29): metric[i] = (1/(i+1)) * log(1 + exp(-pred_i*real_i)) (not cumulative in
// the reference), acc[i] = (1/(i+1)) * #{j <= i : pred_j == real_j}.  The hit count is an integer prefix sum (exact).
__global__ void classification_metric_kernel(const double* __restrict__ pred, const double* __restrict__ real, int64_t n,
                                             double* __restrict__ metric, double* __restrict__ acc) {
    const int lane = threadIdx.x;
    long long hits = 0;
    for (int64_t base = 0; base < n; base += 32) {
        const int64_t i = base + lane;
        const bool valid = i < n;
        const double p = valid ? pred[i] : 0.0, r = valid ? real[i] : 1.0;
        const unsigned m = __ballot_sync(0xffffffffu, valid && p == r);
        if (valid) {
            const double inv = __ddiv_rn(1.0, (double)(i + 1));
            const long long h = hits + __popc(m & ((2u << lane) - 1u));
            acc[i] = __dmul_rn(inv, (double)h);
            metric[i] = __dmul_rn(inv, log(1.0 + exp(-p * r)));
        }
        hits += __popc(m);
    }
}

// confusion counts of a batch of predictions, fm_adam.py:101-111: out = {tp, fp, tn, fn}
__global__ void confusion_kernel(const uint8_t* __restrict__ pred, const float* __restrict__ y, int64_t n,
                                 unsigned long long* __restrict__ out) {
    __shared__ unsigned long long c[4];
    if (threadIdx.x < 4) c[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float yy = y[i];
        const bool pos = yy == 1.0f, hit = (pred[i] ? 1.0f : 0.0f) == yy;
        atomicAdd(&c[hit ? (pos ? 0 : 2) : (pos ? 3 : 1)], 1ull);   // integer counts: order-independent
    }
    __syncthreads();
    if (threadIdx.x < 4 && c[threadIdx.x]) atomicAdd(out + threadIdx.x, c[threadIdx.x]);
}

// exact AUC by counting pairs: out[0] += #{(p, q): y_p = 1, y_q = 0, s_p > s_q}, out[1] += #{... s_p == s_q},
// out[2] = #positives, out[3] = #negatives.  O(n_pos * n_neg) integer work, tiled through shared memory.
__global__ void __launch_bounds__(256) auc_pairs_kernel(const float* __restrict__ s, const float* __restrict__ y, int64_t n,
                                                        unsigned long long* __restrict__ out) {
    __shared__ float ts[1024];
    __shared__ unsigned char ty[1024];
    unsigned long long gt = 0, eq = 0, np_ = 0, nn_ = 0;
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool vi = i < n;
    const float si = vi ? s[i] : 0.f;
    const bool pi = vi && y[i] > 0.f;
    if (vi) { if (pi) np_ = 1; else nn_ = 1; }
    for (int64_t base = 0; base < n; base += 1024) {
        for (int t = threadIdx.x; t < 1024; t += 256) {
            const int64_t j = base + t;
            ts[t] = j < n ? s[j] : 0.f;
            ty[t] = j < n ? (y[j] > 0.f ? 1 : 0) : 2;
        }
        __syncthreads();
        if (pi) {
            const int lim = (int)min((int64_t)1024, n - base);
            for (int t = 0; t < lim; ++t)
                if (ty[t] == 0) { gt += si > ts[t]; eq += si == ts[t]; }
        }
        __syncthreads();
    }
    // warp reduce, then one atomic per warp
    for (int o = 16; o; o >>= 1) {
        gt += __shfl_down_sync(0xffffffffu, gt, o); eq += __shfl_down_sync(0xffffffffu, eq, o);
        np_ += __shfl_down_sync(0xffffffffu, np_, o); nn_ += __shfl_down_sync(0xffffffffu, nn_, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (gt) atomicAdd(out, gt);
        if (eq) atomicAdd(out + 1, eq);
        if (np_) atomicAdd(out + 2, np_);
        if (nn_) atomicAdd(out + 3, nn_);
    }
}

__global__ void sigmoid_vec_kernel(const float* __restrict__ z, int n, float* __restrict__ p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = fmb::sigmoid_at(z[i], i, n);
}

}  // namespace

// utils/metric_manager.py:7-15 -- out_dev [n+1] fp64
FMB_API int fmb_metric_regression(const double* pred, const double* real, int64_t n, double* out, cudaStream_t stream) {
    FMB_CHECK_ARG(pred && real && out && n > 0, "fmb_metric_regression: bad arguments");
    regression_metric_kernel<<<1, 32, 0, stream>>>(pred, real, n, out);
    FMB_CHECK_LAUNCH("regression_metric_kernel");
    return FMB_OK;
}

// utils/metric_manager.py:18-29 -- metric_dev [n], acc_dev [n] fp64
FMB_API int fmb_metric_classification(const double* pred, const double* real, int64_t n, double* metric, double* acc,
                                      cudaStream_t stream) {
    FMB_CHECK_ARG(pred && real && metric && acc && n > 0, "fmb_metric_classification: bad arguments");
    classification_metric_kernel<<<1, 32, 0, stream>>>(pred, real, n, metric, acc);
    FMB_CHECK_LAUNCH("classification_metric_kernel");
    return FMB_OK;
}

// fm_adam.py:101-111 on a batch: conf_dev [4] uint64 = {tp, fp, tn, fn} (accumulated: zero it first)
FMB_API int fmb_confusion(const uint8_t* pred, const float* y, int64_t n, unsigned long long* conf, cudaStream_t stream) {
    FMB_CHECK_ARG(pred && y && conf && n > 0, "fmb_confusion: bad arguments");
    confusion_kernel<<<(unsigned)min((int64_t)592, (n + 255) / 256), 256, 0, stream>>>(pred, y, n, conf);
    FMB_CHECK_LAUNCH("confusion_kernel");
    return FMB_OK;
}

// exact ROC AUC ingredients: counts_dev [4] uint64 += {#(pos > neg), #(pos == neg), #pos, #neg} (zero it first);
// AUC = (counts[0] + counts[1] / 2) / (counts[2] * counts[3]).  n <= 2^20 (pair counting, O(n_pos * n_neg)).
FMB_API int fmb_auc_pairs(const float* scores, const float* labels, int64_t n, unsigned long long* counts,
                          cudaStream_t stream) {
    FMB_CHECK_ARG(scores && labels && counts && n > 0 && n <= (1 << 20), "fmb_auc_pairs: bad arguments (n <= 2^20)");
    auc_pairs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(scores, labels, n, counts);
    FMB_CHECK_LAUNCH("auc_pairs_kernel");
    return FMB_OK;
}

// torch.sigmoid of a contiguous [n] logit vector (predict_proba): ATen's bits (fmb_aten_math.cuh)
FMB_API int fmb_sigmoid(const float* z, int n, float* p, cudaStream_t stream) {
    FMB_CHECK_ARG(z && p && n > 0, "fmb_sigmoid: bad arguments");
    sigmoid_vec_kernel<<<(n + 255) / 256, 256, 0, stream>>>(z, n, p);
    FMB_CHECK_LAUNCH("sigmoid_vec_kernel");
    return FMB_OK;
}
