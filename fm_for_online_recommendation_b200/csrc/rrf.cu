// rrf.cu -- RRF_Online (models/models_online/RRF_Online.py:70-187): online reparameterised random
// Fourier features, fp64, one persistent CTA walking the stream in order (every sample's prediction uses the weights
// all earlier samples left behind -- the reference's Python loop, never mini-batched).
//
// Per sample x (1 x d), y:   p_j = sum_i x_i * exp(gamma_i) * eps[i][j]             j < D         (:70-75)
//                            phi = [cos p | sin p]   (2D);   s = phi . w                           (:140)
//   NaN score: the sample is skipped (no update, no prediction recorded)                           (:162-165)
//   d_w = lr_w * exp(w) + c * phi,  d_phi = c * w,  c = s - y ('reg', l2) or -y ('cls', logit: with one sample the
//         logsumexp over the batch cancels the exponent exactly, :103-106)                          (:96-113)
//   d_gamma_i = exp(gamma_i) * x_i * sum_j eps[i][j] * (-sin p_j * d_phi_j + cos p_j * d_phi_{D+j}) (:77-87)
//   w -= lr_w * d_w;  gamma -= lr_gamma * d_gamma;  prediction = s ('reg') or +-1 ('cls', s >= 0 -> +1)  (:167-172)
#include "fmb_common.cuh"

namespace {

constexpr int RT = 256;

__global__ void __launch_bounds__(RT) rrf_kernel(const double* __restrict__ X, const double* __restrict__ Y, int N, int d,
                                                 int D, int task, double lr_w, double lr_g, double* __restrict__ gamma,
                                                 double* __restrict__ w, const double* __restrict__ eps,
                                                 double* __restrict__ preds, int* __restrict__ nvalid) {
    extern __shared__ double sm[];
    double* p = sm;             // [D]
    double* cs = p + D;         // [2D] cos | sin
    double* dphi = cs + 2 * D;  // [2D]
    double* red = dphi + 2 * D; // [RT]
    double* ws = red + RT;      // [2D] weights
    __shared__ double score;
    const int tid = threadIdx.x;
    for (int j = tid; j < 2 * D; j += RT) ws[j] = w[j];
    __syncthreads();
    int nv = 0;
    for (int t = 0; t < N; ++t) {
        const double* x = X + (size_t)t * d;
        // p_j = sum_i x_i e^{gamma_i} eps[i][j]: threads over i, block reduction per j
        for (int j = 0; j < D; ++j) {
            double a = 0.0;
            for (int i = tid; i < d; i += RT) {
                const double xi = x[i];
                if (xi != 0.0) a += xi * exp(gamma[i]) * eps[(size_t)i * D + j];
            }
            red[tid] = a;
            __syncthreads();
            for (int o = RT / 2; o > 0; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
            if (tid == 0) p[j] = red[0];
            __syncthreads();
        }
        if (tid < D) { cs[tid] = cos(p[tid]); cs[D + tid] = sin(p[tid]); }
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int j = 0; j < 2 * D; ++j) s += cs[j] * ws[j];
            score = s;
        }
        __syncthreads();
        const double s = score;
        if (s != s) continue;   // uniform: every thread sees the same score
        const double y = Y[t];
        const double c = task == 0 ? (s - y) : -y;
        if (tid < 2 * D) dphi[tid] = c * ws[tid];
        __syncthreads();
        // gamma update (uses the pre-update w through dphi), threads over features
        for (int i = tid; i < d; i += RT) {
            const double xi = x[i];
            if (xi != 0.0) {
                double acc = 0.0;
                for (int j = 0; j < D; ++j) acc += eps[(size_t)i * D + j] * (-cs[D + j] * dphi[j] + cs[j] * dphi[D + j]);
                const double g = gamma[i];
                gamma[i] = g - lr_g * (exp(g) * xi * acc);
            }
        }
        if (tid < 2 * D) {
            const double wj = ws[tid];
            const double dw = lr_w * exp(wj) + c * cs[tid];
            ws[tid] = wj - lr_w * dw;
        }
        if (tid == 0) preds[nv] = task == 0 ? s : (s >= 0.0 ? 1.0 : -1.0);
        ++nv;
        __syncthreads();
    }
    for (int j = tid; j < 2 * D; j += RT) w[j] = ws[j];
    if (tid == 0) *nvalid = nv;
}

}  // namespace

// RRF_Online.online_learning (RRF_Online.py:142-187).  X [N,d], Y [N] fp64; gamma [d] (log scale), w [2D], eps [d,D] are
// the learner's state (gamma and w updated in place); preds [N] receives the predictions of the samples whose score was
// not NaN, *nvalid their number (the reference appends only those).  task 0 = 'reg' (l2), 1 = 'cls' (logit).
FMB_API int fmb_rrf_run(const double* X, const double* Y, int N, int d, int D, int task, double lr_w, double lr_gamma,
                        double* gamma, double* w, const double* eps, double* preds, int* nvalid, cudaStream_t stream) {
    FMB_CHECK_ARG(X && Y && gamma && w && eps && preds && nvalid, "fmb_rrf_run: null pointer");
    FMB_CHECK_ARG(N > 0 && d > 0 && D > 0 && D <= 128 && (task == 0 || task == 1), "fmb_rrf_run: bad arguments");
    const size_t smem = (size_t)(7 * D + RT) * sizeof(double);
    rrf_kernel<<<1, RT, smem, stream>>>(X, Y, N, d, D, task, lr_w, lr_gamma, gamma, w, eps, preds, nvalid);
    FMB_CHECK_LAUNCH("rrf_kernel");
    return FMB_OK;
}
