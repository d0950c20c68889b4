// classical.cu -- the fp64 per-example online learners as persistent single-CTA kernels.
//
// Replaces the Python loops of models/models_online/FM_FTRL.py:47-92 (online FM by linearised FTRL),
// SFTRL_CCFM.py:30-121 and SFTRL_Vanila.py:33-123 (sketched FTRL with two Frequent-Directions
// sketches); SURVEY.md 8a rows A9-A11.  The stream is a strict dependency chain (sample i+1 reads
// the state sample i wrote), so one CTA walks it in order and the whole run is ONE launch: no host
// round trip per sample, state stays in L2/shared memory.  Latency-bound by construction; reported
// as samples/s, no roofline fraction claimed.
//
// Only the non-zero features of a sample are touched (ml-100k rows have 3 non-zeros out of 2626):
// adding exact zeros does not change an IEEE sum, so this equals the reference's dense algebra.
#include "fmb_common.cuh"
#include <math_constants.h>

namespace {

constexpr int CT = 256;  // threads of the persistent CTA

// non-zero entries of row x[0:d) -> (nz_idx, nz_val) in ascending index order; returns the count
__device__ int compact_row(const double* __restrict__ x, int d, int* nz_idx, double* nz_val, int* s_cnt) {
    __shared__ int wcnt[CT / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) *s_cnt = 0;
    __syncthreads();
    for (int base = 0; base < d; base += CT) {
        const int j = base + threadIdx.x;
        const double v = j < d ? x[j] : 0.0;
        const bool nzf = j < d && v != 0.0;
        const unsigned bal = __ballot_sync(0xffffffffu, nzf);
        if (lane == 0) wcnt[warp] = __popc(bal);
        __syncthreads();
        int pre = *s_cnt, tot = 0;
        for (int w = 0; w < CT / 32; ++w) { if (w < warp) pre += wcnt[w]; tot += wcnt[w]; }
        if (nzf) { const int o = pre + __popc(bal & ((1u << lane) - 1u)); nz_idx[o] = j; nz_val[o] = v; }
        __syncthreads();
        if (threadIdx.x == 0) *s_cnt += tot;
        __syncthreads();
    }
    return *s_cnt;
}

// deterministic block sum (fixed tree) of one double per thread; result to all threads
__device__ double block_sum(double v, double* red) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int s = CT / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
}

__device__ __forceinline__ double grad_loss(int task, double scalar, double y) {
    // FM_FTRL.py:68-73 with FM_Base._grad_loss (FM_Base.py:44-51)
    return task == 1 ? (-1.0 / (1.0 + exp(scalar * y))) * y : 2.0 * (scalar - y);
}

// ------------------------------------------------------------------------------------------------
// FM_FTRL (A9)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CT) ftrl_fm_kernel(const double* __restrict__ X, const double* __restrict__ y,
                                                     int N, int d, int m2, int task, double eta, double* w1,
                                                     double* W2, double* g_w1, double* g_W2, double* preds,
                                                     int* status) {
    extern __shared__ __align__(16) unsigned char dsm[];
    double* nz_val = reinterpret_cast<double*>(dsm);             // [d]
    double* t_s = nz_val + d;                                    // [m2]
    double* red = t_s + m2;                                      // [CT]
    int* nz_idx = reinterpret_cast<int*>(red + CT);              // [d]
    __shared__ int s_cnt;
    const int dm1 = d - 1;
    for (int idx = 0; idx < N; ++idx) {
        const int nnz = compact_row(X + (size_t)idx * d, d, nz_idx, nz_val, &s_cnt);
        // t = W2 a[:-1]  (one thread per row of W2), wa = w1^T a
        for (int i = threadIdx.x; i < m2; i += CT) {
            double acc = 0.0;
            for (int e = 0; e < nnz; ++e) { const int j = nz_idx[e]; if (j < dm1) acc += W2[(size_t)i * dm1 + j] * nz_val[e]; }
            t_s[i] = acc;
        }
        double part = 0.0;
        for (int e = threadIdx.x; e < nnz; e += CT) part += w1[nz_idx[e]] * nz_val[e];
        const double wa = block_sum(part, red);
        part = 0.0;
        for (int i = threadIdx.x; i < m2; i += CT) part += t_s[i] * t_s[i];
        const double tt = block_sum(part, red);
        const double scalar = wa + tt;
        if (scalar != scalar) { if (threadIdx.x == 0) *status = idx + 1; return; }   // ValueError('Nan contained')
        const double yy = y[idx];
        if (threadIdx.x == 0) preds[idx] = task == 1 ? (scalar >= 0 ? 1.0 : -1.0) : scalar;
        const double sign = grad_loss(task, scalar, yy);
        // g_w1 += sign*a ; g_W2 += (2 t) a'^T ; w = -eta g   (FM_FTRL.py:76-80)
        for (int e = threadIdx.x; e < nnz; e += CT) {
            const int j = nz_idx[e];
            const double g = g_w1[j] + sign * nz_val[e];
            g_w1[j] = g;
            w1[j] = -eta * g;
        }
        for (int it = threadIdx.x; it < m2 * nnz; it += CT) {
            const int i = it / nnz, e = it - i * nnz, j = nz_idx[e];
            if (j < dm1) {
                const double g = g_W2[(size_t)i * dm1 + j] + (2.0 * t_s[i]) * nz_val[e];
                g_W2[(size_t)i * dm1 + j] = g;
                W2[(size_t)i * dm1 + j] = -eta * g;
            }
        }
        if (idx == 0) {
            // the first assignment `self.w1 = -eta*g_w1` replaces the WHOLE randn-initialised tensors
            __syncthreads();
            for (int j = threadIdx.x; j < d; j += CT) w1[j] = -eta * g_w1[j];
            for (size_t q = threadIdx.x; q < (size_t)m2 * dm1; q += CT) W2[q] = -eta * g_W2[q];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// symmetric eigen-decomposition by cyclic Jacobi (n <= 128), A and V in shared memory, one CTA.
// On return A's diagonal holds the eigenvalues, V's columns the eigenvectors.
// ------------------------------------------------------------------------------------------------
__device__ void jacobi_eig(double* A, double* V, int n, double* cs /*[2*n]*/, int* pairs /*[2*n]*/, double* red) {
    for (int q = threadIdx.x; q < n * n; q += CT) V[q] = (q / n == q % n) ? 1.0 : 0.0;
    __syncthreads();
    const int ne = n + (n & 1);  // round-robin tournament needs an even number of players
    const int half = ne / 2;
    for (int sweep = 0; sweep < 30; ++sweep) {
        // off-diagonal norm
        double part = 0.0;
        for (int q = threadIdx.x; q < n * n; q += CT) { const int i = q / n, j = q % n; if (i != j) part += A[q] * A[q]; }
        const double off = block_sum(part, red);
        part = 0.0;
        for (int i = threadIdx.x; i < n; i += CT) part += A[i * n + i] * A[i * n + i];
        const double diag = block_sum(part, red);
        if (off <= 1e-30 * (diag + 1e-300) || off == 0.0) break;
        for (int step = 0; step < ne - 1; ++step) {
            // pairing of the circle method: player ne-1 fixed, the others rotate
            if (threadIdx.x < half) {
                const int t = threadIdx.x;
                int a = (t == 0) ? ne - 1 : (step + t) % (ne - 1);
                int b = (step + ne - 1 - t) % (ne - 1);
                int p = min(a, b), q = max(a, b);
                double c = 1.0, s = 0.0;
                if (q < n) {
                    const double apq = A[p * n + q];
                    if (apq != 0.0) {
                        const double tau = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                        const double tt = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + tt * tt);
                        s = tt * c;
                    }
                } else { p = q = -1; }
                pairs[2 * t] = p; pairs[2 * t + 1] = q; cs[2 * t] = c; cs[2 * t + 1] = s;
            }
            __syncthreads();
            // columns: A <- A J, V <- V J
            for (int it = threadIdx.x; it < half * n; it += CT) {
                const int t = it / n, i = it - t * n;
                const int p = pairs[2 * t], q = pairs[2 * t + 1];
                if (p < 0) continue;
                const double c = cs[2 * t], s = cs[2 * t + 1];
                const double aip = A[i * n + p], aiq = A[i * n + q];
                A[i * n + p] = c * aip - s * aiq;
                A[i * n + q] = s * aip + c * aiq;
                const double vip = V[i * n + p], viq = V[i * n + q];
                V[i * n + p] = c * vip - s * viq;
                V[i * n + q] = s * vip + c * viq;
            }
            __syncthreads();
            // rows: A <- J^T A
            for (int it = threadIdx.x; it < half * n; it += CT) {
                const int t = it / n, j = it - t * n;
                const int p = pairs[2 * t], q = pairs[2 * t + 1];
                if (p < 0) continue;
                const double c = cs[2 * t], s = cs[2 * t + 1];
                const double apj = A[p * n + j], aqj = A[q * n + j];
                A[p * n + j] = c * apj - s * aqj;
                A[q * n + j] = s * apj + c * aqj;
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// SFTRL_CCFM / SFTRL_Vanila (A10, A11)
// ------------------------------------------------------------------------------------------------
struct SftrlParams {
    const double* X; const double* y;
    int N, d, ds, m, task, vanila;
    double eta;
    double* BT[2];      // [ds][2m] row-major: 0 = BT_P, 1 = BT_N
    int* rc;            // [2] row_count_p, row_count_n
    double* w; double* g_w;   // [d] (vanila)
    double* tmp;        // [ds*2m] scratch for the rebuilt sketch
    double* preds;
    int* status;
    int ne;             // eigenproblem size = min(ds, 2m)
};

// Frequent-Directions shrink of sketch BT (SFTRL_CCFM.py:85-99), using the smaller of the two Gram
// matrices: eigenvectors of BT^T BT (2m x 2m) as the reference does, or, when ds < 2m, of
// BT BT^T (ds x ds), whose unit eigenvectors ARE the columns of V = BT U Sigma^-1/2 (up to sign).
__device__ int gfd_shrink(const SftrlParams& p, double* BT, double* A, double* V, double* lam, int* order,
                          double* cs, int* pairs, double* red) {
    const int m2 = 2 * p.m, ds = p.ds, n = p.ne;
    const bool small_side = ds < m2;
    // Gram matrix
    for (int q = threadIdx.x; q < n * n; q += CT) {
        const int r = q / n, c = q % n;
        double acc = 0.0;
        if (!small_side) { for (int j = 0; j < ds; ++j) acc += BT[(size_t)j * m2 + r] * BT[(size_t)j * m2 + c]; }
        else { for (int j = 0; j < m2; ++j) acc += BT[(size_t)r * m2 + j] * BT[(size_t)c * m2 + j]; }
        A[q] = acc;
    }
    __syncthreads();
    jacobi_eig(A, V, n, cs, pairs, red);
    // singular values of a symmetric PSD matrix = |eigenvalues|, sorted descending; <= 1e-12 -> 0
    if (threadIdx.x == 0) {
        for (int i = 0; i < n; ++i) { const double l = fabs(A[i * n + i]); lam[i] = l <= 1e-12 ? 0.0 : l; order[i] = i; }
        for (int i = 1; i < n; ++i) {  // insertion sort, descending, stable
            const int oi = order[i]; const double li = lam[oi];
            int j = i - 1;
            while (j >= 0 && lam[order[j]] < li) { order[j + 1] = order[j]; --j; }
            order[j + 1] = oi;
        }
    }
    __syncthreads();
    int nnz = 0;
    for (int i = 0; i < n; ++i) if (lam[order[i]] != 0.0) ++nnz;   // uniform across threads
    const bool shrink = nnz >= p.m;
    const int keep = shrink ? p.m - 1 : nnz;
    const double cut = (shrink && p.m < n) ? lam[order[p.m]] : 0.0;  // Sigma[m] (0 when the Gram has <= m values)
    // new BT[:, c] = V[:, c] * sqrt(Sigma_c - cut), V[:, c] = BT u_c / sqrt(Sigma_c)   (or w_c directly)
    for (size_t q = threadIdx.x; q < (size_t)ds * m2; q += CT) p.tmp[q] = 0.0;
    __syncthreads();
    for (int it = threadIdx.x; it < ds * keep; it += CT) {
        const int j = it / keep, c = it - j * keep;
        const int ec = order[c];
        const double sg = lam[ec];
        double vjc;
        if (!small_side) {
            double acc = 0.0;
            for (int r = 0; r < m2; ++r) acc += BT[(size_t)j * m2 + r] * V[r * n + ec];
            vjc = acc * (1.0 / sqrt(sg));
        } else {
            vjc = V[j * n + ec];
        }
        p.tmp[(size_t)j * m2 + c] = vjc * sqrt(sg - cut);
    }
    __syncthreads();
    for (size_t q = threadIdx.x; q < (size_t)ds * m2; q += CT) BT[q] = p.tmp[q];
    __syncthreads();
    return keep;
}

__global__ void __launch_bounds__(CT) sftrl_kernel(SftrlParams p) {
    extern __shared__ __align__(16) unsigned char dsm[];
    const int m2 = 2 * p.m, d = p.d, n = p.ne;
    double* nz_val = reinterpret_cast<double*>(dsm);   // [d]
    double* A = nz_val + d;                             // [n*n]
    double* V = A + n * n;                              // [n*n]
    double* lam = V + n * n;                            // [n]
    double* cs = lam + n;                               // [2n+2]
    double* red = cs + 2 * n + 2;                       // [CT]
    double* proj = red + CT;                            // [2*m2]  BP_alpha, BN_alpha
    int* nz_idx = reinterpret_cast<int*>(proj + 2 * m2);   // [d]
    int* order = nz_idx + d;                            // [n]
    int* pairs = order + n;                             // [2n+2]
    __shared__ int s_cnt;
    int rc[2] = {p.rc[0], p.rc[1]};
    for (int idx = 0; idx < p.N; ++idx) {
        const int nnz = compact_row(p.X + (size_t)idx * d, d, nz_idx, nz_val, &s_cnt);
        // BP_alpha = BT_P^T a', BN_alpha = BT_N^T a' (columns beyond row_count are zero)
        for (int it = threadIdx.x; it < 2 * m2; it += CT) {
            const int which = it / m2, c = it - which * m2;
            double acc = 0.0;
            if (c <= rc[which]) {
                const double* B = p.BT[which];
                for (int e = 0; e < nnz; ++e) { const int j = nz_idx[e]; if (j < p.ds) acc += B[(size_t)j * m2 + c] * nz_val[e]; }
            }
            proj[it] = acc;
        }
        __syncthreads();
        double part = 0.0;
        for (int c = threadIdx.x; c < m2; c += CT) part += proj[c] * proj[c];
        const double pp = block_sum(part, red);
        part = 0.0;
        for (int c = threadIdx.x; c < m2; c += CT) part += proj[m2 + c] * proj[m2 + c];
        const double nn = block_sum(part, red);
        double scalar = pp - nn;
        if (p.vanila) {
            part = 0.0;
            for (int e = threadIdx.x; e < nnz; e += CT) part += p.w[nz_idx[e]] * nz_val[e];
            const double wa = block_sum(part, red);
            scalar = (wa + pp) - nn;   // SFTRL_Vanila.py:46
        }
        if (scalar != scalar) { if (threadIdx.x == 0) *p.status = idx + 1; return; }
        const double yy = p.y[idx];
        if (threadIdx.x == 0) p.preds[idx] = p.task == 1 ? (scalar >= 0 ? 1.0 : -1.0) : scalar;
        const double sign = grad_loss(p.task, scalar, yy);
        if (p.vanila) {  // SFTRL_Vanila.py:59-60
            for (int e = threadIdx.x; e < nnz; e += CT) {
                const int j = nz_idx[e];
                const double g = p.g_w[j] + sign * nz_val[e];
                p.g_w[j] = g;
                p.w[j] = -p.eta * g;
            }
        }
        // _GFD (SFTRL_CCFM.py:77-121): the count is incremented BEFORE the insert (column 0 stays empty)
        const int which = sign <= 0 ? 0 : 1;
        const double scale = sqrt(which == 0 ? -p.eta * sign : p.eta * sign);
        rc[which] += 1;
        double* B = p.BT[which];
        for (int e = threadIdx.x; e < nnz; e += CT) { const int j = nz_idx[e]; if (j < p.ds) B[(size_t)j * m2 + rc[which]] = scale * nz_val[e]; }
        __syncthreads();
        if (rc[which] == m2 - 1) rc[which] = gfd_shrink(p, B, A, V, lam, order, cs, pairs, red);
        __syncthreads();
    }
    if (threadIdx.x == 0) { p.rc[0] = rc[0]; p.rc[1] = rc[1]; }
}

}  // namespace

// A9: FM_FTRL.online_learning (FM_FTRL.py:47-92) over the whole stream in one launch.
//   X [N,d] fp64 dense, y [N]; task 0 = 'reg', 1 = 'cls'; m2 = 2*m rows of W2
//   w1 [d], W2 [m2,d-1]: randn initial values in, final values out; g_w1 [d], g_W2 [m2,d-1]: zeroed
//   accumulators (caller-allocated); preds [N]; status [1]: 0 = ok, i+1 = NaN at sample i.
FMB_API int fmb_ftrl_fm_run(const double* X, const double* y, int N, int d, int m2, int task, double eta,
                            double* w1, double* W2, double* g_w1, double* g_W2, double* preds, int* status,
                            cudaStream_t stream) {
    FMB_CHECK_ARG(X && y && w1 && W2 && g_w1 && g_W2 && preds && status, "fmb_ftrl_fm_run: null pointer");
    FMB_CHECK_ARG(N > 0 && d > 1 && m2 > 0 && (task == 0 || task == 1), "fmb_ftrl_fm_run: bad arguments");
    const size_t sm = (size_t)d * 8 + (size_t)m2 * 8 + CT * 8 + (size_t)d * 4 + 16;
    FMB_CHECK_ARG(sm <= 200 * 1024, "fmb_ftrl_fm_run: d=%d too large for one CTA", d);
    cudaFuncSetAttribute(ftrl_fm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaMemsetAsync(status, 0, sizeof(int), stream);
    ftrl_fm_kernel<<<1, CT, sm, stream>>>(X, y, N, d, m2, task, eta, w1, W2, g_w1, g_W2, preds, status);
    FMB_CHECK_LAUNCH("ftrl_fm_kernel");
    return FMB_OK;
}

FMB_API size_t fmb_sftrl_workspace_bytes(int d, int m) { return (size_t)d * 2 * m * sizeof(double) + 256; }

// A10/A11: SFTRL_CCFM.online_learning (vanila = 0) / SFTRL_Vanila.online_learning (vanila = 1).
//   BT_P, BT_N [ds, 2m] row-major (ds = d, or d-1 for vanila), zero-initialised in, final sketches out
//   rc [2] row_count_p / row_count_n in/out; w, g_w [d] (vanila only, zero-initialised)
FMB_API int fmb_sftrl_run(const double* X, const double* y, int N, int d, int m, int task, int vanila, double eta,
                          double* BT_P, double* BT_N, int* rc, double* w, double* g_w, double* preds, int* status,
                          void* ws, size_t ws_bytes, cudaStream_t stream) {
    FMB_CHECK_ARG(X && y && BT_P && BT_N && rc && preds && status && ws, "fmb_sftrl_run: null pointer");
    FMB_CHECK_ARG(N > 0 && d > 1 && m > 1 && (task == 0 || task == 1), "fmb_sftrl_run: bad arguments");
    FMB_CHECK_ARG(!vanila || (w && g_w), "fmb_sftrl_run: vanila needs w and g_w");
    if (ws_bytes < fmb_sftrl_workspace_bytes(d, m)) { fmb_set_error("fmb_sftrl_run: workspace too small"); return FMB_ERR_WS; }
    SftrlParams p;
    p.X = X; p.y = y; p.N = N; p.d = d; p.ds = vanila ? d - 1 : d; p.m = m; p.task = task; p.vanila = vanila;
    p.eta = eta; p.BT[0] = BT_P; p.BT[1] = BT_N; p.rc = rc; p.w = w; p.g_w = g_w; p.tmp = (double*)ws;
    p.preds = preds; p.status = status;
    p.ne = p.ds < 2 * m ? p.ds : 2 * m;
    FMB_CHECK_ARG(p.ne <= 128, "fmb_sftrl_run: min(d, 2m) = %d > 128 not supported", p.ne);
    const int n = p.ne, m2 = 2 * m;
    const size_t sm = (size_t)d * 8 + (size_t)2 * n * n * 8 + (size_t)n * 8 + (size_t)(2 * n + 2) * 8 + CT * 8 +
                      (size_t)2 * m2 * 8 + (size_t)d * 4 + (size_t)n * 4 + (size_t)(2 * n + 2) * 4 + 32;
    FMB_CHECK_ARG(sm <= 220 * 1024, "fmb_sftrl_run: d/m too large for one CTA");
    cudaFuncSetAttribute(sftrl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaMemsetAsync(status, 0, sizeof(int), stream);
    sftrl_kernel<<<1, CT, sm, stream>>>(p);
    FMB_CHECK_LAUNCH("sftrl_kernel");
    return FMB_OK;
}
