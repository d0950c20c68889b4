// dataset.cu -- the device-resident input pipeline: the data set is encoded ONCE into device arrays
// (global row ids int32 [N,F], values fp32 [N,F] or none, labels fp32 [N]) and every batch the reference's builders would
// assemble as Python lists of lists (utils/data_preprocess.py:154-264: _construct_batch_criteo_data, create_ten_iter,
// create_dataset; :46-82 / :120-151 the balance_* functions) is either a view of those arrays or ONE row-gather launch.
// read_svm_file's vocabulary build (utils/data_preprocess.py:100-108: per column, `list.index(value)` with append on a
// miss -- O(N * vocabulary) in Python) becomes a first-seen dictionary encoding with a hash table per column: O(N).
//
// All of it is integer / byte work and bit-exact by construction: codes, sizes, ids and copied rows are compared with the
// reference's outputs in tests/test_dataset.py (fixtures generated from the reference's own functions).
#include "fmb_common.cuh"

namespace {

// ---- row gather: dst row i = src row index[i]; one warp per row (F ids + F values are contiguous: coalesced), the
// number of positive labels of the batch (the reference's ratio_list entry, data_preprocess.py:170-171) counted on the way
__global__ void __launch_bounds__(256) take_rows_kernel(const int32_t* __restrict__ ids, const float* __restrict__ xv,
                                                        const float* __restrict__ y, int F, int64_t n_src,
                                                        const int64_t* __restrict__ index, int64_t n,
                                                        int32_t* __restrict__ ids_o, float* __restrict__ xv_o,
                                                        float* __restrict__ y_o, int32_t* __restrict__ pos_count,
                                                        int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    int pos = 0;
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
        const int64_t r = __ldg(index + i);
        if (r < 0 || r >= n_src) {          // the reference's list indexing would raise IndexError
            if (lane == 0) atomicExch(err, 1);
            continue;
        }
        for (int f = lane; f < F; f += 32) {
            ids_o[i * F + f] = __ldg(ids + r * F + f);
            if (xv) xv_o[i * F + f] = __ldg(xv + r * F + f);
        }
        if (lane == 0 && y) {
            const float yy = __ldg(y + r);
            y_o[i] = yy;
            pos += yy == 1.0f;
        }
    }
    if (pos_count && lane == 0 && pos) atomicAdd(pos_count, pos);
}

// ---- per-field local ids -> global row ids with the range check nn.Embedding performs in the reference
// (IndexError: index out of range): local int64 [N,F] -> ids int32 [N,F]
__global__ void __launch_bounds__(256) encode_ids_kernel(const int64_t* __restrict__ local, int64_t n, int F,
                                                         const int32_t* __restrict__ field_off /*[F+1]*/,
                                                         int32_t* __restrict__ ids, int* __restrict__ err) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * F; e += (int64_t)gridDim.x * blockDim.x) {
        const int f = (int)(e % F);
        const int64_t v = __ldg(local + e);
        const int32_t lo = __ldg(field_off + f), hi = __ldg(field_off + f + 1);
        if (v < 0 || v >= (int64_t)(hi - lo)) { atomicExch(err, 1); ids[e] = lo; }
        else ids[e] = lo + (int32_t)v;
    }
}

// ---- first-seen dictionary encoding --------------------------------------------------------------------------------
// X [N,d] fp64 row-major.  Column c's code of row i = the number of DISTINCT values of column c that appeared before the
// first appearance of X[i,c] (Python's `list.index` compares with ==, so -0.0 and +0.0 are one entry: canonicalised).
constexpr unsigned long long DICT_EMPTY = 0xffffffffffffffffull;   // a NaN pattern: NaNs are rejected (error flag)

__device__ __forceinline__ unsigned long long dict_key(double v) {
    if (v == 0.0) v = 0.0;                                         // -0.0 -> +0.0
    return (unsigned long long)__double_as_longlong(v);
}
__device__ __forceinline__ uint32_t dict_hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (uint32_t)k;
}

// pass 1: insert every (row, column) into its column's table; slot's `first` = the smallest row holding the key.
// codes[i,c] temporarily holds the slot.
__global__ void __launch_bounds__(256) dict_insert_kernel(const double* __restrict__ X, int64_t N, int d, uint32_t cap_mask,
                                                          unsigned long long* __restrict__ keys /*[d][cap]*/,
                                                          int32_t* __restrict__ first /*[d][cap]*/,
                                                          int32_t* __restrict__ codes, int* __restrict__ err) {
    const size_t cap = (size_t)cap_mask + 1;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < N * d; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % d);
        const int64_t i = e / d;
        const double v = __ldg(X + e);
        if (v != v) { atomicExch(err, 2); codes[e] = 0; continue; }
        const unsigned long long key = dict_key(v);
        uint32_t h = dict_hash(key) & cap_mask;
        unsigned long long* kc = keys + (size_t)c * cap;
        for (;;) {
            unsigned long long old = kc[h];
            if (old == DICT_EMPTY) old = atomicCAS(kc + h, DICT_EMPTY, key);
            if (old == DICT_EMPTY || old == key) break;
            h = (h + 1) & cap_mask;
        }
        atomicMin(first + (size_t)c * cap + h, (int32_t)i);
        codes[e] = (int32_t)h;
    }
}
// pass 2: flag[c][i] = 1 when row i is the first appearance of its value in column c
__global__ void __launch_bounds__(256) dict_mark_kernel(int64_t N, int d, uint32_t cap_mask, const int32_t* __restrict__ first,
                                                        const int32_t* __restrict__ codes, int32_t* __restrict__ rank /*[d][N]*/) {
    const size_t cap = (size_t)cap_mask + 1;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < N * d; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % d);
        const int64_t i = e / d;
        rank[(size_t)c * N + i] = first[(size_t)c * cap + (uint32_t)codes[e]] == (int32_t)i;
    }
}
// pass 3: exclusive prefix sum of the flags of one column per CTA (in place); sizes[c] = number of distinct values.
// 1 024 threads x 4 consecutive elements per round, warp shuffles + one shared-memory level, running carry.
__global__ void __launch_bounds__(1024) dict_scan_kernel(int64_t N, int32_t* __restrict__ rank, int32_t* __restrict__ sizes) {
    __shared__ int32_t wsum[32];
    __shared__ int32_t carry_s;
    int32_t* r = rank + (size_t)blockIdx.x * N;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < N; base += 4096) {
        const int64_t i0 = base + (int64_t)threadIdx.x * 4;
        int32_t v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) v[t] = i0 + t < N ? r[i0 + t] : 0;
        const int32_t mine = v[0] + v[1] + v[2] + v[3];
        int32_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        if (w == 0) {
            int32_t s = wsum[lane], si = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int32_t u = __shfl_up_sync(0xffffffffu, si, o); if (lane >= o) si += u; }
            wsum[lane] = si - s;                 // exclusive prefix of the warp totals
        }
        __syncthreads();
        const int32_t carry = carry_s;
        int32_t ex = carry + wsum[w] + inc - mine;
#pragma unroll
        for (int t = 0; t < 4; ++t) { if (i0 + t < N) r[i0 + t] = ex; ex += v[t]; }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = ex;   // the last thread's running value = carry + the round's total
        __syncthreads();
    }
    if (threadIdx.x == 0) sizes[blockIdx.x] = carry_s;
}
// pass 4: code = rank of the value's first appearance
__global__ void __launch_bounds__(256) dict_codes_kernel(int64_t N, int d, uint32_t cap_mask, const int32_t* __restrict__ first,
                                                         const int32_t* __restrict__ rank, int32_t* __restrict__ codes) {
    const size_t cap = (size_t)cap_mask + 1;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < N * d; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % d);
        const int32_t fi = first[(size_t)c * cap + (uint32_t)codes[e]];
        codes[e] = (fi >= 0 && (int64_t)fi < N) ? rank[(size_t)c * N + fi] : -1;   // -1: a NaN cell (error flag set)
    }
}

uint32_t dict_cap(int64_t N) { uint32_t c = 64; while ((int64_t)c < 2 * N) c <<= 1; return c; }
unsigned grid_for(int64_t items) { int64_t g = (items + 255) / 256; if (g > 148 * 16) g = 148 * 16; if (g < 1) g = 1; return (unsigned)g; }

}  // namespace

// take rows `index_dev[0..n)` of a device-resident data set: the batch builders of utils/data_preprocess.py:154-264 and
// the balance_* functions (:46-82, :120-151), which append one Python list per sample.  xv_src / y_src may be NULL (all-one
// values / unlabeled).  pos_count_dev (optional, zeroed by the caller): labels equal to 1 among the taken rows (ratio_list).
// err_dev: int, set to 1 when an index is outside [0, n_src) (the reference raises IndexError).
FMB_API int fmb_dataset_take(const int32_t* ids_src, const float* xv_src, const float* y_src, int F, int64_t n_src,
                             const int64_t* index_dev, int64_t n, int32_t* ids_dst, float* xv_dst, float* y_dst,
                             int32_t* pos_count_dev, int* err_dev, cudaStream_t stream) {
    FMB_CHECK_ARG(ids_src && index_dev && ids_dst && err_dev && F > 0 && n >= 0 && n_src >= 0, "fmb_dataset_take: bad arguments");
    FMB_CHECK_ARG((!xv_src || xv_dst) && (!y_src || y_dst), "fmb_dataset_take: missing destination");
    if (n == 0) return FMB_OK;
    int64_t g = (n + 7) / 8;
    if (g > 148 * 8) g = 148 * 8;
    take_rows_kernel<<<(unsigned)g, 256, 0, stream>>>(ids_src, xv_src, y_src, F, n_src, index_dev, n, ids_dst, xv_dst, y_dst,
                                                      pos_count_dev, err_dev);
    FMB_CHECK_LAUNCH("take_rows_kernel");
    return FMB_OK;
}

// per-field local ids (int64, what torch.LongTensor(Xi) holds in deepfm_adam.py:47) -> global row ids; err_dev set to 1 on
// an id outside [0, feature_sizes[f]) (nn.Embedding's IndexError).  field_off_dev: int32 [F+1] exclusive prefix of the sizes.
FMB_API int fmb_dataset_encode_ids(const int64_t* local_dev, int64_t n, int F, const int32_t* field_off_dev, int32_t* ids_dev,
                                   int* err_dev, cudaStream_t stream) {
    FMB_CHECK_ARG(local_dev && field_off_dev && ids_dev && err_dev && F > 0 && n >= 0, "fmb_dataset_encode_ids: bad arguments");
    if (n == 0) return FMB_OK;
    encode_ids_kernel<<<grid_for(n * F), 256, 0, stream>>>(local_dev, n, F, field_off_dev, ids_dev, err_dev);
    FMB_CHECK_LAUNCH("encode_ids_kernel");
    return FMB_OK;
}

FMB_API size_t fmb_dict_encode_workspace_bytes(int64_t N, int d) {
    if (N <= 0 || d <= 0) return 0;
    const size_t cap = dict_cap(N);
    return (size_t)d * cap * (sizeof(unsigned long long) + sizeof(int32_t)) + (size_t)d * (size_t)N * sizeof(int32_t);
}

// read_svm_file's vocabulary build (utils/data_preprocess.py:100-108) for every column at once: codes_dev [N,d] int32 =
// index of X[i,c] in the list of column c's distinct values in order of first appearance; sizes_dev [d] = lengths of those
// lists (the reference's feature_sizes).  err_dev: set to 2 when X holds a NaN.
FMB_API int fmb_dict_encode_first_seen(const double* X_dev, int64_t N, int d, int32_t* codes_dev, int32_t* sizes_dev,
                                       int* err_dev, void* ws_dev, size_t ws_bytes, cudaStream_t stream) {
    FMB_CHECK_ARG(X_dev && codes_dev && sizes_dev && err_dev && ws_dev && N > 0 && d > 0 && N < ((int64_t)1 << 30),
                  "fmb_dict_encode_first_seen: bad arguments");
    if (ws_bytes < fmb_dict_encode_workspace_bytes(N, d)) { fmb_set_error("fmb_dict_encode_first_seen: workspace too small"); return FMB_ERR_WS; }
    const uint32_t cap = dict_cap(N);
    unsigned long long* keys = (unsigned long long*)ws_dev;
    int32_t* first = (int32_t*)(keys + (size_t)d * cap);
    int32_t* rank = first + (size_t)d * cap;
    cudaError_t e = cudaMemsetAsync(keys, 0xff, (size_t)d * cap * sizeof(unsigned long long), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(first, 0x7f, (size_t)d * cap * sizeof(int32_t), stream);
    if (e != cudaSuccess) { fmb_set_error("fmb_dict_encode_first_seen: %s", cudaGetErrorString(e)); return FMB_ERR_CUDA; }
    const unsigned g = grid_for(N * d);
    dict_insert_kernel<<<g, 256, 0, stream>>>(X_dev, N, d, cap - 1, keys, first, codes_dev, err_dev);
    FMB_CHECK_LAUNCH("dict_insert_kernel");
    dict_mark_kernel<<<g, 256, 0, stream>>>(N, d, cap - 1, first, codes_dev, rank);
    FMB_CHECK_LAUNCH("dict_mark_kernel");
    dict_scan_kernel<<<d, 1024, 0, stream>>>(N, rank, sizes_dev);
    FMB_CHECK_LAUNCH("dict_scan_kernel");
    dict_codes_kernel<<<g, 256, 0, stream>>>(N, d, cap - 1, first, rank, codes_dev);
    FMB_CHECK_LAUNCH("dict_codes_kernel");
    return FMB_OK;
}
