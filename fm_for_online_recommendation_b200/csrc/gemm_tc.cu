// gemm_tc.cu -- tcgen05 (5th-generation tensor core) GEMM for the DeepFM / NFM tower where it is a real
// dense contraction (BASELINE configs[3]: B = 8192, H = 400: 15.9 GFLOP per fit step).
//
// C[M,N] = epilogue( A(M x K) * B(K x N) ), operands addressed through (row, col) strides like the SIMT
// kernel of mlp.cu, so the forward (NT), dX (NN) and dW (TN) products all map onto it.  fp32 in, fp32 out:
// every operand is split x = hi + lo with hi = x truncated to TF32, and three tcgen05.mma.kind::tf32
// products (lo*hi, hi*lo, hi*hi) accumulate in fp32 in tensor memory ("3xTF32": ~1e-6 relative, inside the
// 1e-5 tolerance of BASELINE.json; plain TF32 would be ~1e-3).  One CTA computes a 128 x 128 tile:
// all threads stage K-slices of 32 into shared memory in the canonical K-major core-matrix layout
// (no swizzle), one elected thread issues the MMAs, completion comes back through tcgen05.commit on an
// mbarrier, and the four warps read the accumulator from TMEM with tcgen05.ld for the epilogue.
// This path is NOT bit-comparable with the oracle (the tensor core's summation order is not the oracle's
// left-to-right FMA chain); mlp.cu picks it only for large shapes and tests compare it within tolerance.
#include "fmb_common.cuh"

namespace {

constexpr int TBM = 128, TBN = 128, TBK = 32;
constexpr int TC_THREADS = 256;
constexpr uint32_t LBO_BYTES = 128;                 // K-adjacent core matrices are contiguous
constexpr uint32_t SBO_BYTES = (TBK / 4) * 128;     // next 8-row group: 8 core matrices further
constexpr int OP_BYTES = TBM * TBK * 4;             // one operand tile (128 rows x 32 k) = 16 KB

struct TcParams {
    const float* A; int64_t sam, sak;
    const float* B; int64_t sbk, sbn;
    float* C; int64_t scm;
    int M, N, K;
    int epi;                 // 0 none, 1 bias+relu, 2 mask
    const float* bias;
    const float* mask; int64_t smm;
    float* colsum;           // colsum[m] = sum_k A(m,k) (fp32 adds, k ascending)
    int* error;              // set to 1 if an mbarrier wait times out (never expected)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((LBO_BYTES >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((SBO_BYTES >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version of sm_100
    return d;                 // base offset 0, swizzle mode 0 (none)
}

// element (row r, k) of a [128 x 32] operand tile in the canonical layout: core matrix = 8 rows x 16 bytes
__device__ __forceinline__ int tile_off(int r, int k) {
    return (r >> 3) * (SBO_BYTES / 4) + (k >> 2) * (LBO_BYTES / 4) + (r & 7) * 4 + (k & 3);
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__global__ void __launch_bounds__(TC_THREADS) gemm_tc_kernel(TcParams p) {
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    float* a_hi = reinterpret_cast<float*>(tc_smem);
    float* a_lo = a_hi + TBM * TBK;
    float* b_hi = a_lo + TBM * TBK;
    float* b_lo = b_hi + TBN * TBK;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * TBN;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(TBN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
    float csum = 0.f;   // colsum of row (m0 + tid) when tid < 128
    unsigned phase = 0;
    bool first = true;
    for (int k0 = 0; k0 < p.K; k0 += TBK) {
        // ---- stage the K-slice: split every element into TF32 hi + remainder lo
        const bool a_kcontig = (p.sak == 1), b_kcontig = (p.sbk == 1);
        for (int e = tid; e < TBM * TBK; e += TC_THREADS) {
            int r, kk;
            if (a_kcontig) { kk = e & (TBK - 1); r = e >> 5; } else { r = e & (TBM - 1); kk = e >> 7; }
            const int gm = m0 + r, gk = k0 + kk;
            const float x = (gm < p.M && gk < p.K) ? p.A[gm * p.sam + gk * p.sak] : 0.f;
            const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
            const int o = tile_off(r, kk);
            a_hi[o] = hi;
            a_lo[o] = x - hi;
        }
        for (int e = tid; e < TBN * TBK; e += TC_THREADS) {
            int r, kk;
            if (b_kcontig) { kk = e & (TBK - 1); r = e >> 5; } else { r = e & (TBN - 1); kk = e >> 7; }
            const int gn = n0 + r, gk = k0 + kk;
            const float x = (gn < p.N && gk < p.K) ? p.B[gk * p.sbk + gn * p.sbn] : 0.f;
            const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
            const int o = tile_off(r, kk);
            b_hi[o] = hi;
            b_lo[o] = x - hi;
        }
        if (p.colsum && blockIdx.x == 0 && tid < TBM) {   // exact fp32 column sums, k ascending
            const int gm = m0 + tid;
            if (gm < p.M)
                for (int kk = 0; kk < TBK && k0 + kk < p.K; ++kk) csum = __fadd_rn(csum, p.A[gm * p.sam + (k0 + kk) * p.sak]);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic smem writes -> tensor-core proxy
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
            for (int s = 0; s < TBK / 8; ++s) {   // one MMA consumes K = 8 (two core matrices along K)
                const uint32_t ko = s * 2 * LBO_BYTES;
                mma_tf32(tmem, make_desc(al + ko), make_desc(bh + ko), idesc, first ? 0u : 1u);
                mma_tf32(tmem, make_desc(ah + ko), make_desc(bl + ko), idesc, 1u);
                mma_tf32(tmem, make_desc(ah + ko), make_desc(bh + ko), idesc, 1u);
                first = false;
            }
            // arrive on the mbarrier when every MMA issued so far has finished reading shared memory
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar)) : "memory");
        }
        first = false;
        // everyone waits: the staging buffers are reused by the next K-slice
        {
            const uint32_t a = smem_u32(&mbar);
            bool ok = false;
            for (int spin = 0; spin < (1 << 22) && !ok; ++spin) {
                uint32_t r;
                asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                             : "=r"(r) : "r"(a), "r"(phase) : "memory");
                ok = r != 0;
            }
            if (!ok && tid == 0 && p.error) *p.error = 1;
            phase ^= 1u;
        }
        __syncthreads();
    }
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    // ---- epilogue: warp w reads TMEM lanes [32w, 32w+32) = rows m0 + 32w + lane
    if (warp < 4) {
        const int gm = m0 + warp * 32 + lane;
        for (int c0 = 0; c0 < TBN; c0 += 8) {
            uint32_t v[8];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            if (gm < p.M) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int gn = n0 + c0 + j;
                    if (gn >= p.N) continue;
                    float x = __uint_as_float(v[j]);
                    if (p.epi == 1) { x = __fadd_rn(x, p.bias[gn]); x = x > 0.f ? x : 0.f; }
                    else if (p.epi == 2) { x = p.mask[gm * p.smm + gn] > 0.f ? x : 0.f; }
                    p.C[gm * p.scm + gn] = x;
                }
            }
        }
        if (p.colsum && blockIdx.x == 0 && gm < p.M) p.colsum[gm] = csum;   // tid == warp*32+lane < 128 here
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(TBN));
}

}  // namespace

static int* g_tc_error = nullptr;

// C = epilogue(A * B) on the tensor cores (see file header). Same operand convention as mlp.cu's gemm.
int fmb_gemm_tc_launch(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C,
                       int64_t scm, int M, int N, int K, int epi, const float* bias, const float* mask, int64_t smm,
                       float* colsum, cudaStream_t stream) {
    if (!g_tc_error) {
        if (cudaMalloc(&g_tc_error, sizeof(int)) != cudaSuccess) return FMB_ERR_CUDA;
        cudaMemset(g_tc_error, 0, sizeof(int));
    }
    TcParams p;
    p.A = A; p.sam = sam; p.sak = sak; p.B = B; p.sbk = sbk; p.sbn = sbn; p.C = C; p.scm = scm;
    p.M = M; p.N = N; p.K = K; p.epi = epi; p.bias = bias; p.mask = mask; p.smm = smm; p.colsum = colsum;
    p.error = g_tc_error;
    static bool attr = false;
    const size_t smem = 4 * OP_BYTES + 1024;
    if (!attr) { cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
    dim3 grid((N + TBN - 1) / TBN, (M + TBM - 1) / TBM);
    gemm_tc_kernel<<<grid, TC_THREADS, smem, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fmb_set_error("gemm_tc_kernel: %s", cudaGetErrorString(e)); return FMB_ERR_CUDA; }
    return FMB_OK;
}

// 1 if a tensor-core GEMM ever timed out on its mbarrier (diagnostic)
FMB_API int fmb_gemm_tc_error(void) {
    int v = 0;
    if (g_tc_error) cudaMemcpy(&v, g_tc_error, sizeof(int), cudaMemcpyDeviceToHost);
    return v;
}

// Standalone tensor-core GEMM for tests: C[M,N] = A[M,K] * B[N,K]^T (both row-major, K contiguous).
FMB_API int fmb_gemm_tc_nt(const float* A, const float* B, float* C, int M, int N, int K, cudaStream_t stream) {
    FMB_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0, "fmb_gemm_tc_nt: bad arguments");
    return fmb_gemm_tc_launch(A, K, 1, B, 1, K, C, N, M, N, K, 0, nullptr, nullptr, 0, nullptr, stream);
}
