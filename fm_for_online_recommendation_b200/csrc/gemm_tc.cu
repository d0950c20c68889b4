// gemm_tc.cu -- tcgen05 (5th-generation tensor core) GEMM for the DeepFM / NFM tower where it is a real
// dense contraction (BASELINE configs[3]: B = 8192, H = 400: 15.9 GFLOP per fit step).
//
// C[M,N] = epilogue( A(M x K) * B(K x N) ), operands addressed through (row, col) strides like the SIMT
// kernel of mlp.cu, so the forward (NT), dX (NN) and dW (TN) products all map onto it.  fp32 in, fp32 out:
// every operand is split x = hi + lo with hi = x truncated to TF32, and three tcgen05.mma.kind::tf32
// products (lo*hi, hi*lo, hi*hi) accumulate in fp32 in tensor memory ("3xTF32": ~1e-6 relative, inside the
// 1e-5 tolerance of BASELINE.json; plain TF32 would be ~1e-3).
//
// One CTA computes a 128 x TBN tile (TBN = 128 or 208: 400 = 2 x 208 - 16 wastes 4 % instead of 28 %).
// K is walked in slices of 32 through a two-stage shared-memory ring:
//   * all 256 threads prefetch the NEXT slice from global memory into registers, then split and store the
//     current one into the canonical K-major no-swizzle core-matrix layout.  The layout is chosen for the
//     staging, not the other way round: 8-row groups are contiguous (SBO = 128 B) and the K-direction stride
//     is rows*16 + 16 B (LBO), so a 16-byte chunk (row r, k-chunk c) sits at c*LBO + r*16 and both a K-contiguous
//     source (lanes = 8 k-chunks x 4 rows) and a row-contiguous source (lanes = 32 rows) store conflict-free;
//   * one elected thread issues the 12 MMAs of the slice and a tcgen05.commit on the stage's mbarrier; the
//     stage is refilled only after that commit, so the MMAs of slice t run under the staging of slice t+1.
// K is also split across CTAs (blockIdx.z) when the tile grid alone cannot fill 148 SMs (the dW product:
// M = N = 400, K = batch) and whenever one pass would accumulate more than 1024 terms in TMEM (the tensor
// core's fp32 accumulator truncates; error grows with the K of one pass).  Split partials go to a workspace
// and a second kernel adds them in split order and applies the epilogue, so results are run-to-run identical.
//
// This path is NOT bit-comparable with the oracle (the tensor core's summation order is not the oracle's
// left-to-right FMA chain); mlp.cu picks it only for large shapes and tests compare it within tolerance.
#include "fmb_common.cuh"

namespace {

constexpr int TBM = 128, TBK = 32;
constexpr int TC_THREADS = 256;
constexpr int KCH = TBK / 4;                        // 16-byte k-chunks per slice
constexpr int MAX_K_PER_PASS = 1024;

struct TcParams {
    const float* A; int64_t sam, sak;
    const float* B; int64_t sbk, sbn;
    float* C; int64_t scm;
    int M, N, K;
    int kchunk;              // K range of one split (multiple of TBK); gridDim.z splits
    int epi;                 // 0 none, 1 bias+relu, 2 mask
    const float* bias;
    const float* mask; int64_t smm;
    float* colsum;           // colsum[m] = sum_k A(m,k)
    float* ws;               // split partials [splits][M][N] (gridDim.z > 1)
    float* wsc;              // split partial column sums [splits][M]
    int* error;              // set to 1 if an mbarrier wait times out (never expected)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // K-direction stride between core matrices
    d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;        // SBO: 8-row groups are contiguous
    d |= (uint64_t)1 << 46;                             // descriptor version of sm_100
    return d;                                           // base offset 0, swizzle mode 0 (none)
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__device__ __forceinline__ bool mbar_wait(uint32_t addr, uint32_t parity) {
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t r;
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                     : "=r"(r) : "r"(addr), "r"(parity) : "memory");
        if (r) return true;
    }
    return false;
}

// One operand's slice held in registers: NJ 16-byte chunks per thread.
//   K-contiguous source   : chunk j = (row (lane>>3) + 4*warp + 32*j, k-chunk lane&7)
//   row-contiguous source : chunk j = (row lane + 32*j,               k-chunk warp)
template <int ROWS>
struct Slice {
    static constexpr int NJ = (ROWS + 31) / 32;
    float4 v[NJ];

    __device__ __forceinline__ void load(const float* __restrict__ src, int64_t srow, int64_t sk, bool kcontig, bool vec,
                                         int row0, int rows_valid, int k0, int kend, int warp, int lane) {
        if (kcontig) {
            const int kc = lane & 7, gk = k0 + 4 * kc;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int row = (lane >> 3) + 4 * warp + 32 * j, g = row0 + row;
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < ROWS && g < rows_valid && gk < kend) {
                    const float* q = src + g * srow + gk;
                    if (vec && gk + 3 < kend) x = *reinterpret_cast<const float4*>(q);
                    else {
                        x.x = q[0];
                        if (gk + 1 < kend) x.y = q[1];
                        if (gk + 2 < kend) x.z = q[2];
                        if (gk + 3 < kend) x.w = q[3];
                    }
                }
                v[j] = x;
            }
        } else {
            const int gk = k0 + 4 * warp;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int row = lane + 32 * j, g = row0 + row;
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < ROWS && g < rows_valid) {
                    const float* q = src + g * srow + gk * sk;
                    if (gk < kend) x.x = q[0];
                    if (gk + 1 < kend) x.y = q[sk];
                    if (gk + 2 < kend) x.z = q[2 * sk];
                    if (gk + 3 < kend) x.w = q[3 * sk];
                }
                v[j] = x;
            }
        }
    }

    // split into TF32 hi + remainder lo and store into the stage (byte offsets; lbo = ROWS*16 + 16)
    __device__ __forceinline__ void store(unsigned char* hi, unsigned char* lo, bool kcontig, int warp, int lane) const {
        constexpr int LBO = ROWS * 16 + 16;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int row = kcontig ? (lane >> 3) + 4 * warp + 32 * j : lane + 32 * j;
            const int kc = kcontig ? (lane & 7) : warp;
            if (row >= ROWS) continue;
            const float4 x = v[j];
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
            h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
            h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
            h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
            const int off = kc * LBO + row * 16;
            *reinterpret_cast<float4*>(hi + off) = h;
            *reinterpret_cast<float4*>(lo + off) = l;
        }
    }
};

template <int TBN>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(TcParams p) {
    constexpr int LBO_A = TBM * 16 + 16, LBO_B = TBN * 16 + 16;
    constexpr int A_BYTES = KCH * LBO_A, B_BYTES = KCH * LBO_B;
    constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    constexpr int TMEM_COLS = TBN <= 128 ? 128 : 256;
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * TBN;
    const int kbeg = blockIdx.z * p.kchunk, kend = min(p.K, kbeg + p.kchunk);
    const int T = (kend - kbeg + TBK - 1) / TBK;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = TBN, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);

    const bool a_kc = (p.sak == 1), b_kc = (p.sbk == 1);
    const int64_t a_srow = p.sam, b_srow = p.sbn;
    const bool a_vec = a_kc && (p.sam % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
    const bool b_vec = b_kc && (p.sbn % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0);
    const bool do_colsum = p.colsum && blockIdx.x == 0;
    Slice<TBM> sa;
    Slice<TBN> sb;
    float ps[Slice<TBM>::NJ];
#pragma unroll
    for (int j = 0; j < Slice<TBM>::NJ; ++j) ps[j] = 0.f;
    bool ok = true;

    if (T > 0) {
        sa.load(p.A, a_srow, p.sak, a_kc, a_vec, m0, p.M, kbeg, kend, warp, lane);
        sb.load(p.B, b_srow, p.sbk, b_kc, b_vec, n0, p.N, kbeg, kend, warp, lane);
    }
    for (int t = 0; t < T; ++t) {
        const int s = t & 1;
        unsigned char* st = tc_smem + s * STAGE_BYTES;
        if (t >= 2) ok &= mbar_wait(smem_u32(&mbar[s]), ((t - 2) >> 1) & 1);   // MMAs of slice t-2 have read stage s
        sa.store(st, st + A_BYTES, a_kc, warp, lane);
        sb.store(st + 2 * A_BYTES, st + 2 * A_BYTES + B_BYTES, b_kc, warp, lane);
        if (do_colsum) {
#pragma unroll
            for (int j = 0; j < Slice<TBM>::NJ; ++j) ps[j] += (sa.v[j].x + sa.v[j].y) + (sa.v[j].z + sa.v[j].w);
        }
        if (t + 1 < T) {
            const int k0 = kbeg + (t + 1) * TBK;
            sa.load(p.A, a_srow, p.sak, a_kc, a_vec, m0, p.M, k0, kend, warp, lane);
            sb.load(p.B, b_srow, p.sbk, b_kc, b_vec, n0, p.N, k0, kend, warp, lane);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic smem writes -> tensor-core proxy
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t ah = smem_u32(st), al = ah + A_BYTES, bh = ah + 2 * A_BYTES, bl = bh + B_BYTES;
#pragma unroll
            for (int q = 0; q < TBK / 8; ++q) {   // one MMA consumes K = 8 (two core matrices along K)
                const uint32_t ao = q * 2 * LBO_A, bo = q * 2 * LBO_B;
                mma_tf32(tmem, make_desc(al + ao, LBO_A), make_desc(bh + bo, LBO_B), idesc, (t | q) ? 1u : 0u);
                mma_tf32(tmem, make_desc(ah + ao, LBO_A), make_desc(bl + bo, LBO_B), idesc, 1u);
                mma_tf32(tmem, make_desc(ah + ao, LBO_A), make_desc(bh + bo, LBO_B), idesc, 1u);
            }
            // arrive on the stage's mbarrier when every MMA issued so far has finished reading shared memory
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar[s])) : "memory");
        }
    }
    if (T > 0) ok &= mbar_wait(smem_u32(&mbar[(T - 1) & 1]), ((T - 1) >> 1) & 1);   // commits complete in order
    if (!ok && p.error) *p.error = 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

    // ---- column sums of A: fold the per-thread partials through shared memory (the ring is idle now)
    if (do_colsum) {
        __syncthreads();
        float* red = reinterpret_cast<float*>(tc_smem);   // [8][128]
#pragma unroll
        for (int j = 0; j < Slice<TBM>::NJ; ++j) {
            const int row = a_kc ? (lane >> 3) + 4 * warp + 32 * j : lane + 32 * j;
            const int part = a_kc ? (lane & 7) : warp;
            red[part * TBM + row] = ps[j];
        }
        __syncthreads();
        if (tid < TBM && m0 + tid < p.M) {
            float c = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) c += red[q * TBM + tid];
            if (gridDim.z > 1) p.wsc[(size_t)blockIdx.z * p.M + m0 + tid] = c;
            else p.colsum[m0 + tid] = c;
        }
    }

    // ---- epilogue: warp w reads TMEM lanes [32(w&3), +32) = rows m0 + 32(w&3) + lane, column half w>>2
    {
        const int gm = m0 + (warp & 3) * 32 + lane;
        const bool split = gridDim.z > 1;
        float* crow = split ? p.ws + ((size_t)blockIdx.z * p.M + gm) * p.N : p.C + gm * p.scm;
        const bool vec = split ? (p.N % 4 == 0) : (p.scm % 4 == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
        const int cbeg = (warp >> 2) * (TBN / 2);
        for (int c0 = cbeg; c0 < cbeg + TBN / 2; c0 += 8) {
            if (T == 0 || n0 + c0 >= p.N) break;   // warp-uniform
            uint32_t v[8];
            const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            if (gm >= p.M) continue;
            float x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int gn = n0 + c0 + j;
                x[j] = __uint_as_float(v[j]);
                if (!split && gn < p.N) {
                    if (p.epi == 1) { x[j] = __fadd_rn(x[j], p.bias[gn]); x[j] = x[j] > 0.f ? x[j] : 0.f; }
                    else if (p.epi == 2) { x[j] = p.mask[gm * p.smm + gn] > 0.f ? x[j] : 0.f; }
                }
            }
            const int gn0 = n0 + c0;
            if (vec && gn0 + 7 < p.N) {
                *reinterpret_cast<float4*>(crow + gn0) = make_float4(x[0], x[1], x[2], x[3]);
                *reinterpret_cast<float4*>(crow + gn0 + 4) = make_float4(x[4], x[5], x[6], x[7]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (gn0 + j < p.N) crow[gn0 + j] = x[j];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(TMEM_COLS));
}

// C = epilogue( sum over splits, in split order ); colsum likewise
__global__ void splitk_reduce_kernel(TcParams p, int splits) {
    const int64_t total = (int64_t)p.M * p.N;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) {
        float x = 0.f;
        for (int s = 0; s < splits; ++s) x = __fadd_rn(x, p.ws[(size_t)s * total + i]);
        const int gm = (int)(i / p.N), gn = (int)(i % p.N);
        if (p.epi == 1) { x = __fadd_rn(x, p.bias[gn]); x = x > 0.f ? x : 0.f; }
        else if (p.epi == 2) { x = p.mask[gm * p.smm + gn] > 0.f ? x : 0.f; }
        p.C[gm * p.scm + gn] = x;
    }
    if (p.colsum && i < p.M) {
        float c = 0.f;
        for (int s = 0; s < splits; ++s) c = __fadd_rn(c, p.wsc[(size_t)s * p.M + i]);
        p.colsum[i] = c;
    }
}

template <int TBN>
constexpr size_t tc_smem_bytes() { return (size_t)2 * (2 * KCH * (TBM * 16 + 16) + 2 * KCH * (TBN * 16 + 16)) + 1024; }

}  // namespace

static int* g_tc_error = nullptr;
static float* g_tc_ws = nullptr;
static size_t g_tc_ws_bytes = 0;
static int g_tc_sms = 0;

// C = epilogue(A * B) on the tensor cores (see file header). Same operand convention as mlp.cu's gemm.
int fmb_gemm_tc_launch(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C,
                       int64_t scm, int M, int N, int K, int epi, const float* bias, const float* mask, int64_t smm,
                       float* colsum, cudaStream_t stream) {
    if (!g_tc_error) {
        if (cudaMalloc(&g_tc_error, sizeof(int)) != cudaSuccess) return FMB_ERR_CUDA;
        cudaMemset(g_tc_error, 0, sizeof(int));
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_tc_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_tc_sms <= 0) g_tc_sms = 148;
        cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes<128>());
        cudaFuncSetAttribute(gemm_tc_kernel<208>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes<208>());
    }
    // tile width: the one that wastes less of N (ties -> 128)
    const int n128 = (N + 127) / 128, n208 = (N + 207) / 208;
    const int tbn = (n208 * 208 < n128 * 128) ? 208 : 128;
    const int tiles = ((M + TBM - 1) / TBM) * (tbn == 208 ? n208 : n128);
    // K splits: enough CTAs for one wave, and never more than MAX_K_PER_PASS terms in one TMEM pass
    int splits = 1;
    if (tiles * 2 <= g_tc_sms) splits = g_tc_sms / tiles;
    splits = max(splits, (K + MAX_K_PER_PASS - 1) / MAX_K_PER_PASS);
    splits = min(splits, (K + 4 * TBK - 1) / (4 * TBK));   // at least 4 slices per split
    splits = max(splits, 1);
    int kchunk = ((K + splits - 1) / splits + TBK - 1) / TBK * TBK;
    splits = (K + kchunk - 1) / kchunk;
    TcParams p;
    p.A = A; p.sam = sam; p.sak = sak; p.B = B; p.sbk = sbk; p.sbn = sbn; p.C = C; p.scm = scm;
    p.M = M; p.N = N; p.K = K; p.kchunk = kchunk; p.epi = epi; p.bias = bias; p.mask = mask; p.smm = smm; p.colsum = colsum;
    p.error = g_tc_error; p.ws = nullptr; p.wsc = nullptr;
    if (splits > 1) {
        const size_t need = ((size_t)splits * M * N + (size_t)splits * M) * sizeof(float);
        if (need > g_tc_ws_bytes) {
            cudaStreamSynchronize(stream);   // earlier launches may still read the old buffer
            if (g_tc_ws) cudaFree(g_tc_ws);
            g_tc_ws = nullptr; g_tc_ws_bytes = 0;
            if (cudaMalloc(&g_tc_ws, need) != cudaSuccess) { fmb_set_error("gemm_tc: cannot allocate %zu workspace bytes", need); return FMB_ERR_CUDA; }
            g_tc_ws_bytes = need;
        }
        p.ws = g_tc_ws;
        p.wsc = g_tc_ws + (size_t)splits * M * N;
    }
    dim3 grid(tbn == 208 ? n208 : n128, (M + TBM - 1) / TBM, splits);
    if (tbn == 208) gemm_tc_kernel<208><<<grid, TC_THREADS, tc_smem_bytes<208>(), stream>>>(p);
    else gemm_tc_kernel<128><<<grid, TC_THREADS, tc_smem_bytes<128>(), stream>>>(p);
    if (splits > 1) {
        const int64_t total = (int64_t)M * N;
        splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p, splits);
    }
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

// Same through arbitrary strides and the fused epilogues (tests of the TN / NN forms):
//   C[m*scm + n] = epi( sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] ), colsum[m] = sum_k A(m,k) (nullable)
FMB_API int fmb_gemm_tc_strided(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn,
                                float* C, int64_t scm, int M, int N, int K, int epi, const float* bias,
                                const float* mask, int64_t smm, float* colsum, cudaStream_t stream) {
    FMB_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0, "fmb_gemm_tc_strided: bad arguments");
    FMB_CHECK_ARG(epi >= 0 && epi <= 2 && (epi != 1 || bias) && (epi != 2 || mask), "fmb_gemm_tc_strided: bad epilogue");
    return fmb_gemm_tc_launch(A, sam, sak, B, sbk, sbn, C, scm, M, N, K, epi, bias, mask, smm, colsum, stream);
}
