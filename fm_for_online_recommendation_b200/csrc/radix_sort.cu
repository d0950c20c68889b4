// radix_sort.cu -- deterministic stable sort of (row id, entry index) pairs + segment detection.
//
// The reference sums the gradients of duplicate rows in sample order (torch's CPU
// embedding_dense_backward).  The B200 path gets the same order without float
// atomics: a STABLE least-significant-digit radix sort of the B*F global row ids carries the entry
// index (b*F+f) as payload, so inside every run of equal row ids the entries stay in sample order.
// Integer work only; every pass is three small kernels (tile histogram -> scan -> stable scatter).
// Ranking inside a tile uses __match_any_sync on a warp-striped key arrangement (no atomics on the
// ordering path, so the permutation is bit-reproducible).
#include "fmb_common.cuh"
#include "smem_sort.cuh"

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int KPT = 8;                       // keys per thread
constexpr int TILE = SORT_THREADS * KPT;     // 2048 keys per CTA

__device__ __forceinline__ int digit_of(int32_t key, int shift) { return (key >> shift) & (RADIX - 1); }

using fmb::digit_peers;

// hist[d * ntiles + tile] = number of keys of `tile` whose digit is d
__global__ void __launch_bounds__(SORT_THREADS) radix_hist_kernel(const int32_t* __restrict__ keys, int64_t N,
                                                                  int shift, int ntiles,
                                                                  uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[RADIX];
    for (int i = threadIdx.x; i < RADIX; i += SORT_THREADS) h[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * TILE;
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
        const int64_t idx = base + (int64_t)i * SORT_THREADS + threadIdx.x;
        if (idx < N) atomicAdd(&h[digit_of(keys[idx], shift)], 1u);  // integer: order-independent
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RADIX; i += SORT_THREADS) hist[(size_t)i * ntiles + blockIdx.x] = h[i];
}

// in-place exclusive scan of `n` counters by one CTA of 1024 threads (n = 256 * ntiles)
__global__ void __launch_bounds__(1024) radix_scan_kernel(uint32_t* __restrict__ hist, int64_t n) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    constexpr int IPT = 4;
    for (int64_t base = 0; base < n; base += 1024 * IPT) {
        const int64_t i0 = base + (int64_t)threadIdx.x * IPT;
        uint32_t v[IPT], sum = 0;
#pragma unroll
        for (int j = 0; j < IPT; ++j) { v[j] = (i0 + j < n) ? hist[i0 + j] : 0u; sum += v[j]; }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_tot[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
            warp_tot[lane] = winc - w;  // exclusive over warps
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        uint32_t excl = carry + warp_tot[warp] + (inc - sum);
#pragma unroll
        for (int j = 0; j < IPT; ++j) { if (i0 + j < n) hist[i0 + j] = excl; excl += v[j]; }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl;  // total so far (last thread's running end)
        __syncthreads();
    }
}

// stable scatter of one tile. vals_in == nullptr means "identity payload" (first pass).
__global__ void __launch_bounds__(SORT_THREADS) radix_scatter_kernel(const int32_t* __restrict__ keys_in,
                                                                     const int32_t* __restrict__ vals_in, int64_t N,
                                                                     int shift, int ntiles,
                                                                     const uint32_t* __restrict__ scan,
                                                                     int32_t* __restrict__ keys_out,
                                                                     int32_t* __restrict__ vals_out) {
    __shared__ uint32_t cnt[SORT_WARPS][RADIX];
    __shared__ uint32_t gbase[RADIX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&cnt[0][0])[i] = 0;
    for (int i = threadIdx.x; i < RADIX; i += SORT_THREADS) gbase[i] = scan[(size_t)i * ntiles + blockIdx.x];
    __syncthreads();
    // warp-striped arrangement: warp w owns keys [w*32*KPT, (w+1)*32*KPT) of the tile, slot i is
    // 32 consecutive keys -> original order == (warp, slot, lane) order.
    const int64_t wbase = (int64_t)blockIdx.x * TILE + (int64_t)warp * 32 * KPT;
    int32_t key[KPT];
    uint32_t rank[KPT];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
        const int64_t idx = wbase + i * 32 + lane;
        const bool valid = idx < N;
        key[i] = valid ? keys_in[idx] : 0;
        const unsigned d = valid ? (unsigned)digit_of(key[i], shift) : 0u;
        const unsigned m = digit_peers(d, valid);
        const int leader = __ffs(m) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) { old = cnt[warp][d]; cnt[warp][d] = old + __popc(m); }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[i] = old + __popc(m & lt);
        __syncwarp();
    }
    __syncthreads();
    // exclusive prefix over warps, per digit
    for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) { uint32_t t = cnt[w][d]; cnt[w][d] = run; run += t; }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
        const int64_t idx = wbase + i * 32 + lane;
        if (idx < N) {
            const int d = digit_of(key[i], shift);
            const uint32_t pos = gbase[d] + cnt[warp][d] + rank[i];
            keys_out[pos] = key[i];
            vals_out[pos] = vals_in ? vals_in[idx] : (int32_t)idx;
        }
    }
}

// seg_flag[i] = 1 where a new run of equal keys starts; compacted to seg_start by one CTA (test/API
// path only -- the hot path finds runs inside fm_backward's tiles and never needs this list).
__global__ void __launch_bounds__(1024) segment_starts_kernel(const int32_t* __restrict__ keys, int64_t N,
                                                              int32_t* __restrict__ seg_start,
                                                              int32_t* __restrict__ nseg) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < N; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const uint32_t f = (i < N) && (i == 0 || keys[i] != keys[i - 1]);
        uint32_t inc = f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_tot[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
            warp_tot[lane] = winc - w;
        }
        __syncthreads();
        const uint32_t pos = carry_s + warp_tot[warp] + inc - f;
        if (f) seg_start[pos] = (int32_t)i;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = pos + f;
        __syncthreads();
    }
    if (threadIdx.x == 0) { *nseg = (int32_t)carry_s; seg_start[carry_s] = (int32_t)N; }
}


// ---------------------------------------------------------------------------------------------
// Fast path: ids come as a [B,F] matrix whose column f holds ids of field f only, so the global
// stable sort factors into F independent stable sorts of B keys each.  One CTA per field sorts its
// column entirely in shared memory (keys 32-bit, payload = sample index 16-bit, ping-pong), with
// as many 8-bit passes as the field's cardinality needs (0 for a single-row field), and writes
// sorted_keys/perm at [f*B, (f+1)*B).  The result is bit-identical to the generic global sort.
// ---------------------------------------------------------------------------------------------
constexpr int FS_THREADS = 1024;
constexpr int FS_WARPS = FS_THREADS / 32;
constexpr int FS_MAX_SLOTS = 16;  // B <= 32 warps * 16 slots * 32 lanes = 16384

__global__ void __launch_bounds__(FS_THREADS) sort_fields_kernel(const int32_t* __restrict__ ids, int B, int F,
                                                                 const int32_t* __restrict__ field_off,
                                                                 int32_t* __restrict__ skeys,
                                                                 int32_t* __restrict__ perm) {
    extern __shared__ __align__(16) unsigned char fs_smem[];
    uint32_t* kbuf0 = reinterpret_cast<uint32_t*>(fs_smem);
    uint32_t* kbuf1 = kbuf0 + B;
    uint16_t* pbuf0 = reinterpret_cast<uint16_t*>(kbuf1 + B);
    uint16_t* pbuf1 = pbuf0 + B + (B & 1);
    uint16_t* cnt = pbuf1 + B + (B & 1);                 // [FS_WARPS][RADIX]
    __shared__ uint32_t tot[RADIX];
    const int f = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t off = field_off[f];
    const uint32_t nrows = (uint32_t)(field_off[f + 1] - off);
    const int bits = nrows <= 1 ? 0 : 32 - __clz(nrows - 1);
    const int passes = (bits + RADIX_BITS - 1) / RADIX_BITS;
    for (int i = threadIdx.x; i < B; i += FS_THREADS) {
        kbuf0[i] = (uint32_t)(ids[(size_t)i * F + f] - off);
        pbuf0[i] = (uint16_t)i;
    }
    __syncthreads();
    uint32_t* kc; uint16_t* pc;
    fmb::smem_sort_passes(kbuf0, kbuf1, pbuf0, pbuf1, cnt, tot, B, passes, &kc, &pc);
    for (int i = threadIdx.x; i < B; i += FS_THREADS) {
        skeys[(size_t)f * B + i] = (int32_t)kc[i] + off;
        perm[(size_t)f * B + i] = (int32_t)pc[i] * F + f;
    }
}

}  // namespace

static int sort_ntiles(int64_t N) { return (int)((N + TILE - 1) / TILE); }

// workspace: ping-pong key/val buffers + histogram
FMB_API size_t fmb_sort_workspace_bytes(int64_t N) {
    const size_t n4 = ((size_t)N * 4 + 255) / 256 * 256;
    const size_t h = ((size_t)RADIX * sort_ntiles(N) * 4 + 255) / 256 * 256;
    return 4 * n4 + h;
}

// Stable sort of keys[N] (non-negative int32, < 2^key_bits) -> sorted_keys[N], perm[N] (perm[i] =
// original position of the i-th smallest key; ties keep ascending original position).
// Optional: seg_start[nseg+1] run starts (+ terminating N) and *nseg, for parity checks.
FMB_API int fmb_sort_segment(const int32_t* keys, int64_t N, int key_bits, void* ws, size_t ws_bytes,
                             int32_t* sorted_keys, int32_t* perm, int32_t* seg_start, int32_t* nseg,
                             cudaStream_t stream) {
    FMB_CHECK_ARG(keys && sorted_keys && perm && ws, "fmb_sort_segment: null pointer");
    FMB_CHECK_ARG(N > 0 && N < ((int64_t)1 << 31), "fmb_sort_segment: N out of range");
    FMB_CHECK_ARG(key_bits >= 1 && key_bits <= 31, "fmb_sort_segment: key_bits out of range");
    if (ws_bytes < fmb_sort_workspace_bytes(N)) { fmb_set_error("fmb_sort_segment: workspace too small"); return FMB_ERR_WS; }
    const size_t n4 = ((size_t)N * 4 + 255) / 256 * 256;
    char* w = (char*)ws;
    int32_t* kbuf[2] = {(int32_t*)w, (int32_t*)(w + n4)};
    int32_t* vbuf[2] = {(int32_t*)(w + 2 * n4), (int32_t*)(w + 3 * n4)};
    uint32_t* hist = (uint32_t*)(w + 4 * n4);
    const int ntiles = sort_ntiles(N);
    const int passes = (key_bits + RADIX_BITS - 1) / RADIX_BITS;
    const int32_t* kin = keys;
    const int32_t* vin = nullptr;
    for (int p = 0; p < passes; ++p) {
        const bool last = (p == passes - 1);
        int32_t* kout = last ? sorted_keys : kbuf[p & 1];
        int32_t* vout = last ? perm : vbuf[p & 1];
        const int shift = p * RADIX_BITS;
        radix_hist_kernel<<<ntiles, SORT_THREADS, 0, stream>>>(kin, N, shift, ntiles, hist);
        radix_scan_kernel<<<1, 1024, 0, stream>>>(hist, (int64_t)RADIX * ntiles);
        radix_scatter_kernel<<<ntiles, SORT_THREADS, 0, stream>>>(kin, vin, N, shift, ntiles, hist, kout, vout);
        kin = kout;
        vin = vout;
    }
    FMB_CHECK_LAUNCH("radix sort");
    if (seg_start) {
        FMB_CHECK_ARG(nseg, "fmb_sort_segment: seg_start given without nseg");
        segment_starts_kernel<<<1, 1024, 0, stream>>>(sorted_keys, N, seg_start, nseg);
        FMB_CHECK_LAUNCH("segment_starts_kernel");
    }
    return FMB_OK;
}

// Largest batch the per-field shared-memory sort accepts.
FMB_API int fmb_sort_fields_max_batch(void) { return FS_WARPS * FS_MAX_SLOTS * 32; }

// Stable sort of ids[B,F] (column f holds global row ids of field f, field_off[F+1] device array of
// field offsets) -> sorted_keys[B*F], perm[B*F] (original entry index b*F+f), identical to
// fmb_sort_segment over the flattened matrix.  One CTA per field, everything in shared memory.
FMB_API int fmb_sort_fields(const int32_t* ids, int B, int F, const int32_t* field_off, int32_t* sorted_keys,
                            int32_t* perm, cudaStream_t stream) {
    FMB_CHECK_ARG(ids && field_off && sorted_keys && perm, "fmb_sort_fields: null pointer");
    FMB_CHECK_ARG(B > 0 && B <= fmb_sort_fields_max_batch() && F > 0, "fmb_sort_fields: B=%d out of range", B);
    const size_t sm = (size_t)2 * B * 4 + (size_t)2 * (B + (B & 1)) * 2 + (size_t)FS_WARPS * RADIX * 2;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(sort_fields_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); attr = true; }
    sort_fields_kernel<<<F, FS_THREADS, sm, stream>>>(ids, B, F, field_off, sorted_keys, perm);
    FMB_CHECK_LAUNCH("sort_fields_kernel");
    return FMB_OK;
}
