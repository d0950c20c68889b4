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
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

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


// ---------------------------------------------------------------------------------------------
// Cluster version of the per-field sort: a thread-block cluster of CL CTAs (CL SMs) sorts one
// field.  CTA r keeps positions [r*Bq, (r+1)*Bq) of the field's current ordering in its shared
// memory; every pass ranks its keys locally, exchanges the 256 digit totals through distributed
// shared memory, and scatters (key, payload) straight into the destination CTA's shared memory.
// Same result as sort_fields_kernel on 4x the SMs (39 CTAs of 1024 threads left 109 SMs idle and
// were issue-bound: profiles/r1c), and batches up to 65536.
// ---------------------------------------------------------------------------------------------
constexpr int CL = 4;
__device__ long long* g_sort_dbg = nullptr;   // debug: per-phase cycle counts of one CTA
#define SORT_PROBE(slot) do { if (dbg && threadIdx.x == 0) { long long t_ = clock64(); dbg[slot] += t_ - tlast; tlast = t_; } } while (0)

// THREADS = 256 for batches up to 16384 (8 warps x <= 16 slots per CTA: the per-pass fixed work -- counter
// reset, per-digit prefix over warps -- scales with the warp count and dominated the 1024-thread version),
// 1024 above that.
// STAGED: keys are first scattered into a LOCAL staging buffer in digit order, then copied to their destination
// CTAs with lane-contiguous (coalesced) distributed-shared-memory stores; the direct version issued two
// scattered 4-/2-byte remote stores per key, which is what bounded it (profiles/r1g).
template <int THREADS, bool STAGED>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS)
sort_fields_cluster_kernel(const int32_t* __restrict__ ids, int B, int F, const int32_t* __restrict__ field_off,
                           int32_t* __restrict__ skeys, int32_t* __restrict__ perm, int Bq) {
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(16) unsigned char fs_smem[];
    uint32_t* kbuf0 = reinterpret_cast<uint32_t*>(fs_smem);
    uint32_t* kbuf1 = kbuf0 + Bq;
    uint16_t* pbuf0 = reinterpret_cast<uint16_t*>(kbuf1 + Bq);
    uint16_t* pbuf1 = pbuf0 + Bq + (Bq & 1);
    uint16_t* cnt = pbuf1 + Bq + (Bq & 1);               // [WARPS][RADIX] (sized for 32 warps)
    uint32_t* kstage = reinterpret_cast<uint32_t*>(cnt + 32 * RADIX);   // [Bq]  (STAGED only)
    uint16_t* pstage = reinterpret_cast<uint16_t*>(kstage + Bq);        // [Bq]
    __shared__ uint32_t lstart[RADIX];                   // start of each digit inside the local staging order
    __shared__ uint32_t tot[RADIX];                      // my digit totals (read by the other CTAs)
    __shared__ uint32_t before_s[RADIX];                 // same-digit keys held by lower-ranked CTAs
    __shared__ uint32_t base[RADIX];                     // field-wide start of each digit
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int f = blockIdx.x / CL, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t off = field_off[f];
    const uint32_t nrows = (uint32_t)(field_off[f + 1] - off);
    const int bits = nrows <= 1 ? 0 : 32 - __clz(nrows - 1);
    const int passes = (bits + RADIX_BITS - 1) / RADIX_BITS;
    const int g0 = crank * Bq;                            // first field position held by this CTA
    const int n = max(0, min(Bq, B - g0));                // entries held by this CTA
    long long* dbg = (g_sort_dbg && blockIdx.x == (unsigned)(F - 1) * CL) ? g_sort_dbg : nullptr;
    long long tlast = clock64();
    for (int i = threadIdx.x; i < n; i += THREADS) {
        kbuf0[i] = (uint32_t)(ids[(size_t)(g0 + i) * F + f] - off);
        pbuf0[i] = (uint16_t)(g0 + i);
    }
    __syncthreads();
    SORT_PROBE(0);
    uint32_t* kc = kbuf0; uint32_t* kn = kbuf1;
    uint16_t* pc = pbuf0; uint16_t* pn = pbuf1;
    const int slots = (Bq + THREADS - 1) / THREADS;
    const int wbase = warp * slots * 32;
    const uint32_t lt = (1u << lane) - 1u;
    for (int ps = 0; ps < passes; ++ps) {
        const int shift = ps * RADIX_BITS;
        for (int i = threadIdx.x; i < WARPS * RADIX / 2; i += THREADS) reinterpret_cast<uint32_t*>(cnt)[i] = 0;
        __syncthreads();
        uint32_t key[FS_MAX_SLOTS];
        uint16_t rank[FS_MAX_SLOTS];
        unsigned peers[FS_MAX_SLOTS];
        // peer masks of all slots first: the ballots of different slots are independent, so their latencies
        // overlap (issued slot after slot they were 560 cycles per slot: 9 dependent VOTEs each)
#pragma unroll
        for (int s = 0; s < FS_MAX_SLOTS; ++s) {
            if (s < slots) {
                const int idx = wbase + s * 32 + lane;
                const bool valid = idx < n;
                key[s] = valid ? kc[idx] : 0u;
                peers[s] = digit_peers(valid ? ((key[s] >> shift) & (RADIX - 1)) : 0u, valid);
            }
        }
        // then the (short) serial part: per-warp digit counters
#pragma unroll
        for (int s = 0; s < FS_MAX_SLOTS; ++s) {
            if (s < slots) {
                const int idx = wbase + s * 32 + lane;
                const bool valid = idx < n;
                const unsigned d = (key[s] >> shift) & (RADIX - 1);
                const unsigned m = peers[s];
                const int leader = __ffs(m) - 1;
                uint32_t old = 0;
                if (valid && lane == leader) { old = cnt[warp * RADIX + d]; cnt[warp * RADIX + d] = (uint16_t)(old + __popc(m)); }
                old = __shfl_sync(0xffffffffu, old, leader);
                rank[s] = (uint16_t)(old + __popc(m & lt));
                __syncwarp();
            }
        }
        __syncthreads();
        SORT_PROBE(1);
        if (threadIdx.x < RADIX) {
            const int d = threadIdx.x;
            uint32_t run = 0;
            for (int w = 0; w < WARPS; ++w) { const uint32_t t = cnt[w * RADIX + d]; cnt[w * RADIX + d] = (uint16_t)run; run += t; }
            tot[d] = run;
        }
        if (STAGED) {
            __syncthreads();
            SORT_PROBE(2);
            if (warp == 0) {  // exclusive scan of MY digit totals -> local staging order
                uint32_t v[8], sum = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { v[j] = tot[lane * 8 + j]; sum += v[j]; }
                uint32_t inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                uint32_t ex = inc - sum;
#pragma unroll
                for (int j = 0; j < 8; ++j) { lstart[lane * 8 + j] = ex; ex += v[j]; }
            }
            __syncthreads();
#pragma unroll
            for (int s = 0; s < FS_MAX_SLOTS; ++s) {
                if (s < slots) {
                    const int idx = wbase + s * 32 + lane;
                    if (idx < n) {
                        const unsigned d = (key[s] >> shift) & (RADIX - 1);
                        const uint32_t lp = lstart[d] + cnt[warp * RADIX + d] + rank[s];
                        kstage[lp] = key[s];
                        pstage[lp] = pc[idx];
                    }
                }
            }
        }
        SORT_PROBE(3);
        cluster.sync();   // every CTA's totals are published; the previous pass's remote scatter is complete
        SORT_PROBE(4);
        if (threadIdx.x < RADIX) {
            const int d = threadIdx.x;
            uint32_t all = 0, before = 0;
#pragma unroll
            for (int r = 0; r < CL; ++r) {
                const uint32_t t = *cluster.map_shared_rank(&tot[d], r);
                all += t;
                if (r < crank) before += t;
            }
            base[d] = all;
            before_s[d] = before;
        }
        __syncthreads();
        if (warp == 0) {  // exclusive scan of the 256 column totals, 8 per lane
            uint32_t v[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { v[j] = base[lane * 8 + j]; sum += v[j]; }
            uint32_t inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            uint32_t ex = inc - sum;
#pragma unroll
            for (int j = 0; j < 8; ++j) { base[lane * 8 + j] = ex + before_s[lane * 8 + j]; ex += v[j]; }
        }
        __syncthreads();
        SORT_PROBE(5);
        if (STAGED) {
            // staged order -> destination CTAs: consecutive lanes write consecutive remote addresses
            for (int i = threadIdx.x; i < n; i += THREADS) {
                const uint32_t kk = kstage[i];
                const unsigned d = (kk >> shift) & (RADIX - 1);
                const uint32_t pos = base[d] + ((uint32_t)i - lstart[d]);
                const int dest = (pos >= (uint32_t)Bq) + (pos >= 2u * (uint32_t)Bq) + (pos >= 3u * (uint32_t)Bq);
                const uint32_t li = pos - (uint32_t)dest * (uint32_t)Bq;
                cluster.map_shared_rank(kn, dest)[li] = kk;
                cluster.map_shared_rank(pn, dest)[li] = pstage[i];
            }
        } else {
            // scatter (key, payload) straight into the destination CTA's next buffer
#pragma unroll
            for (int s = 0; s < FS_MAX_SLOTS; ++s) {
                if (s < slots) {
                    const int idx = wbase + s * 32 + lane;
                    if (idx < n) {
                        const unsigned d = (key[s] >> shift) & (RADIX - 1);
                        const uint32_t pos = base[d] + cnt[warp * RADIX + d] + rank[s];
                        const int dest = (pos >= (uint32_t)Bq) + (pos >= 2u * (uint32_t)Bq) + (pos >= 3u * (uint32_t)Bq);
                        const uint32_t li = pos - (uint32_t)dest * (uint32_t)Bq;
                        cluster.map_shared_rank(kn, dest)[li] = key[s];
                        cluster.map_shared_rank(pn, dest)[li] = pc[idx];
                    }
                }
            }
        }
        SORT_PROBE(6);
        cluster.sync();   // all remote writes of this pass have landed (tot[] may be overwritten again)
        SORT_PROBE(7);
        uint32_t* tk = kc; kc = kn; kn = tk;
        uint16_t* tp = pc; pc = pn; pn = tp;
    }
    for (int i = threadIdx.x; i < n; i += THREADS) {
        skeys[(size_t)f * B + g0 + i] = (int32_t)kc[i] + off;
        perm[(size_t)f * B + g0 + i] = (int32_t)pc[i] * F + f;
    }
    SORT_PROBE(8);
}

}  // namespace

// debug hook (not in the public header)
FMB_API void fmb_debug_set_sort_buffer(long long* dev) { cudaMemcpyToSymbol(g_sort_dbg, &dev, sizeof(dev)); }

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

// host-side handles of the three per-field sort kernels (argument 0 = ids, 4 = sorted_keys, 5 = perm; 6 or 7 arguments),
// for graph-node identification in session.cu
FMB_API int fmb_sort_fields_kernel_fns(const void** fns, int* nparams) {
    fns[0] = (const void*)sort_fields_kernel; nparams[0] = 6;
    fns[1] = (const void*)sort_fields_cluster_kernel<256, true>; nparams[1] = 7;
    fns[2] = (const void*)sort_fields_cluster_kernel<1024, false>; nparams[2] = 7;
    return 3;
}

// Largest batch the per-field shared-memory sort accepts (cluster of 4 CTAs, 16-bit sample payload).
FMB_API int fmb_sort_fields_max_batch(void) { return 65536; }

// Stable sort of ids[B,F] (column f holds global row ids of field f, field_off[F+1] device array of
// field offsets) -> sorted_keys[B*F], perm[B*F] (original entry index b*F+f), identical to
// fmb_sort_segment over the flattened matrix.  Everything stays in shared memory: one CTA per field for
// B <= 2048, otherwise a 4-CTA thread-block cluster per field exchanging through distributed shared memory.
FMB_API int fmb_sort_fields(const int32_t* ids, int B, int F, const int32_t* field_off, int32_t* sorted_keys,
                            int32_t* perm, cudaStream_t stream) {
    FMB_CHECK_ARG(ids && field_off && sorted_keys && perm, "fmb_sort_fields: null pointer");
    FMB_CHECK_ARG(B > 0 && B <= fmb_sort_fields_max_batch() && F > 0, "fmb_sort_fields: B=%d out of range", B);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(sort_fields_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(sort_fields_cluster_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(sort_fields_cluster_kernel<1024, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        attr = true;
    }
    if (B <= 2048) {
        sort_fields_kernel<<<F, FS_THREADS, fmb::smem_sort_bytes(B), stream>>>(ids, B, F, field_off, sorted_keys, perm);
        FMB_CHECK_LAUNCH("sort_fields_kernel");
    } else {
        int Bq = (B + CL - 1) / CL;
        Bq = (Bq + 31) / 32 * 32;
        if (Bq <= 256 * FS_MAX_SLOTS)
            sort_fields_cluster_kernel<256, true><<<F * CL, 256, fmb::smem_sort_bytes(Bq) + (size_t)Bq * 6 + 16, stream>>>(ids, B, F, field_off,
                                                                                             sorted_keys, perm, Bq);
        else
            sort_fields_cluster_kernel<1024, false><<<F * CL, 1024, fmb::smem_sort_bytes(Bq), stream>>>(ids, B, F, field_off,
                                                                                               sorted_keys, perm, Bq);
        FMB_CHECK_LAUNCH("sort_fields_cluster_kernel");
    }
    return FMB_OK;
}
