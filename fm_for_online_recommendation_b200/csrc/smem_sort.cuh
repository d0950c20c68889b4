// smem_sort.cuh -- stable LSD radix sort of up to 16384 (key, 16-bit payload) pairs held in the
// shared memory of one 1024-thread CTA (used by the per-field sorts of radix_sort.cu and sharded.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fmb {

constexpr int SS_RADIX_BITS = 8;
constexpr int SS_RADIX = 1 << SS_RADIX_BITS;
constexpr int SS_THREADS = 1024;
constexpr int SS_WARPS = SS_THREADS / 32;
constexpr int SS_MAX_SLOTS = 16;  // n <= 32 warps * 16 slots * 32 lanes = 16384

// lanes of the warp whose (valid) key has the same digit as mine: what __match_any_sync returns, built
// from one ballot per digit bit (MATCH is far slower than 9 VOTEs on sm_100: it was 46 % of the
// sort kernel's stall samples in profiles/r1c).
__device__ __forceinline__ unsigned digit_peers(unsigned d, bool valid) {
    unsigned m = __ballot_sync(0xffffffffu, valid);
    if (!valid) m = ~m;
#pragma unroll
    for (int b = 0; b < SS_RADIX_BITS; ++b) {
        const unsigned bit = (d >> b) & 1u;
        const unsigned bal = __ballot_sync(0xffffffffu, bit);
        m &= bit ? bal : ~bal;
    }
    return m;
}

// shared-memory bytes for n pairs
__host__ __device__ inline size_t smem_sort_bytes(int n) {
    return (size_t)2 * n * 4 + (size_t)2 * (n + (n & 1)) * 2 + (size_t)SS_WARPS * SS_RADIX * 2;
}

// Sorts n pairs (kbuf0/pbuf0 hold the input) by the low `passes`*8 bits of the key, stable.
// On return *kout/*pout point at the buffers holding the result.  All SS_THREADS threads must call.
__device__ __forceinline__ void smem_sort_passes(uint32_t* kbuf0, uint32_t* kbuf1, uint16_t* pbuf0, uint16_t* pbuf1,
                                                 uint16_t* cnt, uint32_t* tot /*[256] shared*/, int n, int passes,
                                                 uint32_t** kout, uint16_t** pout) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int B = n;
    constexpr int RADIX = SS_RADIX, RADIX_BITS = SS_RADIX_BITS, FS_THREADS = SS_THREADS, FS_WARPS = SS_WARPS,
                  FS_MAX_SLOTS = SS_MAX_SLOTS;
    uint32_t* kc = kbuf0; uint32_t* kn = kbuf1;
    uint16_t* pc = pbuf0; uint16_t* pn = pbuf1;
    const int slots = (B + FS_THREADS - 1) / FS_THREADS;  // per-warp slice = slots*32 consecutive keys
    const int wbase = warp * slots * 32;
    const uint32_t lt = (1u << lane) - 1u;
    for (int ps = 0; ps < passes; ++ps) {
        const int shift = ps * RADIX_BITS;
        for (int i = threadIdx.x; i < FS_WARPS * RADIX / 2; i += FS_THREADS) reinterpret_cast<uint32_t*>(cnt)[i] = 0;
        __syncthreads();
        uint32_t key[FS_MAX_SLOTS];
        uint16_t rank[FS_MAX_SLOTS];
#pragma unroll
        for (int s = 0; s < FS_MAX_SLOTS; ++s) {
            if (s < slots) {
                const int idx = wbase + s * 32 + lane;
                const bool valid = idx < B;
                key[s] = valid ? kc[idx] : 0u;
                const unsigned d = valid ? ((key[s] >> shift) & (RADIX - 1)) : 0u;
                const unsigned m = digit_peers(d, valid);
                const int leader = __ffs(m) - 1;
                uint32_t old = 0;
                if (valid && lane == leader) { old = cnt[warp * RADIX + d]; cnt[warp * RADIX + d] = (uint16_t)(old + __popc(m)); }
                old = __shfl_sync(0xffffffffu, old, leader);
                rank[s] = (uint16_t)(old + __popc(m & lt));
                __syncwarp();
            }
        }
        __syncthreads();
        if (threadIdx.x < RADIX) {
            const int d = threadIdx.x;
            uint32_t run = 0;
            for (int w = 0; w < FS_WARPS; ++w) { const uint32_t t = cnt[w * RADIX + d]; cnt[w * RADIX + d] = (uint16_t)run; run += t; }
            tot[d] = run;
        }
        __syncthreads();
        if (warp == 0) {  // exclusive scan of the 256 digit totals, 8 per lane
            uint32_t v[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { v[j] = tot[lane * 8 + j]; sum += v[j]; }
            uint32_t inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            uint32_t ex = inc - sum;
#pragma unroll
            for (int j = 0; j < 8; ++j) { tot[lane * 8 + j] = ex; ex += v[j]; }
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < FS_MAX_SLOTS; ++s) {
            if (s < slots) {
                const int idx = wbase + s * 32 + lane;
                if (idx < B) {
                    const unsigned d = (key[s] >> shift) & (RADIX - 1);
                    const uint32_t pos = tot[d] + cnt[warp * RADIX + d] + rank[s];
                    kn[pos] = key[s];
                    pn[pos] = pc[idx];
                }
            }
        }
        __syncthreads();
        uint32_t* tk = kc; kc = kn; kn = tk;
        uint16_t* tp = pc; pc = pn; pn = tp;
    }
    *kout = kc;
    *pout = pc;
}

// Run list of a sorted key array held in shared memory (what fm_bwd_runs_list_kernel consumes; see fmb_runlist_t in
// include/fmb200.h): kc[0..n) sorted keys, position of entry i = pos0 + i, listed key = kc[i] + key_add.  Runs of >= 2 equal
// keys are written to `seg` (a segment of seg_cap entries): those shorter than 128 entries upwards from its start, the longer
// ones downwards from its end; *n_short / *n_long receive their numbers.  Counts per warp in registers, two passes, no
// atomics.  All THREADS threads of the CTA must call (one __syncthreads inside).
template <int THREADS>
__device__ __forceinline__ void runlist_from_sorted(const uint32_t* kc, int n, int pos0, int32_t key_add, int4* seg, int seg_cap,
                                                    uint32_t* n_short, uint32_t* n_long) {
    __shared__ uint32_t rl_ws[32], rl_wl[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const int nr = (n + 31) & ~31;
    uint32_t wc = 0, wcl = 0;
    for (int i = threadIdx.x; i < nr; i += THREADS) {
        bool start = false, lng = false;
        if (i < n) {
            const uint32_t key = kc[i];
            start = (i == 0 || kc[i - 1] != key) && i + 1 < n && kc[i + 1] == key;
            lng = start && i + 127 < n && kc[i + 127] == key;
        }
        wc += __popc(__ballot_sync(0xffffffffu, start && !lng));
        wcl += __popc(__ballot_sync(0xffffffffu, lng));
    }
    if (lane == 0) { rl_ws[warp] = wc; rl_wl[warp] = wcl; }
    __syncthreads();
    uint32_t base = 0, total = 0, basel = 0, totall = 0;
    for (int w = 0; w < THREADS / 32; ++w) {
        const uint32_t t = rl_ws[w], tl = rl_wl[w];
        if (w < warp) { base += t; basel += tl; }
        total += t; totall += tl;
    }
    if (threadIdx.x == 0) { *n_short = total; *n_long = totall; }
    if (wc + wcl == 0) return;
    for (int i = threadIdx.x; i < nr; i += THREADS) {
        uint32_t key = 0xfffffffeu;
        bool cont = false, start = false, lng = false;
        if (i < n) {
            key = kc[i];
            cont = i > 0 && kc[i - 1] == key;
            start = !cont && i + 1 < n && kc[i + 1] == key;
            lng = start && i + 127 < n && kc[i + 127] == key;
        }
        const unsigned m = __ballot_sync(0xffffffffu, start);
        if (m) {
            const unsigned ml = __ballot_sync(0xffffffffu, lng);
            const unsigned cm = __ballot_sync(0xffffffffu, cont);
            const uint32_t klast = __shfl_sync(0xffffffffu, key, 31);
            const uint32_t k2 = i + 32 < n ? kc[i + 32] : 0xffffffffu;
            const unsigned em = __ballot_sync(0xffffffffu, k2 == klast);
            const int ext = em == 0xffffffffu ? 32 : __ffs(~em) - 1;
            if (start) {
                const unsigned after = lane == 31 ? 0u : (cm >> (lane + 1));
                const int inw = lane == 31 ? 0 : __ffs(~after) - 1;
                int n0 = 1 + inw;
                if (lane + inw == 31) n0 += ext;
                if (n0 > 32) n0 = 32;
                const int4 e4 = make_int4(pos0 + i, (int32_t)key + key_add, n0, lng ? 1 : 0);
                if (lng) seg[seg_cap - 1 - (int)(basel + __popc(ml & lt))] = e4;
                else seg[base + __popc((m & ~ml) & lt)] = e4;
            }
            base += __popc(m & ~ml);
            basel += __popc(ml);
        }
    }
}

}  // namespace fmb
