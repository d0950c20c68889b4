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

}  // namespace fmb
