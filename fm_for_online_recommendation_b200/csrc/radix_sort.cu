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
constexpr int CL_SEGS = 4;   // run-list segments per field = CTAs of a sort cluster
__device__ long long* g_sort_dbg = nullptr;   // debug: per-phase cycle counts of one CTA + per-field globaltimer stamps
__device__ __forceinline__ void sort_stamp(int slot) {
    if (g_sort_dbg && threadIdx.x == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); g_sort_dbg[slot] = (long long)gt; }
}
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

// Exclusive scan of the [RADIX][ntiles] histogram in two levels: CTA d scans row d in place (exclusive over the tiles)
// and writes the row's total to rowtot[d]; the scatter kernel adds the exclusive prefix of the 256 row totals itself.
// (One CTA scanning all 256 * ntiles counters serially was 470 us per pass at B = 262 144: the whole sort's time.)
__global__ void __launch_bounds__(256) radix_scan_rows_kernel(uint32_t* __restrict__ hist, int ntiles, uint32_t* __restrict__ rowtot) {
    __shared__ uint32_t warp_tot[8];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* row = hist + (size_t)blockIdx.x * ntiles;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    constexpr int IPT = 4;
    for (int base = 0; base < ntiles; base += 256 * IPT) {
        const int i0 = base + threadIdx.x * IPT;
        uint32_t v[IPT], sum = 0;
#pragma unroll
        for (int j = 0; j < IPT; ++j) { v[j] = (i0 + j < ntiles) ? row[i0 + j] : 0u; sum += v[j]; }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        uint32_t before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const uint32_t t = warp_tot[w]; if (w < warp) before += t; all += t; }
        const uint32_t carry = carry_s;
        uint32_t excl = carry + before + (inc - sum);
#pragma unroll
        for (int j = 0; j < IPT; ++j) { if (i0 + j < ntiles) row[i0 + j] = excl; excl += v[j]; }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + all;
        __syncthreads();
    }
    if (threadIdx.x == 0) rowtot[blockIdx.x] = carry_s;
}

// stable scatter of one tile. vals_in == nullptr means "identity payload" (first pass).
__global__ void __launch_bounds__(SORT_THREADS) radix_scatter_kernel(const int32_t* __restrict__ keys_in,
                                                                     const int32_t* __restrict__ vals_in, int64_t N,
                                                                     int shift, int ntiles,
                                                                     const uint32_t* __restrict__ scan,
                                                                     const uint32_t* __restrict__ rowtot,
                                                                     int32_t* __restrict__ keys_out,
                                                                     int32_t* __restrict__ vals_out) {
    __shared__ uint32_t cnt[SORT_WARPS][RADIX];
    __shared__ uint32_t gbase[RADIX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&cnt[0][0])[i] = 0;
    for (int i = threadIdx.x; i < RADIX; i += SORT_THREADS) gbase[i] = scan[(size_t)i * ntiles + blockIdx.x];
    __syncthreads();
    if (warp == 0) {   // + exclusive prefix of the 256 digit totals (8 per lane)
        uint32_t v[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[j] = rowtot[lane * 8 + j]; sum += v[j]; }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        uint32_t ex = inc - sum;
#pragma unroll
        for (int j = 0; j < 8; ++j) { gbase[lane * 8 + j] += ex; ex += v[j]; }
    }
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
// Tail of the per-field sorts: the CTA holds positions [g0, g0 + n) of field f's sorted order in shared memory (kc = row
// ids relative to the field, pc = sample index).  Besides sorted_keys / perm it writes what the FM step needs one step
// later (fm_step.cu, fm_backward.cu) -- this used to be a separate kernel (pos_flags_kernel, 7-8 us behind the sort):
//   posflag[entry] = sorted position | 0x80000000 if the entry's row is hit more than once in the batch
//   run list       = {first sorted position, key, number of the run's entries among its first 32 positions, 0} for every
//                    run of >= 2 entries that STARTS in this CTA, appended in any order to the CTA's own segment
//                    (rl_entries + seg * rl_cap); rl_segc[seg] = their number.  No global counter, nothing to zero.
// key_at(gi) returns the key at field position gi when it is held by another CTA of the cluster.
// ---------------------------------------------------------------------------------------------
template <int THREADS, typename KeyAt>
__device__ __forceinline__ void sort_tail(const uint32_t* kc, const uint16_t* pc, int n, int g0, int B, int F, int f,
                                          int32_t off, int32_t* __restrict__ skeys, int32_t* __restrict__ perm,
                                          uint32_t* __restrict__ posflag, int4* __restrict__ rl_entries,
                                          uint32_t* __restrict__ rl_segc, int rl_cap, int rl_nseg, int seg, int seg_long, KeyAt key_at) {
    __shared__ uint32_t wcount_s[32], wcountl_s[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const int nr = (n + 31) & ~31;
    int4* mine = rl_entries ? rl_entries + (size_t)seg * rl_cap : nullptr;
    int4* minel = rl_entries ? rl_entries + (size_t)seg_long * rl_cap : nullptr;   // long runs: downwards from this segment's end
    // long run (>= 128 entries) starting at local position i?  (its own launch of the run kernel)
    auto is_long = [&](int i, uint32_t key) {
        const int g2 = g0 + i + 127, j2 = i + 127;
        return g2 < B && (j2 < n ? kc[j2] : key_at(g2)) == key;
    };
    // pass 1: outputs per position; run starts are only counted (per warp, in registers)
    uint32_t wc = 0, wcl = 0;
    for (int i = threadIdx.x; i < nr; i += THREADS) {     // a warp covers 32 consecutive sorted positions per iteration
        bool start = false, lng = false;
        if (i < n) {
            const int gi = g0 + i;
            const uint32_t key = kc[i];
            const uint32_t prev = gi > 0 ? (i > 0 ? kc[i - 1] : key_at(gi - 1)) : 0xffffffffu;
            const uint32_t next = gi + 1 < B ? (i + 1 < n ? kc[i + 1] : key_at(gi + 1)) : 0xffffffffu;
            const bool cont = key == prev, multi = cont || key == next;
            start = multi && !cont;
            if (start && mine) lng = is_long(i, key);
            const size_t o = (size_t)f * B + gi;
            const int32_t e = (int32_t)pc[i] * F + f;
            skeys[o] = (int32_t)key + off;
            perm[o] = e;
            if (posflag) posflag[e] = (uint32_t)o | (multi ? 0x80000000u : 0u);
        }
        if (mine) {
            wc += __popc(__ballot_sync(0xffffffffu, start && !lng));
            wcl += __popc(__ballot_sync(0xffffffffu, lng));
        }
    }
    if (!mine) return;
    if (lane == 0) { wcount_s[warp] = wc; wcountl_s[warp] = wcl; }
    __syncthreads();
    uint32_t base = 0, total = 0, basel = 0, totall = 0;
    for (int w = 0; w < THREADS / 32; ++w) {
        const uint32_t t = wcount_s[w], tl = wcountl_s[w];
        if (w < warp) { base += t; basel += tl; }
        total += t; totall += tl;
    }
    if (threadIdx.x == 0) { rl_segc[seg] = total; rl_segc[rl_nseg + seg_long] = totall; }
    if (wc + wcl == 0) return;
    // pass 2 (warps that hold run starts): the entries -- short runs from the segment's start upwards, long runs from its
    // end downwards.  No counter in shared memory: same-address ATOMS.ADD with a return value retire one at a time.
    for (int i = threadIdx.x; i < nr; i += THREADS) {
        uint32_t key = 0xfffffffeu;
        bool cont = false, start = false, lng = false;
        const int gi = g0 + i;
        if (i < n) {
            key = kc[i];
            const uint32_t prev = gi > 0 ? (i > 0 ? kc[i - 1] : key_at(gi - 1)) : 0xffffffffu;
            const uint32_t next = gi + 1 < B ? (i + 1 < n ? kc[i + 1] : key_at(gi + 1)) : 0xffffffffu;
            cont = key == prev;
            start = !cont && key == next;
            if (start) lng = is_long(i, key);
        }
        const unsigned m = __ballot_sync(0xffffffffu, start);
        if (m) {
            // entries of each run among its first 32 positions, without a loop: inside this warp's window the run of a
            // start at lane l continues over the lanes whose key equals their predecessor's; only the LAST run of the
            // window can leave it, and the next 32 keys tell how far (a per-start scan of up to 31 shared-memory reads
            // cost 4-20 us per field here)
            const unsigned ml = __ballot_sync(0xffffffffu, lng);
            const unsigned cm = __ballot_sync(0xffffffffu, cont);
            const uint32_t klast = __shfl_sync(0xffffffffu, key, 31);
            const int i2 = i + 32, gi2 = gi + 32;
            const uint32_t k2 = gi2 < B ? (i2 < n ? kc[i2] : key_at(gi2)) : 0xffffffffu;
            const unsigned em = __ballot_sync(0xffffffffu, k2 == klast);
            const int ext = em == 0xffffffffu ? 32 : __ffs(~em) - 1;            // run of the last lane, beyond the window
            if (start) {
                const unsigned after = lane == 31 ? 0u : (cm >> (lane + 1));     // continuation flags of the lanes after me
                const int inw = lane == 31 ? 0 : __ffs(~after) - 1;              // consecutive ones (bit 31 - lane is 0)
                int n0 = 1 + inw;
                if (lane + inw == 31) n0 += ext;                                 // my run reaches the window's last lane
                if (n0 > 32) n0 = 32;
                const int4 e4 = make_int4((int)((size_t)f * B + gi), (int32_t)key + off, n0, lng ? 1 : 0);
                if (lng) minel[rl_cap - 1 - (int)(basel + __popc(ml & lt))] = e4;
                else mine[base + __popc((m & ~ml) & lt)] = e4;
            }
            base += __popc(m & ~ml);
            basel += __popc(ml);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Sparse fields (far more rows than the batch has samples: Criteo's large categorical fields, 1.27 M rows against 8 192
// samples) need no sort at all for the FM step: what the step consumes is (a) per entry, whether its row is hit more than
// once, and (b) for those entries only, a position such that the entries of one row are contiguous and in sample order.
// Rows hit once are updated inside the fused kernel and their position is never used.  One CTA per sparse field:
//   1. the field's keys go into a shared-memory hash table (open addressing, rounds of plain stores and barriers, no
//      atomics); an entry that finds its key already present marks itself and the slot's owner as multi-hit -- the SET of
//      marked entries does not depend on who won a slot;
//   2. the marked entries are compacted in sample order (block scan) -- ~50 of 8 192 on uniform ids;
//   3. their rank by (key, sample) is counted directly (n^2 comparisons, n <= 256) or, on skewed ids where many entries
//      share rows, found by the shared-memory radix sort of smem_sort.cuh over the compacted list;
//   4. outputs: posflag of the multi-hit entries (the caller cleared the buffer), sorted_keys / perm of the multi-hit
//      entries at the head of the field's range [f*B, f*B + nm), sorted_keys = -1 behind them (the run kernel probes keys
//      behind a run's end), the run list.
// Result for the consumers = that of the full sort restricted to multi-hit entries, at 1/5 of its time and without the
// register footprint that kept the fused kernel's tiles off the SMs.
// ---------------------------------------------------------------------------------------------
constexpr int SPARSE_THREADS = 1024;
constexpr int SPARSE_DIRECT = 256;     // largest multi-hit count ranked by direct comparison

__host__ __device__ inline int sparse_table_slots(int B) { int t = 1024; while (t < 2 * B) t <<= 1; return t; }
// keys [B] u32 | flag [B] u8 | compacted keys [B] u32 | compacted samples [B] u16 | region shared by the hash table (step 1)
// and the second key/sample buffers + radix counters (step 3)
__host__ __device__ inline size_t sparse_region_bytes(int B) {
    const size_t a = (size_t)sparse_table_slots(B) * 4, b = (size_t)B * 4 + (size_t)(B + (B & 1)) * 2 + 32 * 256 * 2;
    return a > b ? a : b;
}
// (the two compacted buffers also hold the per-warp lists of unsettled entries during the insertion: sized for >= 2 048)
__host__ __device__ inline int sparse_bp(int B) { return B < 2048 ? 2048 : B; }
__host__ __device__ inline size_t sparse_smem_bytes(int B) {
    return (size_t)B * 4 + (size_t)((B + 3) & ~3) + (size_t)sparse_bp(B) * 4 + (size_t)((sparse_bp(B) + 1) & ~1) * 2 + sparse_region_bytes(B) + 16;
}

__global__ void __launch_bounds__(SPARSE_THREADS) sparse_fields_kernel(const int32_t* __restrict__ ids, int B, int F,
                                                                      const int32_t* __restrict__ field_off,
                                                                      int64_t min_rows, int32_t* __restrict__ skeys,
                                                                      int32_t* __restrict__ perm, uint32_t* __restrict__ posflag,
                                                                      int4* __restrict__ rl_entries, uint32_t* __restrict__ rl_segc,
                                                                      int rl_cap) {
    const int f = blockIdx.x;
    const int32_t off = field_off[f];
    const uint32_t nrows = (uint32_t)(field_off[f + 1] - off);
    if ((int64_t)nrows < min_rows) return;                 // dense field: the radix kernels own it
    extern __shared__ __align__(16) unsigned char sp_smem[];
    const int slots = sparse_table_slots(B);
    uint32_t* keys = reinterpret_cast<uint32_t*>(sp_smem);             // [B]
    uint8_t* flag = reinterpret_cast<uint8_t*>(keys + B);              // [B]
    const int Bp = sparse_bp(B);
    uint32_t* kbuf0 = reinterpret_cast<uint32_t*>(flag + ((B + 3) & ~3));   // compacted multi-hit keys (sample order)
    uint16_t* pbuf0 = reinterpret_cast<uint16_t*>(kbuf0 + Bp);         // their samples
    uint32_t* tab = reinterpret_cast<uint32_t*>(pbuf0 + ((Bp + 1) & ~1));   // [slots] owner entry + 1, 0 = empty (step 1)
    uint32_t* kbuf1 = tab;                                             // step 3 reuses the table's memory
    uint16_t* pbuf1 = reinterpret_cast<uint16_t*>(kbuf1 + B);
    uint16_t* cnt = pbuf1 + B + (B & 1);                               // radix counters (fallback only)
    __shared__ uint32_t tot[256];
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t nm_s, t_fill, t_filll;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    sort_stamp(16 + 2 * F + 2 * f);
    for (int i0 = threadIdx.x; i0 < B; i0 += 8 * SPARSE_THREADS) {   // eight strided loads in flight per thread
        int32_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = i0 + u * SPARSE_THREADS; v[u] = i < B ? __ldg(ids + (size_t)i * F + f) : 0; }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * SPARSE_THREADS;
            if (i < B) { keys[i] = (uint32_t)(v[u] - off); flag[i] = 0; }
        }
    }
    for (int i = threadIdx.x; i < slots; i += SPARSE_THREADS) tab[i] = 0;
    if (threadIdx.x == 0) { t_fill = 0; t_filll = 0; }
    __syncthreads();
    sort_stamp(16 + 6 * F + f);   // keys loaded, table cleared
    // 1. insert, without atomics (8 192 ATOMS.CAS on one SM took 9 us: shared-memory atomics retire a few lanes per
    //    clock).  Rounds of "store, barrier, look": an unresolved entry whose probe slot is empty stores its own index
    //    there (plain store; among racing writers one survives), and after the barrier reads the slot back: its own
    //    index = it owns the slot; an entry with the same key = both are multi-hit; another key = next probe slot.
    //    Entries with equal keys walk the same probe sequence in the same rounds and see the same table, so they always
    //    meet in one slot; the set of marked entries does not depend on who wins a race.
    {
        constexpr int PER = 8;                     // B <= 8 * SPARSE_THREADS is guaranteed by sparse_smem_bytes
        const uint32_t mask = (uint32_t)slots - 1u;
        const int hshift = 32 - (31 - __clz(slots));
        uint32_t key[PER], slot[PER], step[PER];
        unsigned open = 0;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int i = threadIdx.x + u * SPARSE_THREADS;
            key[u] = 0; slot[u] = 0; step[u] = 1;
            if (i < B) {
                key[u] = keys[i];
                slot[u] = (key[u] * 2654435761u) >> hshift;
                step[u] = ((key[u] * 0x85ebca6bu) >> 17) | 1u;     // odd: the probe sequence visits every slot
                open |= 1u << u;
            }
        }
        // round 0: every entry (three quarters are settled here at load factor 1/2)
#pragma unroll
        for (int u = 0; u < PER; ++u)
            if ((open >> u) & 1u) { if (tab[slot[u]] == 0u) tab[slot[u]] = (uint32_t)(threadIdx.x + u * SPARSE_THREADS) + 1u; }
        __syncthreads();
        // unsettled entries and their current probe slot, one private list per warp (a warp owns 8 x 32 entries): no
        // counter in shared memory -- 256 ATOMS.ADD on one address took 12 us here
        const int perb = (B + SPARSE_THREADS - 1) / SPARSE_THREADS;   // entries per thread: a warp owns 32 * perb
        uint16_t* ulist = reinterpret_cast<uint16_t*>(kbuf0) + warp * (perb * 32);
        uint16_t* uslot = ulist + SPARSE_THREADS * perb;           // (kbuf0 | pbuf0 are free until the compaction of step 2:
                                                                   //  3 * Bp halfwords >= 2 * (B + 1023))
        int wn = 0;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            bool unsettled = false;
            const uint32_t me = (uint32_t)(threadIdx.x + u * SPARSE_THREADS) + 1u;
            if ((open >> u) & 1u) {
                const uint32_t cur = tab[slot[u]];
                if (cur == me) { }
                else if (keys[cur - 1] == key[u]) { flag[cur - 1] = 1; flag[me - 1] = 1; }
                else unsettled = true;
            }
            const unsigned um = __ballot_sync(0xffffffffu, unsettled);
            if (unsettled) {
                const int j = wn + __popc(um & lt);
                ulist[j] = (uint16_t)(me - 1u);
                uslot[j] = (uint16_t)((slot[u] + step[u]) & mask);
            }
            wn += __popc(um);
        }
        __syncthreads();
        // the unsettled quarter: classic atomicCAS probing from their next slot (equal keys walk the same sequence, so
        // the second one meets the first).  Doing ALL entries this way took 9.5 us (ATOMS retire ~1 per ns per SM),
        // doing all later rounds with plain stores and two barriers each took 12.
        for (int j = lane; j < wn; j += 32) {
            const uint32_t i = ulist[j], k = keys[i];
            const uint32_t st = ((k * 0x85ebca6bu) >> 17) | 1u;
            uint32_t sl = uslot[j];
            for (;;) {
                const uint32_t old = atomicCAS(&tab[sl], 0u, i + 1u);
                if (old == 0u) break;
                if (keys[old - 1] == k) { flag[old - 1] = 1; flag[i] = 1; break; }
                sl = (sl + st) & mask;
            }
        }
    }
    __syncthreads();
    sort_stamp(16 + 7 * F + f);   // inserted
    // 2. compact the marked entries in sample order: thread t owns entries [t*per, (t+1)*per)
    const int per = (B + SPARSE_THREADS - 1) / SPARSE_THREADS;
    const int i0 = threadIdx.x * per;
    uint32_t mine = 0;
    for (int u = 0; u < per; ++u) if (i0 + u < B) mine += flag[i0 + u];
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = wsum[lane];
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
        wsum[lane] = winc - w;
        if (lane == 31) nm_s = winc;
    }
    __syncthreads();
    {
        uint32_t o = wsum[warp] + inc - mine;
        for (int u = 0; u < per; ++u)
            if (i0 + u < B && flag[i0 + u]) { kbuf0[o] = keys[i0 + u]; pbuf0[o] = (uint16_t)(i0 + u); ++o; }
    }
    __syncthreads();
    const int nm = (int)nm_s;
    sort_stamp(16 + 5 * F + f);   // hash + compaction done
    // 3. order by (key, sample).  The compacted list is in sample order, so a stable sort by key is what is needed.
    uint32_t* ks = kbuf1;     // sorted keys
    uint16_t* ps = pbuf1;     // their samples
    if (nm <= SPARSE_DIRECT) {
        for (int a = threadIdx.x; a < nm; a += SPARSE_THREADS) {
            const uint32_t key = kbuf0[a];
            int r = 0;
            for (int b2 = 0; b2 < nm; ++b2) { const uint32_t kb = kbuf0[b2]; r += (kb < key) || (kb == key && b2 < a); }
            kbuf1[r] = key;
            pbuf1[r] = pbuf0[a];
        }
        __syncthreads();
    } else {
        const int bits = nrows <= 1 ? 0 : 32 - __clz(nrows - 1);
        fmb::smem_sort_passes(kbuf0, kbuf1, pbuf0, pbuf1, cnt, tot, nm, (bits + RADIX_BITS - 1) / RADIX_BITS, &ks, &ps);
    }
    // 4. outputs
    sort_stamp(16 + 8 * F + f);   // ranked
    const size_t fb = (size_t)f * B;
    // posflag of the rows hit once stays 0 (the caller cleared the buffer: 8 192 scattered 4-byte stores per field from
    // one SM took 4 us, a 1.3 MB memset takes 1); perm is only defined for the placed entries
    for (int i = nm + threadIdx.x; i < B; i += SPARSE_THREADS) skeys[fb + i] = -1;
    int4* seg = rl_entries ? rl_entries + (size_t)(f * CL_SEGS) * rl_cap : nullptr;
    const int nr = (nm + 31) & ~31;
    for (int a = threadIdx.x; a < nr; a += SPARSE_THREADS) {
        bool start = false;
        uint32_t key = 0;
        if (a < nm) {
            key = ks[a];
            const int32_t e = (int32_t)ps[a] * F + f;
            skeys[fb + a] = (int32_t)key + off;
            perm[fb + a] = e;
            posflag[e] = (uint32_t)(fb + a) | 0x80000000u;
            start = a == 0 || ks[a - 1] != key;
        }
        if (seg && start) {        // a few dozen runs per field: plain shared-memory counters
            int n0 = 1;
            while (n0 < 32 && a + n0 < nm && ks[a + n0] == key) ++n0;
            const bool lng = a + 127 < nm && ks[a + 127] == key;
            const int4 e4 = make_int4((int)(fb + a), (int32_t)key + off, n0, lng ? 1 : 0);
            if (lng) seg[CL_SEGS * rl_cap - 1 - (int)atomicAdd(&t_filll, 1u)] = e4;   // the field's four segments are contiguous
            else seg[atomicAdd(&t_fill, 1u)] = e4;
        }
    }
    if (seg) {
        __syncthreads();
        // short runs counted in the field's first segment, long ones (stored downwards from the end of the field's
        // four contiguous segments) in its last
        if (threadIdx.x < CL_SEGS) {
            rl_segc[f * CL_SEGS + threadIdx.x] = threadIdx.x == 0 ? t_fill : 0u;
            rl_segc[F * CL_SEGS + f * CL_SEGS + threadIdx.x] = threadIdx.x == CL_SEGS - 1 ? t_filll : 0u;
        }
    }
    sort_stamp(17 + 2 * F + 2 * f);
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
                                                                 int32_t* __restrict__ perm, uint32_t* __restrict__ posflag,
                                                                 int4* __restrict__ rl_entries, uint32_t* __restrict__ rl_segc,
                                                                 int rl_cap, int64_t max_rows) {
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
    if ((int64_t)nrows >= max_rows) return;      // sparse field: sparse_fields_kernel owns it
    const int bits = nrows <= 1 ? 0 : 32 - __clz(nrows - 1);
    const int passes = (bits + RADIX_BITS - 1) / RADIX_BITS;
    (void)lane; (void)warp;
    for (int i = threadIdx.x; i < B; i += FS_THREADS) {
        kbuf0[i] = (uint32_t)(ids[(size_t)i * F + f] - off);
        pbuf0[i] = (uint16_t)i;
    }
    __syncthreads();
    uint32_t* kc; uint16_t* pc;
    fmb::smem_sort_passes(kbuf0, kbuf1, pbuf0, pbuf1, cnt, tot, B, passes, &kc, &pc);
    // one CTA owns the field's four (contiguous) segments: short runs upwards from the first, long runs downwards from
    // the end of the last
    sort_tail<FS_THREADS>(kc, pc, B, 0, B, F, f, off, skeys, perm, posflag, rl_entries, rl_segc, rl_cap, F * CL_SEGS, f * CL_SEGS,
                          f * CL_SEGS + CL_SEGS - 1, [](int) { return 0xffffffffu; });
    if (rl_entries && threadIdx.x >= 1 && threadIdx.x < CL_SEGS) {
        rl_segc[f * CL_SEGS + threadIdx.x] = 0u;
        rl_segc[F * CL_SEGS + f * CL_SEGS + threadIdx.x - 1] = 0u;
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
constexpr int CL = CL_SEGS;
#define SORT_PROBE(slot) do { if (dbg && threadIdx.x == 0) { long long t_ = clock64(); dbg[slot] += t_ - tlast; tlast = t_; } } while (0)

// THREADS = 256 for batches up to 16384 (8 warps x <= 16 slots per CTA: the per-pass fixed work -- counter
// reset, per-digit prefix over warps -- scales with the warp count and dominated the 1024-thread version),
// 1024 above that.
// STAGED: keys are first scattered into a LOCAL staging buffer in digit order, then copied to their destination
// CTAs with lane-contiguous (coalesced) distributed-shared-memory stores; the direct version issued two
// scattered 4-/2-byte remote stores per key, which is what bounded it (profiles/r1g).
// MAXS: register slots per thread (keys of one CTA / THREADS, rounded up): 8 covers batches up to 8 192 with 256 threads and
// keeps the kernel at <= 80 registers (3 CTAs per SM by registers instead of 2: its CTAs share the SMs with the fused kernel's tiles).
template <int THREADS, bool STAGED, int MAXS>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, THREADS == 256 && MAXS == 8 ? 3 : 1)
sort_fields_cluster_kernel(const int32_t* __restrict__ ids, int B, int F, const int32_t* __restrict__ field_off,
                           int32_t* __restrict__ skeys, int32_t* __restrict__ perm, int Bq, uint32_t* __restrict__ posflag,
                           int4* __restrict__ rl_entries, uint32_t* __restrict__ rl_segc, int rl_cap, int64_t max_rows) {
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
    if ((int64_t)nrows >= max_rows) return;               // sparse field (the whole cluster leaves): sparse_fields_kernel owns it
    const int bits = nrows <= 1 ? 0 : 32 - __clz(nrows - 1);
    const int passes = (bits + RADIX_BITS - 1) / RADIX_BITS;
    const int g0 = crank * Bq;                            // first field position held by this CTA
    const int n = max(0, min(Bq, B - g0));                // entries held by this CTA
    long long* dbg = (g_sort_dbg && blockIdx.x == (unsigned)(F - 1) * CL) ? g_sort_dbg : nullptr;
    long long tlast = clock64();
    if (crank == 0) sort_stamp(16 + 2 * f);
    if (g_sort_dbg && threadIdx.x == 0) {   // debug: earliest CTA start / latest CTA end of the launch (globaltimer ns)
        unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        atomicMin(reinterpret_cast<unsigned long long*>(g_sort_dbg) + 14, gt);
    }
    for (int i0 = threadIdx.x; i0 < n; i0 += 8 * THREADS) {   // eight strided loads in flight per thread
        int32_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = i0 + u * THREADS; v[u] = i < n ? __ldg(ids + (size_t)(g0 + i) * F + f) : 0; }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * THREADS;
            if (i < n) { kbuf0[i] = (uint32_t)(v[u] - off); pbuf0[i] = (uint16_t)(g0 + i); }
        }
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
        uint32_t key[MAXS];
        uint16_t rank[MAXS];
        unsigned peers[MAXS];
        // peer masks of all slots first: the ballots of different slots are independent, so their latencies
        // overlap (issued slot after slot they were 560 cycles per slot: 9 dependent VOTEs each)
#pragma unroll
        for (int s = 0; s < MAXS; ++s) {
            if (s < slots) {
                const int idx = wbase + s * 32 + lane;
                const bool valid = idx < n;
                key[s] = valid ? kc[idx] : 0u;
                peers[s] = digit_peers(valid ? ((key[s] >> shift) & (RADIX - 1)) : 0u, valid);
            }
        }
        // then the (short) serial part: per-warp digit counters
#pragma unroll
        for (int s = 0; s < MAXS; ++s) {
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
            for (int s = 0; s < MAXS; ++s) {
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
            for (int s = 0; s < MAXS; ++s) {
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
    if (crank == 0) sort_stamp(16 + 4 * F + f);   // passes done
    if (passes == 0) cluster.sync();   // the neighbours' keys are read below: their initial loads must have landed
    // every CTA holds the same buffer parity (same number of passes), so kc names the sorted keys in all of them
    sort_tail<THREADS>(kc, pc, n, g0, B, F, f, off, skeys, perm, posflag, rl_entries, rl_segc, rl_cap, F * CL_SEGS, f * CL_SEGS + crank, f * CL_SEGS + crank, [&](int gi) {
        const int dest = (gi >= Bq) + (gi >= 2 * Bq) + (gi >= 3 * Bq);
        return cluster.map_shared_rank(kc, dest)[gi - dest * Bq];
    });
    SORT_PROBE(8);
    cluster.sync();                    // no CTA leaves while a neighbour may still read its keys
    if (crank == 0) sort_stamp(17 + 2 * f);
    if (g_sort_dbg && threadIdx.x == 0) {
        unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        atomicMax(reinterpret_cast<unsigned long long*>(g_sort_dbg) + 15, gt);
        if (nrows > 100000u && crank == 0) atomicMin(reinterpret_cast<unsigned long long*>(g_sort_dbg) + 13, gt);   // first large field done
        if (nrows <= 256u) atomicMax(reinterpret_cast<unsigned long long*>(g_sort_dbg) + 12, gt);                    // last small field done
    }
}

}  // namespace

// debug hook (not in the public header)
FMB_API void fmb_debug_set_sort_buffer(long long* dev) { cudaMemcpyToSymbol(g_sort_dbg, &dev, sizeof(dev)); }

static int sort_ntiles(int64_t N) { return (int)((N + TILE - 1) / TILE); }

// workspace: ping-pong key/val buffers + histogram
FMB_API size_t fmb_sort_workspace_bytes(int64_t N) {
    const size_t n4 = ((size_t)N * 4 + 255) / 256 * 256;
    const size_t h = ((size_t)RADIX * (sort_ntiles(N) + 1) * 4 + 255) / 256 * 256;   // histogram + the 256 row totals
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
        radix_scan_rows_kernel<<<RADIX, 256, 0, stream>>>(hist, ntiles, hist + (size_t)RADIX * ntiles);
        radix_scatter_kernel<<<ntiles, SORT_THREADS, 0, stream>>>(kin, vin, N, shift, ntiles, hist, hist + (size_t)RADIX * ntiles, kout, vout);
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

// host-side handles of the per-field sort kernels (argument 0 = ids; sorted_keys = argument 4, or 5 for the sparse-field kernel),
// for graph-node identification in session.cu
FMB_API int fmb_sort_fields_kernel_fns(const void** fns, int* nparams, int* skeys_arg) {
    fns[0] = (const void*)sort_fields_kernel; nparams[0] = 11; skeys_arg[0] = 4;
    fns[1] = (const void*)sort_fields_cluster_kernel<256, true, 16>; nparams[1] = 12; skeys_arg[1] = 4;
    fns[2] = (const void*)sort_fields_cluster_kernel<1024, false, 16>; nparams[2] = 12; skeys_arg[2] = 4;
    fns[3] = (const void*)sparse_fields_kernel; nparams[3] = 11; skeys_arg[3] = 5;
    fns[4] = (const void*)sort_fields_cluster_kernel<256, true, 8>; nparams[4] = 12; skeys_arg[4] = 4;
    return 5;
}

// Largest batch the per-field shared-memory sort accepts (cluster of 4 CTAs, 16-bit sample payload).
FMB_API int fmb_sort_fields_max_batch(void) { return 65536; }

// Stable sort of ids[B,F] (column f holds global row ids of field f, field_off[F+1] device array of
// field offsets) -> sorted_keys[B*F], perm[B*F] (original entry index b*F+f), identical to
// fmb_sort_segment over the flattened matrix.  Everything stays in shared memory: one CTA per field for
// B <= 2048, otherwise a 4-CTA thread-block cluster per field exchanging through distributed shared memory.
// _ex: optional outputs for the FM step of fm_step.cu / fm_backward.cu (see sort_tail): posflag[B*F] and the run list rl
// (shape from fmb_runlist_shape; nothing to zero: every segment's count is written).  flags & FMB_SORT_SPARSE_OK: fields
// with at least 16*B rows may skip the sort (sparse_fields_kernel): sorted_keys / perm then hold only the entries of
// rows hit more than once, at the head of the field's range, and -1 behind them -- all the FM step reads.
struct fmb_runlist_t { int32_t* entries; uint32_t* seg_count; int nseg, seg_cap; };   // include/fmb200.h
#define FMB_SORT_SPARSE_OK 1

FMB_API void fmb_runlist_shape(int B, int F, int* nseg, int* seg_cap) {
    int Bq = (B + CL - 1) / CL;
    Bq = (Bq + 31) / 32 * 32;
    *nseg = F * CL_SEGS;
    *seg_cap = Bq;
}

// rows a field must have for FMB_SORT_SPARSE_OK to skip its sort at batch size B; 0 = never at this batch size
FMB_API int64_t fmb_sort_fields_sparse_min_rows(int B) {
    return (B > 0 && sparse_smem_bytes(B) <= 200 * 1024) ? (int64_t)16 * B : 0;
}

// flags & FMB_SORT_PART_DENSE / FMB_SORT_PART_SPARSE (internal, with FMB_SORT_SPARSE_OK): launch only the radix kernel of
// the dense fields / only the hash kernel of the sparse fields -- the two are independent, so a caller with two streams
// (session.cu) runs them side by side; neither bit = both, one after the other on `stream`.
#define FMB_SORT_PART_DENSE 2
#define FMB_SORT_PART_SPARSE 4
FMB_API int fmb_sort_fields_ex(const int32_t* ids, int B, int F, const int32_t* field_off, int32_t* sorted_keys,
                               int32_t* perm, uint32_t* posflag, const fmb_runlist_t* rl, int flags, cudaStream_t stream) {
    FMB_CHECK_ARG(ids && field_off && sorted_keys && perm, "fmb_sort_fields: null pointer");
    FMB_CHECK_ARG(B > 0 && B <= fmb_sort_fields_max_batch() && F > 0, "fmb_sort_fields: B=%d out of range", B);
    int Bq = (B + CL - 1) / CL;
    Bq = (Bq + 31) / 32 * 32;
    FMB_CHECK_ARG(!rl || (rl->entries && rl->seg_count && rl->nseg == F * CL_SEGS && rl->seg_cap >= Bq && F * CL_SEGS <= 2048),
                  "fmb_sort_fields: run list shape must be that of fmb_runlist_shape");
    int4* rle = rl ? reinterpret_cast<int4*>(rl->entries) : nullptr;
    uint32_t* rlc = rl ? rl->seg_count : nullptr;
    const int rcap = rl ? rl->seg_cap : 0;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(sort_fields_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(sort_fields_cluster_kernel<256, true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(sort_fields_cluster_kernel<256, true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(sort_fields_cluster_kernel<1024, false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(sparse_fields_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        attr = true;
    }
    // sparse fields: the hash kernel, when the caller allows the reduced contract and the batch fits its shared memory
    const bool sparse = (flags & FMB_SORT_SPARSE_OK) && posflag && fmb_sort_fields_sparse_min_rows(B) > 0;
    const int64_t split = sparse ? fmb_sort_fields_sparse_min_rows(B) : ((int64_t)1 << 40);   // fields with >= split rows are sparse
    const bool do_dense = !(flags & FMB_SORT_PART_SPARSE), do_sparse = sparse && !(flags & FMB_SORT_PART_DENSE);
    if (!do_dense) {
    } else if (B <= 2048) {
        sort_fields_kernel<<<F, FS_THREADS, fmb::smem_sort_bytes(B), stream>>>(ids, B, F, field_off, sorted_keys, perm, posflag,
                                                                               rle, rlc, rcap, split);
        FMB_CHECK_LAUNCH("sort_fields_kernel");
    } else {
        if (Bq <= 256 * 8)
            sort_fields_cluster_kernel<256, true, 8><<<F * CL, 256, fmb::smem_sort_bytes(Bq) + (size_t)Bq * 6 + 16, stream>>>(
                ids, B, F, field_off, sorted_keys, perm, Bq, posflag, rle, rlc, rcap, split);
        else if (Bq <= 256 * FS_MAX_SLOTS)
            sort_fields_cluster_kernel<256, true, 16><<<F * CL, 256, fmb::smem_sort_bytes(Bq) + (size_t)Bq * 6 + 16, stream>>>(
                ids, B, F, field_off, sorted_keys, perm, Bq, posflag, rle, rlc, rcap, split);
        else
            sort_fields_cluster_kernel<1024, false, 16><<<F * CL, 1024, fmb::smem_sort_bytes(Bq), stream>>>(
                ids, B, F, field_off, sorted_keys, perm, Bq, posflag, rle, rlc, rcap, split);
        FMB_CHECK_LAUNCH("sort_fields_cluster_kernel");
    }
    if (do_sparse) {
        sparse_fields_kernel<<<F, SPARSE_THREADS, sparse_smem_bytes(B), stream>>>(ids, B, F, field_off, split, sorted_keys, perm,
                                                                                   posflag, rle, rlc, rcap);
        FMB_CHECK_LAUNCH("sparse_fields_kernel");
    }
    return FMB_OK;
}

FMB_API int fmb_sort_fields(const int32_t* ids, int B, int F, const int32_t* field_off, int32_t* sorted_keys,
                            int32_t* perm, cudaStream_t stream) {
    return fmb_sort_fields_ex(ids, B, F, field_off, sorted_keys, perm, nullptr, nullptr, 0, stream);
}
