// xs_sort.cuh -- hand-written LSD radix sort of the sampled lookups (key -> permutation).
//
// Replaces the Thrust calls of the reference's sorted variants: 12x thrust::count +
// thrust::sort_by_key by material (cuda/Simulation.cu:792-797), thrust::partition (:936) and
// the 12 per-material sort_by_key by energy (:1070-1075).  Here ONE packed 32-bit key
//     key = material << 28 | top 28 bits of the LCG state behind the energy
// is sorted, 8 bits per pass, carrying a 32-bit permutation index; the lookup kernel reads
// its samples through the permutation.  A material-only sort (-k 4) is a single 4-bit pass
// over bits 28..31, the fuel-first partition (-k 5) a single 1-bit pass.  The passes are
// stable, so equal keys keep lookup-id order.
//
// Each pass = histogram (per 2048-key tile) -> exclusive scan of the digit-major
// [digit][tile] table -> stable scatter using warp match_any ranking; 32-bit keys are reordered
// inside the tile (shared memory) before they are written, so the global writes are runs.
#pragma once

#include "xs_device.cuh"

namespace xs {

#ifndef XS_SCATTER_BLOCKS
#define XS_SCATTER_BLOCKS 6
#endif
constexpr int kSortThreads = 256;
#ifndef XS_SORT_ITEMS
#define XS_SORT_ITEMS 8
#endif
constexpr int kSortItems = XS_SORT_ITEMS;                   // keys per thread
constexpr int kSortTile = kSortThreads * kSortItems;        // 2048 keys per block: small tiles, 6 blocks/SM (the scatter is latency-bound)
constexpr int kRadix = 256;
constexpr int kSortWarps = kSortThreads / 32;

struct SortScratch {
    unsigned int *tile_hist = nullptr;      // [kRadix][n_tiles]
    unsigned int *chunk_sum = nullptr;      // scan partials
    long n_tiles_capacity = 0;
};

// (the struct must be declared before the helpers below)
inline int sort_scratch_alloc(SortScratch &s, long capacity_keys)
{
    const long tiles = (capacity_keys + kSortTile - 1) / kSortTile;
    if (tiles <= s.n_tiles_capacity) return 0;
    cudaFree(s.tile_hist);
    s.tile_hist = nullptr;
    if (cudaMalloc(&s.tile_hist, (size_t)tiles * kRadix * sizeof(unsigned int)) != cudaSuccess) return -1;
    cudaFree(s.chunk_sum);
    s.chunk_sum = nullptr;
    const long chunks = (tiles * kRadix + 4095) / 4096;
    if (cudaMalloc(&s.chunk_sum, (size_t)chunks * sizeof(unsigned int)) != cudaSuccess) return -1;
    s.n_tiles_capacity = tiles;
    return 0;
}
inline void sort_scratch_free(SortScratch &s)
{
    cudaFree(s.tile_hist); cudaFree(s.chunk_sum);
    s.tile_hist = nullptr; s.chunk_sum = nullptr; s.n_tiles_capacity = 0;
}

template <typename KeyT>
XS_DEV uint32_t sort_digit(KeyT key, int shift, uint32_t mask, int nonzero_flag)
{
    const uint32_t d = (uint32_t)(key >> shift) & mask;
    return nonzero_flag ? (uint32_t)(d != 0) : d;
}

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads)
sort_hist_kernel(const KeyT *__restrict__ keys, long n, int shift, uint32_t mask, int nonzero_flag,
                 unsigned int *__restrict__ tile_hist, int n_tiles)
{
    __shared__ unsigned int h[kRadix];
    h[threadIdx.x] = 0;
    __syncthreads();
    const long base = (long)blockIdx.x * kSortTile;
#pragma unroll 4
    for (int i = 0; i < kSortItems; i++) {
        const long idx = base + i * kSortThreads + threadIdx.x;
        if (idx < n) atomicAdd(&h[sort_digit(keys[idx], shift, mask, nonzero_flag)], 1u);
    }
    __syncthreads();
    tile_hist[(long)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

// In-place exclusive scan of `total` counters in two launches: per-chunk sums, then each block
// adds the sums of the chunks before it and scans its own chunk (4096 counters per block).
constexpr int kScanThreads = 1024;
constexpr int kScanChunk = kScanThreads * 4;

XS_DEV unsigned int block_exclusive_scan(unsigned int v, unsigned int *warp_sums, unsigned int *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned int t = __shfl_up_sync(kFullMask, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const unsigned int w = warp_sums[lane];
        unsigned int wi = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned int t = __shfl_up_sync(kFullMask, wi, off);
            if (lane >= off) wi += t;
        }
        warp_sums[lane] = wi - w;
        if (lane == 31 && total) *total = wi;
    }
    __syncthreads();
    const unsigned int r = warp_sums[warp] + incl - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads)
sort_chunk_sum_kernel(const unsigned int *data, long total, unsigned int *chunk_sum)
{
    __shared__ unsigned int warp_sums[32];
    __shared__ unsigned int block_total;
    const long base = (long)blockIdx.x * kScanChunk + threadIdx.x * 4;
    unsigned int v = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) if (base + i < total) v += data[base + i];
    block_exclusive_scan(v, warp_sums, &block_total);
    if (threadIdx.x == 0) chunk_sum[blockIdx.x] = block_total;
}

__global__ void __launch_bounds__(kScanThreads)
sort_scan_kernel(unsigned int *data, long total, const unsigned int *chunk_sum)
{
    __shared__ unsigned int warp_sums[32];
    __shared__ unsigned int prefix;
    // sum of all chunks before this one
    unsigned int before = 0;
    for (int c = threadIdx.x; c < (int)blockIdx.x; c += kScanThreads) before += chunk_sum[c];
    block_exclusive_scan(before, warp_sums, &prefix);      // prefix (shared) <- block total
    const long base = (long)blockIdx.x * kScanChunk + threadIdx.x * 4;
    unsigned int v[4], sum = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) { v[i] = base + i < total ? data[base + i] : 0u; sum += v[i]; }
    unsigned int run = prefix + block_exclusive_scan(sum, warp_sums, nullptr);
#pragma unroll
    for (int i = 0; i < 4; i++) if (base + i < total) { data[base + i] = run; run += v[i]; }
}

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads, sizeof(KeyT) == 4 ? XS_SCATTER_BLOCKS : 2)
sort_scatter_kernel(const KeyT *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                    KeyT *__restrict__ keys_out, uint32_t *__restrict__ vals_out, long n, int shift,
                    uint32_t mask, int nonzero_flag, const unsigned int *__restrict__ tile_base, int n_tiles)
{
    __shared__ unsigned int warp_cnt[kSortWarps][kRadix];
    for (int i = threadIdx.x; i < kSortWarps * kRadix; i += kSortThreads) (&warp_cnt[0][0])[i] = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const long base = (long)blockIdx.x * kSortTile + (long)warp * (32 * kSortItems);
    KeyT key[kSortItems];
    unsigned peers[kSortItems];
    unsigned short rank[kSortItems];

    // all key loads of the tile first (16 independent requests in flight per thread), then all
    // warp votes (independent of each other), and only then the short serial chain through the
    // per-warp digit counters: no round waits for memory or for a vote
#pragma unroll
    for (int r = 0; r < kSortItems; r++) {
        const long idx = base + r * 32 + lane;
        key[r] = idx < n ? keys_in[idx] : (KeyT)0;
    }
#pragma unroll
    for (int r = 0; r < kSortItems; r++) {
        const long idx = base + r * 32 + lane;
        const uint32_t d = idx < n ? sort_digit(key[r], shift, mask, nonzero_flag) : (uint32_t)kRadix;
        peers[r] = __match_any_sync(kFullMask, d);
    }
#pragma unroll
    for (int r = 0; r < kSortItems; r++) {
        const long idx = base + r * 32 + lane;
        const bool valid = idx < n;
        const uint32_t d = valid ? sort_digit(key[r], shift, mask, nonzero_flag) : (uint32_t)kRadix;
        const int leader = __ffs(peers[r]) - 1;
        unsigned int before = 0;
        if (valid && lane == leader) {
            before = warp_cnt[warp][d];
            warp_cnt[warp][d] = before + __popc(peers[r]);
        }
        before = __shfl_sync(kFullMask, before, leader);
        rank[r] = (unsigned short)(before + __popc(peers[r] & lt_mask));
        __syncwarp();
    }
    __syncthreads();
    if constexpr (sizeof(KeyT) == 4) {
        // 32-bit keys (the lookup sort): reorder the tile in shared memory first, then write it out
        // in tile order -- consecutive threads then write consecutive addresses of one digit's run
        // instead of 32 scattered words per store instruction.
        __shared__ KeyT s_key[kSortTile];
        __shared__ uint32_t s_val[kSortTile];
        __shared__ unsigned int s_gbase[kRadix];            // global position of a digit's run minus its start in the tile
        __shared__ unsigned int s_warp_tot[kSortWarps];
        const int d = threadIdx.x;                          // kSortThreads == kRadix
        unsigned int tot = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) tot += warp_cnt[w][d];
        unsigned int incl = tot;                            // exclusive scan of the 256 digit totals
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned int t = __shfl_up_sync(kFullMask, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) s_warp_tot[warp] = incl;
        __syncthreads();
        unsigned int start = incl - tot;
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) start += w < warp ? s_warp_tot[w] : 0u;
        s_gbase[d] = tile_base[(long)d * n_tiles + blockIdx.x] - start;
        unsigned int run = start;
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) {
            const unsigned int c = warp_cnt[w][d];
            warp_cnt[w][d] = run;
            run += c;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kSortItems; r++) {
            const long idx = base + r * 32 + lane;
            if (idx < n) {
                const uint32_t dg = sort_digit(key[r], shift, mask, nonzero_flag);
                const unsigned int at = warp_cnt[warp][dg] + rank[r];
                s_key[at] = key[r];
                s_val[at] = vals_in ? vals_in[idx] : (uint32_t)idx;
            }
        }
        __syncthreads();
        const long tile_first = (long)blockIdx.x * kSortTile;
        const int in_tile = (int)((n - tile_first < kSortTile) ? n - tile_first : kSortTile);
#pragma unroll 4
        for (int i = threadIdx.x; i < in_tile; i += kSortThreads) {
            const KeyT k = s_key[i];
            const unsigned int pos = s_gbase[sort_digit(k, shift, mask, nonzero_flag)] + (unsigned int)i;
            keys_out[pos] = k;
            vals_out[pos] = s_val[i];
        }
    } else {
        {   // per digit: warp offsets within the tile + the tile's global base
            const int d = threadIdx.x;
            unsigned int run = tile_base[(long)d * n_tiles + blockIdx.x];
#pragma unroll
            for (int w = 0; w < kSortWarps; w++) {
                const unsigned int c = warp_cnt[w][d];
                warp_cnt[w][d] = run;
                run += c;
            }
        }
        __syncthreads();
        uint32_t val[kSortItems];
#pragma unroll
        for (int r = 0; r < kSortItems; r++) {
            const long idx = base + r * 32 + lane;
            val[r] = (vals_in && idx < n) ? vals_in[idx] : (uint32_t)idx;
        }
#pragma unroll
        for (int r = 0; r < kSortItems; r++) {
            const long idx = base + r * 32 + lane;
            if (idx < n) {
                const uint32_t d = sort_digit(key[r], shift, mask, nonzero_flag);
                const unsigned int pos = warp_cnt[warp][d] + rank[r];
                keys_out[pos] = key[r];
                vals_out[pos] = val[r];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// One-launch-per-digit variant for the 32-bit lookup sort ("onesweep": chained scan with decoupled
// look-back).  One upfront kernel counts the digits of ALL passes (the global digit histogram of a pass does
// not depend on the order the earlier passes leave); then one kernel per pass: a block takes the next tile
// by an atomic ticket (tile index = ticket, so every tile it may have to wait for has been started), ranks
// its keys (the same warp match_any ranking as sort_scatter_kernel), publishes its per-digit counts as an
// AGGREGATE word, adds up its predecessors' words until it meets an INCLUSIVE one (thread d looks back for
// digit d, kLookBack predecessors per round trip), publishes its own INCLUSIVE word and scatters.  5 launches
// per 24-bit sort instead of 12, and the keys are read once per pass instead of twice.
//
// Round 1 had tried this with the 2,048-key tiles of the three-kernel scheme (888 tiles resident: the first
// wave walks hundreds of predecessors, 0.66-1.21 ms against 0.62); the tiles here are 4x-8x larger, so a few
// hundred are resident and a look-back is a handful of words.
// A status word = state (2 bits: 0 = nothing yet, 1 = aggregate, 2 = inclusive prefix) | count (30 bits):
// one 32-bit store publishes both, one 32-bit load reads both.  n < 2^30 keys (larger sorts take the
// three-kernel path).
// ---------------------------------------------------------------------------------------
#ifndef XS_ONESWEEP_ITEMS
#define XS_ONESWEEP_ITEMS 16
#endif
#ifndef XS_ONESWEEP_BLOCKS
#define XS_ONESWEEP_BLOCKS 3
#endif
constexpr int kSweepItems = XS_ONESWEEP_ITEMS;               // keys per thread
constexpr int kSweepTile = kSortThreads * kSweepItems;      // 4096 keys per block
#ifndef XS_ONESWEEP_LOOKBACK
#define XS_ONESWEEP_LOOKBACK 8
#endif
constexpr int kLookBack = XS_ONESWEEP_LOOKBACK;                                 // predecessors' words in flight per look-back round
constexpr int kMaxSortPasses = 4;

struct OnesweepScratch {
    unsigned int *ticket = nullptr;       // [16] next tile of each pass              } one allocation, in this order:
    unsigned int *digit_hist = nullptr;   // [kMaxSortPasses][kRadix] global digit counts of every pass   } a sort zeroes the
    unsigned int *status = nullptr;       // [n_passes][n_tiles][kRadix] look-back words                  } prefix it uses
    long n_tiles_capacity = 0;
};

inline int onesweep_scratch_alloc(OnesweepScratch &s, long capacity_keys)
{
    const long tiles = (capacity_keys + kSweepTile - 1) / kSweepTile;
    if (tiles <= s.n_tiles_capacity) return 0;
    cudaFree(s.ticket);
    s.ticket = nullptr;
    const size_t words = 16 + (size_t)kMaxSortPasses * kRadix + (size_t)kMaxSortPasses * tiles * kRadix;
    if (cudaMalloc(&s.ticket, words * sizeof(unsigned int)) != cudaSuccess) return -1;
    s.digit_hist = s.ticket + 16;
    s.status = s.digit_hist + kMaxSortPasses * kRadix;
    s.n_tiles_capacity = tiles;
    return 0;
}
inline void onesweep_scratch_free(OnesweepScratch &s)
{
    cudaFree(s.ticket);
    s = OnesweepScratch{};
}

__global__ void __launch_bounds__(kSortThreads)
onesweep_hist_kernel(const uint32_t *__restrict__ keys, long n, int lo_bit, int n_passes, int last_bits, unsigned int *__restrict__ digit_hist)
{
    __shared__ unsigned int h[kMaxSortPasses][kRadix];
    for (int i = threadIdx.x; i < kMaxSortPasses * kRadix; i += kSortThreads) (&h[0][0])[i] = 0;
    __syncthreads();
    const long stride = (long)gridDim.x * kSortThreads;
    for (long i = (long)blockIdx.x * kSortThreads + threadIdx.x; i < n; i += stride) {
        const uint32_t k = keys[i];
        for (int p = 0; p < n_passes; p++) {
            const uint32_t mask = (p == n_passes - 1) ? (1u << last_bits) - 1u : 0xffu;
            atomicAdd(&h[p][(k >> (lo_bit + 8 * p)) & mask], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_passes * kRadix; i += kSortThreads)
        if ((&h[0][0])[i]) atomicAdd(digit_hist + i, (&h[0][0])[i]);
}

// Lanes of the warp holding the same 8-bit digit (valid lanes only).  Eight ballots instead of one
// match.any: the MATCH instruction is slow on this GPU (every one of the 16 per thread showed up as a
// short-scoreboard stall, 35 % of the kernel's samples); votes issue at full rate.
XS_DEV unsigned warp_same_digit(uint32_t d, bool valid)
{
    unsigned peers = __ballot_sync(kFullMask, valid);
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const bool bit = (d >> b) & 1u;
        const unsigned m = __ballot_sync(kFullMask, bit);
        peers &= bit ? m : ~m;
    }
    return peers;
}

__global__ void __launch_bounds__(kSortThreads, XS_ONESWEEP_BLOCKS)
onesweep_pass_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                     uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, long n, int shift, uint32_t mask,
                     const unsigned int *__restrict__ digit_hist, unsigned int *status, unsigned int *ticket, int n_tiles)
{
    __shared__ unsigned int warp_cnt[kSortWarps][kRadix];
    extern __shared__ uint32_t s_tile_buf[];                // [2][kSweepTile]: the tile's keys and payloads in digit order (dynamic: > 48 KB for tiles beyond 4,096 keys)
    uint32_t *s_key = s_tile_buf, *s_val = s_tile_buf + kSweepTile;
    __shared__ unsigned int s_gbase[kRadix];                // global position of a digit's run minus its start in the tile
    __shared__ unsigned int s_warp_tot[kSortWarps];
    __shared__ unsigned int s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;

    // thread d: start of digit d in the whole array = keys with a smaller digit (exclusive scan of the pass's global
    // histogram, complete before the pass starts): the same for every tile, so once per block
    unsigned int digit_start;
    {
        const unsigned int mine = digit_hist[threadIdx.x];   // kSortThreads == kRadix
        unsigned int incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned int t = __shfl_up_sync(kFullMask, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) s_warp_tot[warp] = incl;
        __syncthreads();
        digit_start = incl - mine;
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) digit_start += w < warp ? s_warp_tot[w] : 0u;
    }

    for (;;) {
        __syncthreads();                                    // (the previous tile's shared state is no longer read)
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        for (int i = threadIdx.x; i < kSortWarps * kRadix; i += kSortThreads) (&warp_cnt[0][0])[i] = 0;
        __syncthreads();
        const int tile = (int)s_tile;
        if (tile >= n_tiles) break;

        const long base = (long)tile * kSweepTile + (long)warp * (32 * kSweepItems);
        uint32_t key[kSweepItems];
        unsigned short rank[kSweepItems];
#pragma unroll
        for (int r = 0; r < kSweepItems; r++) {
            const long idx = base + r * 32 + lane;
            key[r] = idx < n ? keys_in[idx] : 0u;
        }
        // ranking in rounds of 8 items: votes first (independent), then the short serial chain through the
        // per-warp digit counters
#pragma unroll
        for (int r0 = 0; r0 < kSweepItems; r0 += 8) {
            unsigned peers[8];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const long idx = base + (r0 + r) * 32 + lane;
                peers[r] = warp_same_digit((key[r0 + r] >> shift) & mask, idx < n);
            }
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const long idx = base + (r0 + r) * 32 + lane;
                const bool valid = idx < n;
                const uint32_t d = valid ? (key[r0 + r] >> shift) & mask : (uint32_t)kRadix;
                const int leader = __ffs(peers[r]) - 1;
                unsigned int before = 0;
                if (valid && lane == leader) {
                    before = warp_cnt[warp][d];
                    warp_cnt[warp][d] = before + __popc(peers[r]);
                }
                before = __shfl_sync(kFullMask, before, leader);
                rank[r0 + r] = (unsigned short)(before + __popc(peers[r] & lt_mask));
                __syncwarp();
            }
        }
        // the payloads: all loads in flight now, consumed after the look-back (one at a time inside the scatter
        // loop they were 29 % of the kernel's stall samples)
        uint32_t val[kSweepItems];
#pragma unroll
        for (int r = 0; r < kSweepItems; r++) {
            const long idx = base + r * 32 + lane;
            val[r] = (vals_in && idx < n) ? vals_in[idx] : (uint32_t)idx;
        }
        __syncthreads();

        // thread d: this tile's count of digit d -> publish, look back, publish again
        const int d = threadIdx.x;                          // kSortThreads == kRadix
        unsigned int tot = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) tot += warp_cnt[w][d];
        unsigned int *my_status = status + (size_t)tile * kRadix + d;
        if (tile > 0) {
            __stcg(my_status, (1u << 30) | tot);            // AGGREGATE
        }
        const unsigned int below = digit_start;
        unsigned int exclusive = 0;                          // keys of digit d in the tiles before this one
        int t = tile - 1;
        while (t >= 0) {
            unsigned int word[kLookBack];
#pragma unroll
            for (int i = 0; i < kLookBack; i++)
                word[i] = t - i >= 0 ? __ldcg(status + (size_t)(t - i) * kRadix + d) : (2u << 30);
            bool done = false;
#pragma unroll
            for (int i = 0; i < kLookBack; i++) {
                if (done) break;
                const unsigned int state = word[i] >> 30;
                if (state == 0) break;                      // not published yet: read again from here
                exclusive += word[i] & 0x3fffffffu;
                t--;
                if (state == 2) done = true;
            }
            if (done) break;
        }
        __stcg(my_status, (2u << 30) | (exclusive + tot));   // INCLUSIVE
        // tile-local layout: digit-major, warps in order
        {
            unsigned int incl = tot;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned int v = __shfl_up_sync(kFullMask, incl, off);
                if (lane >= off) incl += v;
            }
            if (lane == 31) s_warp_tot[warp] = incl;
            __syncthreads();
            unsigned int start = incl - tot;
#pragma unroll
            for (int w = 0; w < kSortWarps; w++) start += w < warp ? s_warp_tot[w] : 0u;
            s_gbase[d] = below + exclusive - start;
            unsigned int run = start;
#pragma unroll
            for (int w = 0; w < kSortWarps; w++) {
                const unsigned int c = warp_cnt[w][d];
                warp_cnt[w][d] = run;
                run += c;
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kSweepItems; r++) {
            const long idx = base + r * 32 + lane;
            if (idx < n) {
                const uint32_t dg = (key[r] >> shift) & mask;
                const unsigned int at = warp_cnt[warp][dg] + rank[r];
                s_key[at] = key[r];
                s_val[at] = val[r];
            }
        }
        __syncthreads();
        const long tile_first = (long)tile * kSweepTile;
        const int in_tile = (int)((n - tile_first < kSweepTile) ? n - tile_first : kSweepTile);
#pragma unroll 4
        for (int i = threadIdx.x; i < in_tile; i += kSortThreads) {
            const uint32_t k = s_key[i];
            const unsigned int pos = s_gbase[(k >> shift) & mask] + (unsigned int)i;
            if (keys_out) keys_out[pos] = k;                // (null in the last pass: nothing reads the sorted keys)
            vals_out[pos] = s_val[i];
        }
    }
}

// What a kernel that WRITES the keys needs to count their digits on the way (the sampler / locate kernels do:
// the sort then starts without onesweep_hist_kernel's pass over the keys).  digit_hist == nullptr: do not count.
struct DigitSpec {
    unsigned int *digit_hist = nullptr;   // [n_passes][kRadix], zeroed by onesweep_begin
    int lo_bit = 0, n_passes = 0, last_bits = 8;
};

// Shared-memory side of that: zero, count one key, flush (all threads of the block call zero / flush).
XS_DEV void digit_count_zero(unsigned int (*h)[kRadix])
{
    for (int i = threadIdx.x; i < kMaxSortPasses * kRadix; i += blockDim.x) (&h[0][0])[i] = 0;
}
XS_DEV void digit_count_key(unsigned int (*h)[kRadix], const DigitSpec &spec, uint32_t k)
{
    for (int p = 0; p < spec.n_passes; p++) {
        const uint32_t mask = (p == spec.n_passes - 1) ? (1u << spec.last_bits) - 1u : 0xffu;
        atomicAdd(&h[p][(k >> (spec.lo_bit + 8 * p)) & mask], 1u);
    }
}
// The keys the lookup pipeline sorts are (material << 28 | energy bits): when the last pass looks at bits 24..31,
// its counts, summed over the 16 energy prefixes of a material, are the material histogram.
XS_DEV bool digit_top_is_material(const DigitSpec &spec)
{
    return spec.digit_hist != nullptr && spec.last_bits == 8 && spec.lo_bit + 8 * (spec.n_passes - 1) == 24;
}
XS_DEV unsigned int digit_material_count(unsigned int (*h)[kRadix], const DigitSpec &spec, int m)
{
    unsigned int n = 0;
    for (int i = 0; i < 16; i++) n += h[spec.n_passes - 1][16 * m + i];
    return n;
}
XS_DEV void digit_count_flush(unsigned int (*h)[kRadix], const DigitSpec &spec)
{
    for (int i = threadIdx.x; i < spec.n_passes * kRadix; i += blockDim.x)
        if ((&h[0][0])[i]) atomicAdd(spec.digit_hist + i, (&h[0][0])[i]);
}

// Zero the scratch a sort of n keys over bits [lo_bit, hi_bit) uses and describe its digits.  Returns false when
// onesweep_sort would not take this sort (the caller's kernel then does not count: spec->digit_hist stays null).
inline bool onesweep_begin(OnesweepScratch &s, long n, int lo_bit, int hi_bit, cudaStream_t stream, DigitSpec *spec)
{
    *spec = DigitSpec{};
    const int n_tiles = (int)((n + kSweepTile - 1) / kSweepTile);
    const int n_passes = (hi_bit - lo_bit + 7) / 8;
    if (n_tiles == 0 || n_tiles > s.n_tiles_capacity || n_passes > kMaxSortPasses || n >= (1L << 30)) return false;
    const size_t used_words = 16 + (size_t)kMaxSortPasses * kRadix + (size_t)n_passes * n_tiles * kRadix;
    if (cudaMemsetAsync(s.ticket, 0, used_words * sizeof(unsigned int), stream) != cudaSuccess) return false;
    spec->digit_hist = s.digit_hist;
    spec->lo_bit = lo_bit;
    spec->n_passes = n_passes;
    spec->last_bits = hi_bit - lo_bit - 8 * (n_passes - 1);
    return true;
}

// 32-bit keys, identity payload, bits [lo_bit, hi_bit) in passes of 8 (the last one may be narrower).
// digits_counted: onesweep_begin ran for exactly this sort and the kernel that wrote the keys counted their digits.
inline int onesweep_sort(OnesweepScratch &s, uint32_t *key[2], uint32_t *perm[2], long n, int lo_bit, int hi_bit,
                         int sm_count, cudaStream_t stream, uint32_t **sorted_perm, int *launches, bool digits_counted = false)
{
    const int n_tiles = (int)((n + kSweepTile - 1) / kSweepTile);
    const int n_passes = (hi_bit - lo_bit + 7) / 8;
    if (n_tiles == 0) { *sorted_perm = perm[0]; return 0; }
    if (n_tiles > s.n_tiles_capacity || n_passes > kMaxSortPasses || n >= (1L << 30)) return -2;
    const int last_bits = hi_bit - lo_bit - 8 * (n_passes - 1);
    if (!digits_counted) {
        const size_t used_words = 16 + (size_t)kMaxSortPasses * kRadix + (size_t)n_passes * n_tiles * kRadix;
        if (cudaMemsetAsync(s.ticket, 0, used_words * sizeof(unsigned int), stream) != cudaSuccess) return -1;
        const int hist_blocks = (int)std::min<long>((n + kSortThreads * 16 - 1) / (kSortThreads * 16), (long)sm_count * 8);
        onesweep_hist_kernel<<<hist_blocks, kSortThreads, 0, stream>>>(key[0], n, lo_bit, n_passes, last_bits, s.digit_hist);
        if (launches) *launches += 1;
    }
    int cur = 0;
    const int blocks = std::min(n_tiles, sm_count * XS_ONESWEEP_BLOCKS);
    constexpr size_t kTileBytes = 2 * sizeof(uint32_t) * kSweepTile;
    if (kTileBytes > 48 * 1024 &&                            // (per device: set on whichever device is current)
        cudaFuncSetAttribute((const void *)onesweep_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileBytes) != cudaSuccess)
        return -1;
    for (int p = 0; p < n_passes; p++) {
        const uint32_t mask = (p == n_passes - 1) ? (1u << last_bits) - 1u : 0xffu;
        onesweep_pass_kernel<<<blocks, kSortThreads, kTileBytes, stream>>>(key[cur], p == 0 ? nullptr : perm[cur], p == n_passes - 1 ? nullptr : key[cur ^ 1], perm[cur ^ 1], n,
                                                                  lo_bit + 8 * p, mask, s.digit_hist + p * kRadix,
                                                                  s.status + (size_t)p * n_tiles * kRadix, s.ticket + p, n_tiles);
        cur ^= 1;
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    if (launches) *launches += n_passes;
    *sorted_perm = perm[cur];
    return 0;
}

// Sort bits [lo_bit, hi_bit) of key[0][0..n), 8 bits per pass, stable.  key[]/perm[] are
// ping-pong buffers; the payload starts as the identity unless `payload_ready` says perm[0]
// already holds one.  On return *sorted_perm / *sorted_key point at the buffers holding the
// final payload and keys.  Returns 0 on success.
template <typename KeyT>
inline int radix_sort(SortScratch &s, KeyT *key[2], uint32_t *perm[2], long n, int lo_bit, int hi_bit,
                      int nonzero_flag, bool payload_ready, cudaStream_t stream, uint32_t **sorted_perm,
                      KeyT **sorted_key, int *launches)
{
    const int n_tiles = (int)((n + kSortTile - 1) / kSortTile);
    if (n_tiles == 0) { *sorted_perm = perm[0]; if (sorted_key) *sorted_key = key[0]; return 0; }
    if (n_tiles > s.n_tiles_capacity) return -1;
    int cur = 0;
    bool first = !payload_ready;
    for (int shift = lo_bit; shift < hi_bit; shift += 8) {
        const int bits = std::min(8, hi_bit - shift);
        const uint32_t mask = (1u << bits) - 1u;
        sort_hist_kernel<KeyT><<<n_tiles, kSortThreads, 0, stream>>>(key[cur], n, shift, mask, nonzero_flag, s.tile_hist, n_tiles);
        const long n_counters = (long)n_tiles * kRadix;
        const int n_chunks = (int)((n_counters + kScanChunk - 1) / kScanChunk);
        sort_chunk_sum_kernel<<<n_chunks, kScanThreads, 0, stream>>>(s.tile_hist, n_counters, s.chunk_sum);
        sort_scan_kernel<<<n_chunks, kScanThreads, 0, stream>>>(s.tile_hist, n_counters, s.chunk_sum);
        sort_scatter_kernel<KeyT><<<n_tiles, kSortThreads, 0, stream>>>(key[cur], first ? nullptr : perm[cur], key[cur ^ 1],
                                                                        perm[cur ^ 1], n, shift, mask, nonzero_flag,
                                                                        s.tile_hist, n_tiles);
        if (cudaGetLastError() != cudaSuccess) return -1;
        if (launches) *launches += 4;
        cur ^= 1;
        first = false;
    }
    *sorted_perm = perm[cur];
    if (sorted_key) *sorted_key = key[cur];
    return 0;
}

// The lookup sort of -k 6 / xs_gpu_sort_keys: 32-bit keys, identity payload.
inline int sort_lookups(SortScratch &s, uint32_t *key[2], uint32_t *perm[2], long n, int lo_bit, int hi_bit,
                        int nonzero_flag, cudaStream_t stream, uint32_t **sorted_perm, int *launches)
{
    return radix_sort<uint32_t>(s, key, perm, n, lo_bit, hi_bit, nonzero_flag, false, stream, sorted_perm, nullptr, launches);
}

}  // namespace xs
