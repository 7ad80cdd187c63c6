// xs_tile.cuh -- xs_tile_kernel: the in-order variants (-k 0..3, xs_gpu_dump) in ONE launch with
// deliberate L2 locality.
//
// Replaces xs_lookup_kernel_baseline / _optimization_1..3 (cuda/Simulation.cu:44-99, 464-509, 576-626,
// 692-744): thread-per-lookup kernels that walk each lookup's nuclide list front to back.  The
// reference gets 70 % L2 hits out of that by accident -- its blocks all take equally long, so waves of
// blocks walk the fuel list in step and the working set at any moment is the few nuclides around the
// current position (profiles/r01_notes.md).  A persistent, dynamically scheduled kernel destroys that
// (round 1's xs_event_kernel: 113 GB of DRAM traffic per run, 1.17 x the algorithmic gather bytes).
// Here the locality is made on purpose, inside one launch and without a global regrouping pass:
//
//   * the grid is persistent and works in ROUNDS: in round r block b takes tile r * gridDim.x + b
//     (kTile lookups);
//   * a tile is sampled (or read from the sample arrays), its unionized rows / hash bins are located
//     and its lookups are grouped by material IN SHARED MEMORY (counting sort of tile-local indices);
//   * fuel (every material with more nuclides than one window) is swept window by window -- all of the
//     tile's fuel lookups pass nuclides [32 w, 32 w + 32) before any of them advances; the five partial
//     sums of a lookup wait in shared memory between windows.  Every block does this at the same time
//     on the same window, so a window's pair records (32 x 1.45 MB) stay in the 126 MB L2 while the
//     whole grid reads them;
//   * the small materials (<= one window of nuclides) follow, one sweep each;
//   * a grid barrier (one atomic counter, all blocks co-resident: cooperative launch) ends the round:
//     the blocks cannot drift apart by more than one tile.
//
// The sweep itself is xs_window_kernel's (window_sweep_group: 8 lookups x 4 lanes per warp, one
// 256-bit load per lane and step, reference-order accumulation, Newton-Markstein f): macro_xs is
// bit-identical to the reference on this path too.
#pragma once

#include "xs_kernels.cuh"

namespace xs {

// Tile size and residency, measured (large, unionized, -k 0, ms per 17 M): 4096 lookups x 2 blocks/SM (128 registers)
// 19.0; 2048 x 3 (80 registers) 13.6; 1024 x 4 (64 registers) 13.8; 1792 x 4 21.3; 1024 x 5 (48 registers, spills)
// 25.1 -- the sweep is latency-bound (L2 gather), so warps per SM count until the register cap bites.
#ifndef XS_TILE_LOOKUPS
#define XS_TILE_LOOKUPS 2048
#endif
#ifndef XS_TILE_WIDE
#define XS_TILE_WIDE 1               // nuclide-grid searches probe their last <= 4 candidates at once (-k 0, 17 M: 371 -> 470 M lookups/s;
#endif                               // the hash grid's bracket search is better off one probe at a time: 399 vs 336)
#ifndef XS_TILE_WINDOW_SYNC
#define XS_TILE_WINDOW_SYNC 1
#endif
#ifndef XS_TILE_BLOCKS
#define XS_TILE_BLOCKS 3
#endif
constexpr int kTile = XS_TILE_LOOKUPS;                  // lookups per tile (a multiple of the block size)
constexpr int kTilePartials = kTile * 3 / 16;                    // lookups of a multi-window material whose partial sums fit in shared memory at once
static_assert(kTile % kBlockThreads == 0 && kTile <= 65536, "tile geometry");

struct TileShared {
    unsigned long long part[kWarpsPerBlock];
    int    mat_count[kNumMaterials + 4];                 // lookups per material in this tile (after the filter)
    int    mat_begin[kNumMaterials + 4];                 // exclusive prefix: start of each material in order[]
    int    cursor[kNumMaterials + 4];
    uint32_t rec[kWarpsPerBlock][kSweepSlots][kMaxWindow + 1];
    double   energy[kTile];
    uint32_t where[kTile];
    unsigned short order[kTile];                         // tile-local lookup indices grouped by material
    signed char mat[kTile];                              // material, or -1 = not performed (filtered out / past the end)
    double2 partial[3 * kTilePartials];                  // (total, elastic) (absorbtion, fission) (nu_fission, -) per lookup
};

// All blocks of the (co-resident) grid arrive; thread 0 of each waits until the round's count is
// complete.  The counter only grows: no reset, no sense reversal.
XS_DEV void grid_barrier(unsigned int *counter, unsigned int target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (*(volatile unsigned int *)counter < target) __nanosleep(64);
        __threadfence();
    }
    __syncthreads();
}

template <int GRID>
__global__ void __launch_bounds__(kBlockThreads, XS_TILE_BLOCKS)
xs_tile_kernel(const __grid_constant__ Problem P, const BatchSource src, const BatchSink sink, const __grid_constant__ ConcTable C,
               int window, int use_barrier)
{
    extern __shared__ __align__(16) unsigned char tile_smem[];
    TileShared &T = *reinterpret_cast<TileShared *>(tile_smem);
    int *s_nuc = reinterpret_cast<int *>(tile_smem + sizeof(TileShared));
    int *s_first = s_nuc + P.mat_total;
    for (int i = threadIdx.x; i < P.mat_total; i += blockDim.x) s_nuc[i] = P.mat_nuc[i];
    for (int i = threadIdx.x; i <= kNumMaterials; i += blockDim.x) s_first[i] = P.mat_first[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = lane >> 2, quarter = lane & 3;
    unsigned long long my_sum = 0, my_count = 0;
    const long n_tiles = (src.count + kTile - 1) / kTile;
    const Affine hop = lcg_jump(2ULL * kBlockThreads);      // from one of a thread's lookups to its next one in the tile

    for (long round = 0; round * gridDim.x < n_tiles; round++) {
        const long tile = round * gridDim.x + blockIdx.x;
        const long tile_first = tile * kTile;
        // ---- sample (or fetch) the tile, locate rows / bins, count materials ----------------------------
        if (threadIdx.x < kNumMaterials + 4) T.mat_count[threadIdx.x] = 0;
        __syncthreads();
        if (tile < n_tiles) {
            uint64_t s = 0;
            if (!src.energy) s = lcg_skip(kStartSeed, 2ULL * (uint64_t)(src.first_id + tile_first + threadIdx.x));
#pragma unroll 1
            for (int i = threadIdx.x; i < kTile; i += kBlockThreads) {
                const long t = tile_first + i;
                double e = 0.5;
                int m = -1;
                if (t < src.count) {
                    if (src.energy) {
                        const long at = src.perm ? (long)src.perm[t] : t;
                        e = src.energy[at];
                        m = src.mat[at];
                    } else {
                        const uint64_t s1 = lcg_step(s), s2 = lcg_step(s1);
                        e = lcg_to_double(s1);
                        m = pick_material(P, lcg_to_double(s2));
                    }
                    if (sink.energy_out) sink.energy_out[t] = e;
                    if (sink.mat_out) sink.mat_out[t] = m;
                    if (m < src.mat_lo || m > src.mat_hi) m = -1;
                }
                s = apply(hop, s);
                T.energy[i] = e;
                if (m >= 0) {
                    const uint32_t w = (uint32_t)locate<GRID>(P, e);
                    if (src.row_end && !(w >= src.row_begin && w < src.row_end)) m = -1;     // another device's energy band
                    else { T.where[i] = w; atomicAdd(&T.mat_count[m], 1); }
                }
                T.mat[i] = (signed char)m;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int run = 0;
            for (int m = 0; m < kNumMaterials; m++) { T.mat_begin[m] = run; T.cursor[m] = run; run += T.mat_count[m]; }
            T.mat_begin[kNumMaterials] = run;
        }
        __syncthreads();
        if (tile < n_tiles) {                               // (a block without a tile in the last round only keeps the barrier company)
            for (int i = threadIdx.x; i < kTile; i += kBlockThreads) {
                const int m = T.mat[i];
                if (m >= 0) T.order[atomicAdd(&T.cursor[m], 1)] = (unsigned short)i;
            }
        }
        __syncthreads();

        // ---- material by material; a material with more nuclides than one window goes window by window ----
        for (int m = 0; m < kNumMaterials; m++) {
            const int n_look = T.mat_count[m];
            const int first = s_first[m], n_nuc = s_first[m + 1] - first;
            if (n_nuc == 0 || n_look == 0) continue;        // (block-uniform: the counts live in shared memory)
            int passes = (n_nuc + window - 1) / window;
            // (a remainder of up to one staging row is folded into the last window, like launch_sweep does)
            if (passes > 1 && n_nuc - (passes - 2) * window <= kMaxWindow) passes--;
            // a multi-window material keeps kTilePartials lookups' partial sums in shared memory at a time
            const int batch = passes > 1 ? kTilePartials : n_look;
            for (int b0 = 0; b0 < n_look; b0 += batch) {
                const int nb = min(batch, n_look - b0);
                for (int p = 0; p < passes; p++) {
                    const int j_begin = p * window, j_end = (p == passes - 1) ? n_nuc : (p + 1) * window;
                    const bool first_window = p == 0, last_window = p == passes - 1;
                    for (int g0 = warp * kSweepSlots; g0 < nb; g0 += kWarpsPerBlock * kSweepSlots) {
                        const int in_batch = g0 + slot;
                        const bool on = in_batch < nb;
                        const int local = on ? T.order[T.mat_begin[m] + b0 + in_batch] : 0;
                        const double e = on ? T.energy[local] : 0.5;
                        const uint32_t where32 = on ? T.where[local] : 0u;
                        double acc_x = 0.0, acc_y = 0.0;
                        if (on && !first_window && quarter < 3) {
                            const double2 part = T.partial[3 * in_batch + quarter];
                            acc_x = part.x;
                            acc_y = part.y;
                        }
                        window_sweep_group<GRID, XS_TILE_WIDE != 0 && GRID == kNuclide>(P, C, T.rec[warp], s_nuc + first + j_begin, j_end - j_begin, C.first[m] + j_begin,
                                                 min(kSweepSlots, nb - g0), where32, e, lane, acc_x, acc_y);
                        if (!last_window) {
                            if (on && quarter < 3) T.partial[3 * in_batch + quarter] = make_double2(acc_x, acc_y);
                            continue;                        // warp-uniform
                        }
                        // lane 0 = (total, elastic)  lane 1 = (absorbtion, fission)  lane 2 = (nu_fission, -)
                        const double q1x = __shfl_down_sync(kFullMask, acc_x, 1), q1y = __shfl_down_sync(kFullMask, acc_y, 1);
                        const double q2x = __shfl_down_sync(kFullMask, acc_x, 2);
                        if (on && quarter == 0) {
                            const double v[5] = {acc_x, acc_y, q1x, q1y, q2x};
                            double gap;
                            const int am = argmax5(v, gap);
                            my_sum += (unsigned long long)(am + 1);
                            my_count += 1;
                            const long t = tile_first + local;
                            if (sink.macro_out) {
#pragma unroll
                                for (int k = 0; k < 5; k++) sink.macro_out[5 * t + k] = v[k];
                            }
                            if (sink.argmax_out) sink.argmax_out[t] = am;
                        }
                    }
                    // (multi-window materials: the block moves to the next window together.  Nothing is exchanged
                    // -- a warp keeps its own groups through all windows -- and with 36 fuel groups over 8 warps
                    // the barrier makes the warps with 4 groups wait for those with 5 (23 % of the stall samples);
                    // but without it the warps drift apart and the window falls out of L2: 15.2 instead of 13.6 ms)
                    if (XS_TILE_WINDOW_SYNC && passes > 1) __syncthreads();
                }
            }
        }
        if (use_barrier) grid_barrier(sink.batch_counter, (unsigned int)(round + 1) * gridDim.x);
        else __syncthreads();
    }

    const unsigned long long bs = block_sum(my_sum, T.part);
    const unsigned long long bc = block_sum(my_count, T.part);
    if (threadIdx.x == 0 && bc) {
        atomicAdd(sink.accum, bs);
        atomicAdd(sink.accum + 1, bc);
    }
}

}  // namespace xs
