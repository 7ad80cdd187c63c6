// xs_kernels.cuh -- the lookup kernels.
//
// Three families (dispatch in xs_gpu.cu):
//
//  * SORTED pipeline (-k 6, xs_gpu_lookup_samples) -- the fastest path, the bench headline:
//      xs_sample_kernel / xs_locate_kernel        energy, material, UEG row, sort key, histogram
//      radix sort (xs_sort.cuh)                   order by (material, energy)
//      xs_dense_kernel / xs_sorted_kernel         lane per lookup; the lookups of a warp share their
//                                                 pair records (shared-memory ring, LDS broadcasts).
//                                                 dense materials (many lookups per grid interval):
//                                                 only the group's lowest / highest energy is
//                                                 resolved, a lookup finds its record by comparing
//                                                 its energy with the records' bounds; the others:
//                                                 every lookup's record number is staged
//  * SWEEP pipeline (-k 4/5, history mode) -- lookups grouped by material only:
//      xs_sample_kernel / xs_history_step_kernel  energy, material, UEG row, histogram
//      xs_partition_kernel                        group the lookups by material
//      xs_window_kernel                           per material and nuclide window:
//                                                 pair records from L2, 4 lanes per lookup
//  * IN-ORDER kernel (-k 0..3, xs_gpu_dump): xs_event_kernel -- one kernel body; what changes
//    between variants is where a warp's batch of 32 lookups comes from (sampled in-kernel
//    from the lookup id, or read from sample arrays, optionally through a material filter).
//      - persistent grid, warps fetch BATCHES of 32 lookups from a global atomic counter;
//      - lane-parallel scalar work: lane l draws the sample of lookup l and finds its row/bin;
//      - small materials one lookup per lane in reference order (lane_macro_small), fuel by the
//        whole warp, 3 lanes per nuclide (warp_macro_big), shuffle reduction, near-tie guard;
//      - verification sum reduced warp -> block -> one atomicAdd(u64) per block, so the
//        reference's 136 MB verification[] array and its thrust::reduce pass do not exist.
#pragma once

#include "xs_device.cuh"
#include "xs_sort.cuh"
#include <type_traits>

namespace xs {

constexpr int kBlockThreads = 256;
constexpr int kWarpsPerBlock = kBlockThreads / 32;
constexpr double kTieGuard = 1e-10;      // relative gap below which order of summation could
                                         // change an integer decision -> settle serially

// Gather strategies of the in-order kernel (XSB200_GATHER).
//   kLanePerNuclide (0): every lookup by the whole warp, lane j handles nuclide j (+32, ...):
//                     6 divergent 16-B loads per lane and round.  Simple; kept for comparison.
//   kTriple (1, default, "hybrid"): small materials (n <= 32) are evaluated one lookup per lane
//                     (phase A, lane_macro_small); big ones (fuel) by the whole warp with 3
//                     lanes per nuclide and kBigUnroll x 10 nuclides in flight (phase B,
//                     warp_macro_big).
constexpr int kLanePerNuclide = 0, kTriple = 1;
constexpr int kBigUnroll = 4;            // phase B: 4 x 10 nuclides in flight per warp
constexpr int kSmallMax = 32;            // materials with at most this many nuclides go to phase A

// Where a batch's (energy, material) pairs come from.
struct BatchSource {
    const double *energy;      // sample arrays (nullptr => sample in-kernel from lookup ids)
    const int    *mat;
    const uint32_t *perm;      // optional permutation: slot t reads sample perm[t]
    long   first_id;           // in-kernel sampling: id of slot 0
    long   count;              // number of slots
    int    mat_lo, mat_hi;     // only lookups with mat in [mat_lo, mat_hi] are performed
    uint32_t row_begin, row_end; // energy-band sharding: only lookups whose unionized row lies in [row_begin, row_end)
                                 // are performed (row_end == 0: no filter)
};

struct BatchSink {
    unsigned long long *accum; // [0] += sum(argmax+1), [1] += lookups performed
    unsigned int *batch_counter;
    double *macro_out;         // optional [count*5]
    double *energy_out;        // optional [count]
    int    *mat_out;           // optional [count]
    int    *argmax_out;        // optional [count]
    unsigned char *fwd_out;    // optional [count]: history mode, #{k : macro_xs[k] > 1.0} per sample id
};

// Material tables staged in shared memory (compact CSR: 484 entries at "large").
struct SharedTables {
    int    first[kNumMaterials + 1];
    int    pad[3];
    double exchange[kWarpsPerBlock][8];   // per-warp scratch for the 3-lane reduction
    int    stage[kWarpsPerBlock][2 * 10 * 4];     // per-warp hand-over of resolved grid points (phase B)
};

XS_DEV void stage_tables(const Problem &P, SharedTables &T, int *s_nuc, double *s_conc)
{
    for (int i = threadIdx.x; i <= kNumMaterials; i += blockDim.x) T.first[i] = P.mat_first[i];
    for (int i = threadIdx.x; i < P.mat_total; i += blockDim.x) {
        s_nuc[i] = P.mat_nuc[i];
        s_conc[i] = P.mat_conc[i];
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// Warp-cooperative macroscopic lookup.  e / where / (first,n) are warp-uniform.  On return
// every lane holds the same five sums.
// ---------------------------------------------------------------------------------------
template <int GRID>
XS_DEV void warp_macro_lane_per_nuclide(const Problem &P, const int *s_nuc, const double *s_conc,
                                        int first, int n, double e, long where, int lane,
                                        double out[5])
{
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int j = lane; j < n; j += 32) {
        const int nuc = s_nuc[first + j];
        const double conc = s_conc[first + j];
        const int low = nuclide_low<GRID>(P, e, where, nuc);
        const double2 *p = P.grid + 3 * ((long)nuc * P.n_gp + low);
        const double2 l0 = ldg_grid(p), l1 = ldg_grid(p + 1), l2 = ldg_grid(p + 2);
        const double2 h0 = ldg_grid(p + 3), h1 = ldg_grid(p + 4), h2 = ldg_grid(p + 5);
        const double f = (h0.x - e) / (h0.x - l0.x);
        acc[0] += lerp_xs(l0.y, h0.y, f) * conc;
        acc[1] += lerp_xs(l1.x, h1.x, f) * conc;
        acc[2] += lerp_xs(l1.y, h1.y, f) * conc;
        acc[3] += lerp_xs(l2.x, h2.x, f) * conc;
        acc[4] += lerp_xs(l2.y, h2.y, f) * conc;
    }
#pragma unroll
    for (int k = 0; k < 5; k++) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1)
            acc[k] += __shfl_xor_sync(kFullMask, acc[k], off);
        out[k] = acc[k];
    }
}

// ---------------------------------------------------------------------------------------
// Phase A of a batch: every lane evaluates ITS OWN lookup when that lookup's material is small
// (n <= 32 nuclides: everything except fuel).  Two nuclides per iteration are in flight (12
// 16-byte loads per lane) and the index entries of the next pair are requested before the
// current pair is consumed.  Contributions are added in reference order j = 0,1,2,... so
// these lookups are bit-identical to the reference.  n == 0 switches a lane off.
// ---------------------------------------------------------------------------------------
template <int GRID>
XS_DEV void lane_macro_small(const Problem &P, const int *s_nuc, const double *s_conc, int first, int n,
                             double e, long where, double acc[5])
{
#pragma unroll
    for (int k = 0; k < 5; k++) acc[k] = 0.0;
    const int n_max = __reduce_max_sync(kFullMask, n);
    int nuc0 = 0, nuc1 = 0, low0 = 0, low1 = 0;
    if (0 < n) { nuc0 = s_nuc[first];     low0 = nuclide_low<GRID>(P, e, where, nuc0); }
    if (1 < n) { nuc1 = s_nuc[first + 1]; low1 = nuclide_low<GRID>(P, e, where, nuc1); }
    for (int j = 0; j < n_max; j += 2) {
        const bool on0 = j < n, on1 = j + 1 < n;
        double2 a0, a1, a2, a3, a4, a5, b0, b1, b2, b3, b4, b5;
        a0 = a1 = a2 = a4 = a5 = b0 = b1 = b2 = b4 = b5 = make_double2(0.0, 0.0);
        a3 = b3 = make_double2(1.0, 0.0);
        if (on0) {
            const double2 *p = P.grid + 3 * ((long)nuc0 * P.n_gp + low0);
            a0 = ldg_grid(p); a1 = ldg_grid(p + 1); a2 = ldg_grid(p + 2);
            a3 = ldg_grid(p + 3); a4 = ldg_grid(p + 4); a5 = ldg_grid(p + 5);
        }
        if (on1) {
            const double2 *p = P.grid + 3 * ((long)nuc1 * P.n_gp + low1);
            b0 = ldg_grid(p); b1 = ldg_grid(p + 1); b2 = ldg_grid(p + 2);
            b3 = ldg_grid(p + 3); b4 = ldg_grid(p + 4); b5 = ldg_grid(p + 5);
        }
        const double conc0 = on0 ? s_conc[first + j] : 0.0;
        const double conc1 = on1 ? s_conc[first + j + 1] : 0.0;
        // request the next two index entries before touching the data above
        if (j + 2 < n) { nuc0 = s_nuc[first + j + 2]; low0 = nuclide_low<GRID>(P, e, where, nuc0); }
        if (j + 3 < n) { nuc1 = s_nuc[first + j + 3]; low1 = nuclide_low<GRID>(P, e, where, nuc1); }
        if (on0) {
            const double f = (a3.x - e) / (a3.x - a0.x);
            acc[0] += lerp_xs(a0.y, a3.y, f) * conc0;
            acc[1] += lerp_xs(a1.x, a4.x, f) * conc0;
            acc[2] += lerp_xs(a1.y, a4.y, f) * conc0;
            acc[3] += lerp_xs(a2.x, a5.x, f) * conc0;
            acc[4] += lerp_xs(a2.y, a5.y, f) * conc0;
        }
        if (on1) {
            const double f = (b3.x - e) / (b3.x - b0.x);
            acc[0] += lerp_xs(b0.y, b3.y, f) * conc1;
            acc[1] += lerp_xs(b1.x, b4.x, f) * conc1;
            acc[2] += lerp_xs(b1.y, b4.y, f) * conc1;
            acc[3] += lerp_xs(b2.x, b5.x, f) * conc1;
            acc[4] += lerp_xs(b2.y, b5.y, f) * conc1;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Phase B: one big lookup (fuel: 321 nuclides) by the whole warp.  Three adjacent lanes share
// a nuclide -- lane c loads 16-byte chunk c of the low and of the high grid point, so one
// load instruction touches 48 contiguous bytes per nuclide (about 1.3 L1 wavefronts per
// nuclide and instruction instead of 6 for one lane per nuclide).  U steps of 10 nuclides are
// in flight per iteration (2U loads per lane).  The lower grid-point index of every nuclide is
// resolved LANE-PARALLEL (one nuclide per lane: 32 index loads / searches in flight, coalesced
// index-row segments) one iteration ahead and handed over through `stage` (2 x 10U ints of
// shared memory per warp).  e / where / (first, n) are warp-uniform; on return every lane
// holds the same five sums.
// ---------------------------------------------------------------------------------------
template <int GRID, int U>
XS_DEV void warp_macro_big(const Problem &P, const int *s_nuc, const double *s_conc, int first, int n,
                           double e, long where, int lane, double *exchange /* 8 doubles, per warp */,
                           int *stage /* 2 * 10U ints, per warp */, double out[5])
{
    constexpr int kPerIter = 10 * U;
    const int slot = lane / 3;                 // 10 nuclides per step; lanes 30,31 idle
    const int chunk = lane - 3 * slot;         // which 16-byte chunk of a grid point
    const bool lane_on = lane < 30;
    const int f_src = lane - chunk;            // lane holding both energies of this slot
    double acc_x = 0.0, acc_y = 0.0;

    auto resolve = [&](int j0, int *buf) {     // stage[j - j0] = low point of nuclide j, j in [j0, j0+10U)
#pragma unroll
        for (int r = 0; r < (kPerIter + 31) / 32; r++) {
            const int i = r * 32 + lane;
            if (i < kPerIter && j0 + i < n) {
                const int nuc = s_nuc[first + j0 + i];
                buf[i] = nuc * P.n_gp + nuclide_low<GRID>(P, e, where, nuc);      // < 2^31 up to XXL
            }
        }
    };
    __syncwarp();
    resolve(0, stage);
    __syncwarp();
    int it = 0;
    for (int j0 = 0; j0 < n; j0 += kPerIter, it ^= 1) {
        const int *cur = stage + it * kPerIter;
        double2 lo[U], hi[U];
        double conc[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            lo[u] = make_double2(0.0, 0.0);
            hi[u] = make_double2(1.0, 0.0);
            conc[u] = 0.0;
            const int i = u * 10 + slot;
            if (lane_on && j0 + i < n) {
                const double2 *p = P.grid + 3 * (long)cur[i] + chunk;
                lo[u] = ldg_grid(p);
                hi[u] = ldg_grid(p + 3);
                conc[u] = s_conc[first + j0 + i];
            }
        }
        if (j0 + kPerIter < n) resolve(j0 + kPerIter, stage + (it ^ 1) * kPerIter);   // next iteration's points
        double f[U];
#pragma unroll
        for (int u = 0; u < U; u++) f[u] = (hi[u].x - e) / (hi[u].x - lo[u].x);   // used from chunk 0
#pragma unroll
        for (int u = 0; u < U; u++) {
            const double fu = __shfl_sync(kFullMask, f[u], f_src);
            acc_x += lerp_xs(lo[u].x, hi[u].x, fu) * conc[u];                     // conc == 0 when off
            acc_y += lerp_xs(lo[u].y, hi[u].y, fu) * conc[u];
        }
        __syncwarp();
    }
    // Sum the 10 slots: lanes 0,1,2 end up with the totals of chunk 0,1,2.
#pragma unroll
    for (int off = 24; off >= 3; off >>= 1) {
        const double tx = __shfl_down_sync(kFullMask, acc_x, off);
        const double ty = __shfl_down_sync(kFullMask, acc_y, off);
        if (lane + off < 30) { acc_x += tx; acc_y += ty; }
    }
    // chunk0 = (energy, total) chunk1 = (elastic, absorbtion) chunk2 = (fission, nu_fission)
    __syncwarp();
    if (lane < 3) { exchange[2 * lane] = acc_x; exchange[2 * lane + 1] = acc_y; }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 5; k++) out[k] = exchange[k + 1];
}


template <int GRID, int GATHER>
XS_DEV void warp_macro(const Problem &P, const int *s_nuc, const double *s_conc, int first, int n,
                       double e, long where, int lane, double *exchange, int *stage, double out[5])
{
    if (GATHER == kTriple) warp_macro_big<GRID, kBigUnroll>(P, s_nuc, s_conc, first, n, e, where, lane, exchange, stage, out);
    else                   warp_macro_lane_per_nuclide<GRID>(P, s_nuc, s_conc, first, n, e, where, lane, out);
}

XS_DEV unsigned long long block_sum(unsigned long long v, unsigned long long *s_part)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(kFullMask, v, off);
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) s_part[warp] = v;
    __syncthreads();
    unsigned long long t = 0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kWarpsPerBlock; w++) t += s_part[w];
    __syncthreads();
    return t;      // valid on thread 0
}

// ---------------------------------------------------------------------------------------
// Event-mode kernel (all -k variants and the host-sample path).
// ---------------------------------------------------------------------------------------
template <int GRID, int GATHER>
__global__ void __launch_bounds__(kBlockThreads, 2)
xs_event_kernel(const Problem P, const BatchSource src, const BatchSink sink)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SharedTables &T = *reinterpret_cast<SharedTables *>(smem_raw);
    double *s_conc = reinterpret_cast<double *>(smem_raw + sizeof(SharedTables));
    int *s_nuc = reinterpret_cast<int *>(s_conc + P.mat_total);
    __shared__ unsigned long long s_part[kWarpsPerBlock];
    stage_tables(P, T, s_nuc, s_conc);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *exchange = T.exchange[warp];
    unsigned long long my_sum = 0, my_count = 0;
    const long n_batches = (src.count + 31) / 32;

    for (;;) {
        long batch = 0;
        if (lane == 0) batch = atomicAdd(sink.batch_counter, 1u);
        batch = __shfl_sync(kFullMask, batch, 0);
        if (batch >= n_batches) break;

        // ---- lane-parallel part: sample (or fetch) lookup `slot`, locate its row/bin ----
        const long slot = batch * 32 + lane;
        const bool have = slot < src.count;
        double e_l = 0.5;
        int mat_l = -1;
        if (have) {
            if (src.energy) {
                const long at = src.perm ? (long)src.perm[slot] : slot;
                e_l = src.energy[at];
                mat_l = src.mat[at];
            } else {
                uint64_t s = lcg_skip(kStartSeed, 2ULL * (uint64_t)(src.first_id + slot));
                s = lcg_step(s); e_l = lcg_to_double(s);
                s = lcg_step(s); mat_l = pick_material(P, lcg_to_double(s));
            }
        }
        const bool want = have && mat_l >= src.mat_lo && mat_l <= src.mat_hi;
        long where_l = want ? locate<GRID>(P, e_l) : 0;
        const int first_l = want ? T.first[mat_l] : 0;
        const int n_l = want ? T.first[mat_l + 1] - first_l : 0;
        int my_argmax = 0;
        double mine[5] = {0, 0, 0, 0, 0};

        // ---- phase A: small materials, one lookup per lane --------------------------------
        const bool small_l = GATHER == kTriple && want && n_l <= kSmallMax;
        if (GATHER == kTriple && __any_sync(kFullMask, small_l))
            lane_macro_small<GRID>(P, s_nuc, s_conc, first_l, small_l ? n_l : 0, e_l, where_l, mine);

        // ---- phase B: big materials, the whole warp on one lookup at a time ---------------
        unsigned todo = __ballot_sync(kFullMask, want && !small_l);
        while (todo) {
            const int i = __ffs(todo) - 1;
            todo &= todo - 1;
            const double e = __shfl_sync(kFullMask, e_l, i);
            const int mat = __shfl_sync(kFullMask, mat_l, i);
            const long where = __shfl_sync(kFullMask, where_l, i);
            const int first = T.first[mat], n = T.first[mat + 1] - first;
            double out[5];
            warp_macro<GRID, GATHER>(P, s_nuc, s_conc, first, n, e, where, lane, exchange, T.stage[warp], out);
            if (lane == i) {
#pragma unroll
                for (int k = 0; k < 5; k++) mine[k] = out[k];
            }
        }
        if (want) {
            double gap;
            my_argmax = argmax5(mine, gap);
            if (gap <= kTieGuard) {            // near-tie: redo in reference order, one thread
                macro_xs_serial<GRID>(P, e_l, mat_l, mine);
                my_argmax = argmax5(mine, gap);
            }
            my_sum += (unsigned long long)(my_argmax + 1);
            my_count += 1;
            if (sink.macro_out) {
#pragma unroll
                for (int k = 0; k < 5; k++) sink.macro_out[5 * slot + k] = mine[k];
            }
            if (sink.argmax_out) sink.argmax_out[slot] = my_argmax;
        }
        if (have) {
            if (sink.energy_out) sink.energy_out[slot] = e_l;
            if (sink.mat_out) sink.mat_out[slot] = mat_l;
        }
    }

    const unsigned long long bs = block_sum(my_sum, s_part);
    const unsigned long long bc = block_sum(my_count, s_part);
    if (threadIdx.x == 0 && bc) {
        atomicAdd(sink.accum, bs);
        atomicAdd(sink.accum + 1, bc);
    }
}

// ---------------------------------------------------------------------------------------
// Window kernel: lookups grouped by material, restricted to one nuclide window [j_begin, j_end)
// of the material's nuclide list ("windowed nuclide sweep").
//
// Why: the gather is 96 B per (lookup, nuclide) at a random grid point of that nuclide.  If
// every warp works on its own lookup front to back, the whole grid is the working set and
// (measured, profiles/r01_notes.md) almost every pair comes from DRAM.  Here ONE LAUNCH only
// touches the nuclides of one window, which stays resident in the 126 MB L2 while all lookups
// of the material stream past it; DRAM only streams the index rows.  A material with more
// nuclides than one window (fuel: 321) takes several launches; the five partial sums of a
// lookup travel between launches through a 48-byte record in global memory (2 x 113 MB per
// pass for fuel -- noise next to the gather).  Stream order is the only synchronisation.
// Materials that fit one window (all but fuel) share ONE launch: a launch works on up to 12
// segments (material, slot range, window).
//
// Layout: the gather reads PAIR RECORDS built once at init (xs_build_pairs_kernel), one
// 128-byte line per (nuclide, k), holding for grid points lo = k and hi = k+1:
//     quarter 0: hi.total, hi.total-lo.total, hi.elastic, hi.elastic-lo.elastic
//     quarter 1: hi.absorbtion, d(absorbtion), hi.fission, d(fission)
//     quarter 2: hi.nu_fission, d(nu_fission), lo.energy, 0
//     quarter 3: hi.energy, d = hi.energy-lo.energy, 1/d, 0
// The differences are the reference's own intermediate values (hi - lo rounded once), so
// xs = hi - f*(hi-lo) is computed with the reference's roundings.
//
// Mapping: a warp owns 8 lookups (slots) x 4 lanes; lane q of a slot loads quarter q of the
// slot's record with ONE 256-bit load: 128 contiguous bytes per slot, one L1 wavefront per
// (lookup, nuclide).  Lane 3 derives the interpolation factor f = (hi.E - E)/d from the stored
// reciprocal with one Newton-Markstein correction (q = n*inv; r = fma(-d,q,n); f = fma(r,inv,q)
// -- the correctly rounded quotient, see tests/test_gpu_parity.py::test_division_is_exact) and
// broadcasts it; lanes 0..2 accumulate two channels each in reference order j = 0,1,2,....
// The record numbers of a whole window are resolved up front by all 32 lanes (coalesced
// index-row segments) and staged in shared memory; the gather loop is software-pipelined
// (2 x kSweepUnroll loads in flight per lane) and free of predication (padded steps carry
// concentration 0).
// ---------------------------------------------------------------------------------------
#ifndef XS_SWEEP_UNROLL
#define XS_SWEEP_UNROLL 2
#endif
#ifndef XS_SWEEP_BLOCKS
#define XS_SWEEP_BLOCKS 4
#endif
constexpr int kSweepUnroll = XS_SWEEP_UNROLL;
constexpr int kSweepSlots = 8;
constexpr int kSweepQuantum = 2 * kSweepUnroll;                       // steps per trip of the gather loop
constexpr int kFoldShift = kSweepQuantum <= 4 ? 2 : 3;                // staging row of a folded remainder
constexpr int kMaxWindow = 32 + (1 << kFoldShift);   // nuclides per window (staging capacity: 32 + one folded quantum)
#ifndef XS_SWEEP_FENCE
#define XS_SWEEP_FENCE 1
#endif
constexpr int kMaxSegments = 12;

struct WindowSegment {
    long offset;               // first slot of this material in the grouped arrays
    int  count;                // lookups of this material (a batch holds < 2^31)
    int  group_begin;          // first warp-group of this segment inside the launch
    int  mat;                  // material of this segment
    int  first;                // mat_first[material]
    int  j_begin, j_end;       // nuclide window inside the material's list
};

struct SegTable;
struct WindowArgs {
    const SegTable *dev_table; // lane-per-lookup kernels: segments computed ON THE DEVICE from the material histogram
                               // (xs_build_segments_kernel) -- the host then never waits for the histogram; null = seg[] below
    const double   *energy;    // energies grouped by material
    const uint32_t *where;     // same order: UEG row / hash bin
    const uint32_t *sample_id; // same order: original sample index (only for macro_xs dumps)
    double2        *partial;   // [3 * slots] partial sums between windows
    int   n_groups;            // warp-groups in this launch
    int   n_seg;
    int   first_window, last_window;
    int   indirect;            // sorted kernel: energy / where are the UNGROUPED sample arrays, read through sample_id
    const double2 *pack;       // indirect mode: (energy, row as the low word of .y) per sample -- one sector per random read
    WindowSegment seg[kMaxSegments];
};

struct Quarter { double a, da, b, db; };

// Segment table of one launch of the lane-per-lookup kernels, in device memory (see WindowArgs::dev_table).
struct SegTable {
    int n_groups, n_seg;
    WindowSegment seg[kMaxSegments];
};

// The table a block works from: copied to shared memory from the device table or from the arguments.
XS_DEV void load_seg_table(const WindowArgs &A, SegTable &T)
{
    if (A.dev_table) {
        if (threadIdx.x == 0) { T.n_groups = A.dev_table->n_groups; T.n_seg = A.dev_table->n_seg; }
        if (threadIdx.x < kMaxSegments) T.seg[threadIdx.x] = A.dev_table->seg[threadIdx.x];
    } else {
        if (threadIdx.x == 0) { T.n_groups = A.n_groups; T.n_seg = A.n_seg; }
        if (threadIdx.x < kMaxSegments) T.seg[threadIdx.x] = A.seg[threadIdx.x];
    }
}

struct MatShape { int first[kNumMaterials + 1]; int n_nuc[kNumMaterials]; };

// Device-side replacement of the host loop in launch_sorted: which materials are dense (>= dense_min lookups
// per grid interval), their slot ranges in the sorted batch (prefix of the histogram: the sort key's top
// bits are the material) and their warp-groups.  One thread; 12 materials.
// (Narrower warp-groups for materials under the threshold -- 1 or 2 lookups per lane, so that a group again spans
// few records -- were measured in round 2 and lost to xs_sorted_kernel at every size: profiles/r02_notes.md.)
__global__ void xs_build_segments_kernel(const unsigned int *hist, MatShape shape, long dense_threshold, int dense_group,
                                         int sparse_group, SegTable *dense, SegTable *sparse)
{
    if (threadIdx.x || blockIdx.x) return;
    SegTable *tab[2] = { sparse, dense };
    int groups[2] = { 0, 0 }, n_seg[2] = { 0, 0 };
    long offset = 0;
    for (int m = 0; m < kNumMaterials; m++) {
        const long count = hist[m];
        if (count > 0) {
            const int d = dense_threshold > 0 && count >= dense_threshold;
            WindowSegment &sgm = tab[d]->seg[n_seg[d]++];
            sgm.offset = offset; sgm.count = (int)count; sgm.group_begin = groups[d];
            sgm.mat = m; sgm.first = shape.first[m]; sgm.j_begin = 0; sgm.j_end = shape.n_nuc[m];
            const int per = d ? dense_group : sparse_group;
            groups[d] += (int)((count + per - 1) / per);
        }
        offset += count;
    }
    for (int d = 0; d < 2; d++) { tab[d]->n_groups = groups[d]; tab[d]->n_seg = n_seg[d]; }
}


// Asynchronous global->shared copies (LDGSTS): the next group's sample is fetched without
// tying up registers for the duration of the current group.
XS_DEV void cp_async_8(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
XS_DEV void cp_async_4(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
XS_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

XS_DEV Quarter ldg_quarter(const double2 *p)
{
    Quarter v;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v.a), "=d"(v.da), "=d"(v.b), "=d"(v.db) : "l"(p));
    return v;
}

// One step: lane 3 turns (hi.E, d, 1/d) into f and broadcasts it; every lane folds its two
// channels.  conc == 0 for padded steps.
XS_DEV double sweep_step(const Quarter &v, double e, double conc, int f_src, double &acc_x, double &acc_y)
{
    const double n = v.a - e;                          // hi.E - E          (lane 3)
    const double q = n * v.b;                          // n * (1/d)
    const double r = __fma_rn(-v.da, q, n);            // n - d*q, exact
    const double f_own = __fma_rn(r, v.b, q);          // correctly rounded n/d
    const double f = __shfl_sync(kFullMask, f_own, f_src);
    acc_x += (v.a - f * v.da) * conc;
    acc_y += (v.b - f * v.db) * conc;
    return f;
}

// ptxas hoists the first consumer of a just-issued record load above the arithmetic of the
// previous half of the loop, which turns the two-deep software pipeline into a one-deep one
// (ncu source page: the warp then sits on that DADD for a full L2 round trip).  Making the
// energy operand of the next half depend on a result of this half pins the order: fma(0, x, e)
// is e exactly (x finite, e >= 0) and cannot be folded.  Measured: energy-sorted lookups
// (-k 6) 8.76 -> 8.13 ms, unsorted (-k 4, bound by L2->SM gather throughput) unchanged.
XS_DEV double order_after(double e, double x)
{
#if XS_SWEEP_FENCE
    return __fma_rn(0.0, x, e);
#else
    return e;
#endif
}

// Resolve the record numbers of one window for the 8 lookups of a warp.  The 8 x n_steps
// (slot, step) pairs are spread over the 32 lanes, ROW = 2^ROW_SHIFT (>= n_steps) pairs per
// slot: a 4-nuclide material needs one round, a 32-nuclide window eight.  A lane keeps its step
// j in every round; rounds walk over the slots.  All index loads are issued before the first
// one is consumed.  Steps [jn, n_steps) are padding: record 0 (concentration 0).
template <int GRID, int ROW_SHIFT, bool WIDE = false>
XS_DEV void stage_records(const Problem &P, uint32_t (*rec_rows)[kMaxWindow + 1], const int *nucs, int jn, int n_steps,
                          int slots_on, uint32_t where32, double e, int lane, int j_off = 0)
{
    constexpr int kPerRound = 32 >> ROW_SHIFT;
    constexpr int kRounds = kSweepSlots / kPerRound;
    const int j = j_off + (lane & ((1 << ROW_SHIFT) - 1));
    const int s0 = lane >> ROW_SHIFT;
    const int nuc = j < jn ? nucs[j] : 0;
    int low[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; r++) {
        const int s = s0 + r * kPerRound;
        const uint32_t w_s = __shfl_sync(kFullMask, where32, 4 * s);
        double e_s = 0.0;
        if (GRID != kUnionized) e_s = __shfl_sync(kFullMask, e, 4 * s);
        low[r] = 0;
        if (s < slots_on && j < jn) low[r] = nuclide_low<GRID, GRID == kHash, WIDE>(P, e_s, (long)w_s, nuc);   // hash: probe the records (measured faster); nuclide: the compact grid
    }
#pragma unroll
    for (int r = 0; r < kRounds; r++) {
        const int s = s0 + r * kPerRound;
        if (j < n_steps) rec_rows[s][j] = (s < slots_on && j < jn) ? (uint32_t)((long)nuc * P.n_gp + low[r]) : 0u;
    }
}

// One window of one warp-group: 8 lookups x 4 lanes sweep the nuclides nucs[0..jn) (jn <= kMaxWindow).
// `e` / `where32` are the slot's sample (the same on its 4 lanes), `ci` the index of the window's
// first concentration in C.v; acc_x / acc_y carry the two channels of lane `quarter` in and out.
// Shared by xs_window_kernel (lookups grouped in global memory, one window per launch) and
// xs_tile_kernel (-k 0..3: lookups grouped per tile in shared memory, all windows in one launch).
template <int GRID, bool WIDE = false>
XS_DEV void window_sweep_group(const Problem &P, const ConcTable &C, uint32_t (*rec_rows)[kMaxWindow + 1], const int *nucs,
                               int jn, int ci, int slots_on, uint32_t where32, double e, int lane, double &acc_x, double &acc_y)
{
    const int slot = lane >> 2, quarter = lane & 3;
    const int f_src = lane | 3;
    const double2 *my_pairs = P.pairs + 2 * quarter;
    const int n_steps = (jn + 2 * kSweepUnroll - 1) / (2 * kSweepUnroll) * (2 * kSweepUnroll);
    // ---- resolve the record numbers of the window for the 8 lookups of this warp ------
    __syncwarp();
    {
        if (n_steps <= 4)       stage_records<GRID, 2, WIDE>(P, rec_rows, nucs, jn, n_steps, slots_on, where32, e, lane);
        else if (n_steps <= 8)  stage_records<GRID, 3, WIDE>(P, rec_rows, nucs, jn, n_steps, slots_on, where32, e, lane);
        else if (n_steps <= 16) stage_records<GRID, 4, WIDE>(P, rec_rows, nucs, jn, n_steps, slots_on, where32, e, lane);
        else {
            stage_records<GRID, 5, WIDE>(P, rec_rows, nucs, jn, n_steps, slots_on, where32, e, lane);
            if (n_steps > 32)   // a folded remainder: steps 32..32 + 2^kFoldShift - 1
                stage_records<GRID, kFoldShift, WIDE>(P, rec_rows, nucs, jn, n_steps, slots_on, where32, e, lane, 32);
        }
    }
    __syncwarp();

    // ---- software-pipelined gather: no predication, padded steps multiply by 0 ----------
    const uint32_t *my_rec = rec_rows[slot];
    Quarter A0[kSweepUnroll], A1[kSweepUnroll];
#pragma unroll
    for (int u = 0; u < kSweepUnroll; u++) A0[u] = ldg_quarter(my_pairs + 8 * (long)my_rec[u]);
    int j0 = 0;
    double e0 = e, e1 = e;
    for (; j0 + 2 * kSweepUnroll < n_steps; j0 += 2 * kSweepUnroll) {
#pragma unroll
        for (int u = 0; u < kSweepUnroll; u++)
            A1[u] = ldg_quarter(my_pairs + 8 * (long)my_rec[j0 + kSweepUnroll + u]);
#pragma unroll
        for (int u = 0; u < kSweepUnroll; u++)
            sweep_step(A0[u], e0, C.v[ci + j0 + u], f_src, acc_x, acc_y);
        e1 = order_after(e, acc_x);
#pragma unroll
        for (int u = 0; u < kSweepUnroll; u++)
            A0[u] = ldg_quarter(my_pairs + 8 * (long)my_rec[j0 + 2 * kSweepUnroll + u]);
#pragma unroll
        for (int u = 0; u < kSweepUnroll; u++)
            sweep_step(A1[u], e1, C.v[ci + j0 + kSweepUnroll + u], f_src, acc_x, acc_y);
        e0 = order_after(e, acc_x);
    }
    {   // last iteration: nothing left to prefetch
#pragma unroll
        for (int u = 0; u < kSweepUnroll; u++)
            A1[u] = ldg_quarter(my_pairs + 8 * (long)my_rec[j0 + kSweepUnroll + u]);
#pragma unroll
        for (int u = 0; u < kSweepUnroll; u++)
            sweep_step(A0[u], e0, C.v[ci + j0 + u], f_src, acc_x, acc_y);
        e1 = order_after(e, acc_x);
#pragma unroll
        for (int u = 0; u < kSweepUnroll; u++)
            sweep_step(A1[u], e1, C.v[ci + j0 + kSweepUnroll + u], f_src, acc_x, acc_y);
    }
}

template <int GRID>
__global__ void __launch_bounds__(kBlockThreads, XS_SWEEP_BLOCKS)
xs_window_kernel(const Problem P, const WindowArgs A, const BatchSink sink, const ConcTable C)
{
    __shared__ unsigned long long s_part[kWarpsPerBlock];
    __shared__ uint32_t s_rec[kWarpsPerBlock][kSweepSlots][kMaxWindow + 1];
    __shared__ double s_next_e[kBlockThreads];
    __shared__ uint32_t s_next_where[kBlockThreads];
    extern __shared__ int s_nuc[];                           // [mat_total] nuclide ids of all materials
    for (int i = threadIdx.x; i < P.mat_total; i += blockDim.x) s_nuc[i] = P.mat_nuc[i];
    __syncthreads();
    const int n_groups = A.n_groups;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = lane >> 2, quarter = lane & 3;
    const int warp_global = blockIdx.x * kWarpsPerBlock + warp;
    const int warp_stride = gridDim.x * kWarpsPerBlock;
    unsigned int my_sum = 0;                                 // per thread: far below 2^32

    // The sample (energy, row) of a warp's NEXT group is fetched into shared memory with
    // cp.async while the current group is processed.
    auto fetch_sample = [&](int gq, int seg_hint) {
        bool valid = false;
        if (gq < n_groups) {
            int sq = seg_hint;
            while (sq + 1 < A.n_seg && gq >= A.seg[sq + 1].group_begin) sq++;
            const int in_q = (gq - A.seg[sq].group_begin) * kSweepSlots + slot;
            if (in_q < A.seg[sq].count) {
                valid = true;
                cp_async_8(&s_next_e[threadIdx.x], A.energy + A.seg[sq].offset + in_q);
                cp_async_4(&s_next_where[threadIdx.x], A.where + A.seg[sq].offset + in_q);
            }
        }
        if (!valid) { s_next_e[threadIdx.x] = 0.5; s_next_where[threadIdx.x] = 0u; }
    };
    fetch_sample(warp_global, 0);

    int sg = 0;                                              // segments are visited in order
    for (int g = warp_global; g < n_groups; g += warp_stride) {
        while (sg + 1 < A.n_seg && g >= A.seg[sg + 1].group_begin) sg++;      // warp-uniform
        const WindowSegment &S = A.seg[sg];
        const int in_seg = (g - A.seg[sg].group_begin) * kSweepSlots + slot;
        const long t = A.seg[sg].offset + in_seg;                // global slot
        const bool on = in_seg < A.seg[sg].count;
        const int jn = S.j_end - S.j_begin;

        cp_async_wait_all();
        const double e = s_next_e[threadIdx.x];
        const uint32_t where32 = s_next_where[threadIdx.x];
        fetch_sample(g + warp_stride, sg);                   // next group's sample, asynchronously
        double acc_x = 0.0, acc_y = 0.0;
        if (on && !A.first_window && quarter < 3) {
            const double2 part = A.partial[3 * t + quarter];
            acc_x = part.x;
            acc_y = part.y;
        }

        window_sweep_group<GRID>(P, C, s_rec[warp], s_nuc + S.first + S.j_begin, jn, C.first[S.mat] + S.j_begin,
                                 min(kSweepSlots, A.seg[sg].count - (g - A.seg[sg].group_begin) * kSweepSlots), where32, e, lane,
                                 acc_x, acc_y);

        if (!A.last_window) {
            if (on && quarter < 3) A.partial[3 * t + quarter] = make_double2(acc_x, acc_y);
            continue;                                        // warp-uniform
        }
        // lane 0 = (total, elastic)  lane 1 = (absorbtion, fission)  lane 2 = (nu_fission, -)
        const double q1x = __shfl_down_sync(kFullMask, acc_x, 1), q1y = __shfl_down_sync(kFullMask, acc_y, 1);
        const double q2x = __shfl_down_sync(kFullMask, acc_x, 2);
        if (on && quarter == 0) {
            const double v[5] = {acc_x, acc_y, q1x, q1y, q2x};
            double gap;
            const int am = argmax5(v, gap);
            my_sum += (unsigned int)(am + 1);
            if (sink.macro_out) {
                const long id = A.sample_id ? (long)A.sample_id[t] : t;
#pragma unroll
                for (int k = 0; k < 5; k++) sink.macro_out[5 * id + k] = v[k];
            }
            if (sink.fwd_out) {            // history mode feedback (openmp-threading/Simulation.c:225-228)
                const long id = A.sample_id ? (long)A.sample_id[t] : t;
                int fwd = 0;
#pragma unroll
                for (int k = 0; k < 5; k++) fwd += v[k] > 1.0;
                sink.fwd_out[id] = (unsigned char)fwd;
            }
        }
    }
    const unsigned long long bs = block_sum(my_sum, s_part);
    if (threadIdx.x == 0) {
        if (bs) atomicAdd(sink.accum, bs);
        if (blockIdx.x == 0 && A.last_window) {              // lookups completed by this launch
            unsigned long long done = 0;
            for (int i = 0; i < A.n_seg; i++) done += (unsigned long long)A.seg[i].count;
            atomicAdd(sink.accum + 1, done);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Lane-per-lookup sweep for ENERGY-SORTED lookups (-k 6).
//
// Once the lookups of a material are sorted by energy, the 32 lookups of a warp sit in the same
// or neighbouring grid intervals of every nuclide (large/fuel: ~209 lookups per interval), i.e.
// they need the SAME pair record.  The windowed kernel above still moves 128 bytes per (lookup,
// nuclide) through the L1 data stage -- 8 wavefronts per warp-load whatever the addresses, and
// that is why sorting did not make it faster.  Here every lane owns one lookup and all five
// channels; the lanes read the record with (mostly) identical addresses, which the L1 serves
// as a broadcast: a 256-bit warp-load costs 3.4 cycles instead of 8 (scripts/exp/bcast_bench.cu),
// i.e. 4 loads = 13 cycles per 32 (lookup, nuclide) pairs instead of 8 per 8.  No cross-lane
// traffic at all in the gather loop: f, the five interpolations and the argmax are per lane,
// in the reference's order and with the reference's roundings (same operations as sweep_step).
//
// All nuclides of a material are swept in one pass (no windows, no partial sums): at any moment
// the resident warps cover a narrow energy band, so the records in flight fit L1/L2 by
// construction.  Unionized grid: the index entries of 32 lookups x 32 nuclides are read as 32
// coalesced row segments and transposed through shared memory.
// ---------------------------------------------------------------------------------------
#ifndef XS_SORTED_BLOCKS
#define XS_SORTED_BLOCKS 2
#endif
#ifndef XS_SORTED_PER_LANE
#define XS_SORTED_PER_LANE 2
#endif
// A lane owns kPerLane CONSECUTIVE lookups: they almost always share the record too (large/fuel:
// 99.5 % per nuclide), so one record load serves them all and the bytes moved into registers
// per (lookup, nuclide) -- the L1 data stage is the busiest unit of this kernel -- drop by
// kPerLane; a lane whose next lookup sits in the next grid interval reloads (rare, divergent).
#ifndef XS_SORTED_STAGE_PF
#define XS_SORTED_STAGE_PF 1       // prefetch a chunk's records into L2 at staging time (3.62 vs 4.07 ms; an in-loop L1 prefetch made it slower)
#endif
constexpr int kPerLane = XS_SORTED_PER_LANE;
XS_DEV void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
constexpr int kSortedGroup = 32 * kPerLane;        // lookups per warp-group
#ifndef XS_SORTED_CHUNK
#define XS_SORTED_CHUNK 32
#endif
constexpr int kChunk = XS_SORTED_CHUNK;            // nuclides (steps) staged at a time: 32, 16 or 8
constexpr int kChunkShift = kChunk == 32 ? 5 : kChunk == 16 ? 4 : 3;
constexpr int kLaneWords = kChunk * kPerLane + 1;  // staged record numbers of one lane: [which][step] + pad (bank = lane + step)

#ifndef XS_SORTED_STAGE_UNROLL
#define XS_SORTED_STAGE_UNROLL 32
#endif
#ifndef XS_SORTED_RING
#define XS_SORTED_RING 8
#endif
constexpr int kStageUnroll = XS_SORTED_STAGE_UNROLL;   // lookups (x kPerLane index loads) in flight while staging a chunk
constexpr int kRing = XS_SORTED_RING;              // steps of records in flight per warp (cp.async ring in shared memory)
#ifndef XS_SORTED_SPAN
#define XS_SORTED_SPAN 2
#endif
constexpr int kSpan = XS_SORTED_SPAN;              // consecutive records per nuclide held in the ring (2, 4 or 8)
constexpr int kSlotBytes = kSpan * 128;
constexpr int kRingBytes = kRing * kSlotBytes;     // per warp: [step % kRing][record of the group's first lookup + 0..kSpan-1][128 B]

struct PairRecord { double hi[5], dlt[5], hi_e, d, inv, pad; };

XS_DEV void cp_async_16(uint32_t smem_addr, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_addr), "l"(gmem));
}
XS_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> XS_DEV void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// A record from the ring: all lanes of a warp read the same 128 bytes (broadcast, 7 LDS.128).
XS_DEV PairRecord lds_record(uint32_t smem_addr)
{
    PairRecord r;
    // (hi.E, d, 1/d) first: the interpolation factor is the head of the dependency chain
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+96];" : "=d"(r.hi_e), "=d"(r.d) : "r"(smem_addr));
    asm volatile("ld.shared.f64 %0, [%1+112];" : "=d"(r.inv) : "r"(smem_addr));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r.hi[0]), "=d"(r.dlt[0]) : "r"(smem_addr));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+16];" : "=d"(r.hi[1]), "=d"(r.dlt[1]) : "r"(smem_addr));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+32];" : "=d"(r.hi[2]), "=d"(r.dlt[2]) : "r"(smem_addr));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+48];" : "=d"(r.hi[3]), "=d"(r.dlt[3]) : "r"(smem_addr));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+64];" : "=d"(r.hi[4]), "=d"(r.dlt[4]) : "r"(smem_addr));
    r.pad = 0.0;
    return r;
}

XS_DEV PairRecord ldg_record(const double2 *rec)
{
    PairRecord r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.hi[0]), "=d"(r.dlt[0]), "=d"(r.hi[1]), "=d"(r.dlt[1]) : "l"(rec));
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.hi[2]), "=d"(r.dlt[2]), "=d"(r.hi[3]), "=d"(r.dlt[3]) : "l"(rec + 2));
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(r.hi[4]), "=d"(r.dlt[4]) : "l"(rec + 4));
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.hi_e), "=d"(r.d), "=d"(r.inv), "=d"(r.pad) : "l"(rec + 6));
    return r;
}

XS_DEV void record_step(const PairRecord &r, double e, double conc, double acc[5])
{
    const double n = r.hi_e - e;
    const double q = n * r.inv;
    const double rem = __fma_rn(-r.d, q, n);
    const double f = __fma_rn(rem, r.inv, q);          // correctly rounded (hi.E - E) / d
#pragma unroll
    for (int k = 0; k < 5; k++) acc[k] += (r.hi[k] - f * r.dlt[k]) * conc;
}

// ---- shared by the two lane-per-lookup kernels (xs_sorted_kernel, xs_dense_kernel) ----

// The segment (material) a warp-group belongs to: warp-uniform, <= 11 steps.
XS_DEV int segment_of_group(const SegTable &A, int g, int sg = 0)
{
    while (sg + 1 < A.n_seg && g >= A.seg[sg + 1].group_begin) sg++;
    return sg;
}

// Warp-groups are handed out by an atomic counter, in order: fuel (321 steps per group) comes
// first, the cheap materials fill the tail.  A static split leaves 16-vs-15 fuel groups per warp
// and cost 12 % of the kernel.  sink.batch_counter: [0] next group, [1] warps that are done -- the
// last warp of a launch re-arms both, so launches on one stream share the two words.
XS_DEV int warp_next_group(const BatchSink &sink, int lane)
{
    int v = 0;
    if (lane == 0) v = (int)atomicAdd(sink.batch_counter, 1u);
    return __shfl_sync(kFullMask, v, 0);
}
XS_DEV void warp_groups_done(const BatchSink &sink, int lane)
{
    if (lane == 0) {
        __threadfence();
        if (atomicAdd(sink.batch_counter + 1, 1u) == gridDim.x * kWarpsPerBlock - 1) {
            sink.batch_counter[0] = 0;
            sink.batch_counter[1] = 0;
            __threadfence();
        }
    }
}

// A lane's PL consecutive lookups of a group: energy and UEG row / hash bin.  Slots past
// the end of the segment repeat the lookup in slot `idle_slot` (their results are dropped).
// indirect: the sort's permutation is applied here instead of by a gather pass (two random reads
// per lookup either way; -0.5 ms of gather kernel, +0.25 ms in here).
template <int PL>
XS_DEV void load_lane_samples(const WindowArgs &A, const WindowSegment &S, int first_in_seg, long idle_slot,
                              double e[PL], uint32_t where32[PL], bool on[PL])
{
#pragma unroll
    for (int w = 0; w < PL; w++) {
        on[w] = first_in_seg + w < S.count;
        const long t = on[w] ? S.offset + first_in_seg + w : idle_slot;
        const long src = A.indirect ? (long)A.sample_id[t] : t;
        if (A.indirect && A.pack) {
            const double2 s = __ldg(A.pack + src);
            e[w] = s.x;
            where32[w] = (uint32_t)__double_as_longlong(s.y);
        } else {
            e[w] = A.energy[src];
            where32[w] = A.where[src];
        }
    }
}

// A lane's finished lookups: argmax into the checksum, optional macro_xs dump, optional history
// feedback n_forward = #{k : macro_xs[k] > 1.0} (openmp-threading/Simulation.c:225-228).
template <int PL>
XS_DEV void finish_lane_lookups(const WindowArgs &A, const BatchSink &sink, long t0, const bool on[PL],
                                const double acc[PL][5], unsigned int &my_sum)
{
#pragma unroll
    for (int w = 0; w < PL; w++) {
        if (!on[w]) continue;
        double gap;
        const int am = argmax5(acc[w], gap);
        my_sum += (unsigned int)(am + 1);
        if (sink.macro_out) {
            const long id = A.sample_id ? (long)A.sample_id[t0 + w] : t0 + w;
#pragma unroll
            for (int k = 0; k < 5; k++) sink.macro_out[5 * id + k] = acc[w][k];
        }
        if (sink.fwd_out) {
            const long id = A.sample_id ? (long)A.sample_id[t0 + w] : t0 + w;
            int fwd = 0;
#pragma unroll
            for (int k = 0; k < 5; k++) fwd += acc[w][k] > 1.0;
            sink.fwd_out[id] = (unsigned char)fwd;
        }
    }
}

// Block's share of the verification sum, and (block 0) the lookups this launch completes.
XS_DEV void finish_launch(const SegTable &A, const BatchSink &sink, unsigned int my_sum, unsigned long long *s_part)
{
    const unsigned long long bs = block_sum(my_sum, s_part);
    if (threadIdx.x == 0) {
        if (bs) atomicAdd(sink.accum, bs);
        if (blockIdx.x == 0) {
            unsigned long long done = 0;
            for (int i = 0; i < A.n_seg; i++) done += (unsigned long long)A.seg[i].count;
            atomicAdd(sink.accum + 1, done);
        }
    }
}

template <int GRID>
__global__ void __launch_bounds__(kBlockThreads, XS_SORTED_BLOCKS)
xs_sorted_kernel(const Problem P, const WindowArgs A, const BatchSink sink, const ConcTable C)
{
    constexpr bool kStaged = GRID == kUnionized;
    __shared__ unsigned long long s_part[kWarpsPerBlock];
    extern __shared__ __align__(128) uint32_t s_dyn[];       // [record rings][staged record numbers][nuclide ids]
    uint32_t *s_rec = s_dyn + (kStaged ? kWarpsPerBlock * kRingBytes / 4 : 0);
    int *s_nuc = (int *)(s_rec + (kStaged ? kWarpsPerBlock * 32 * kLaneWords : 0));
    __shared__ SegTable T;
    load_seg_table(A, T);
    for (int i = threadIdx.x; i < P.mat_total; i += blockDim.x) s_nuc[i] = P.mat_nuc[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int my_sum = 0;
    uint32_t *warp_rec = s_rec + (kStaged ? warp * 32 * kLaneWords : 0);
    const uint32_t *my_rec = warp_rec + lane * kLaneWords;   // [which * kChunk + step]
    const uint32_t *first_rec = warp_rec;                                        // lookup 0 of the group
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(s_dyn) + (kStaged ? warp * kRingBytes : 0);

    // groups are handed out in order by an atomic counter (see warp_next_group)
    for (int g = warp_next_group(sink, lane); g < T.n_groups; g = warp_next_group(sink, lane)) {
        const int sg = segment_of_group(T, g);
        const WindowSegment &S = T.seg[sg];
        const int first_in_seg = (g - S.group_begin) * kSortedGroup + lane * kPerLane;
        const long t0 = S.offset + first_in_seg;
        double e[kPerLane];
        uint32_t where32[kPerLane];
        bool on[kPerLane];
        // idle slots repeat the segment's last lookup: the group's last lookup then still bounds
        // the records of all the others
        load_lane_samples<kPerLane>(A, S, first_in_seg, S.offset + S.count - 1, e, where32, on);
        const int n_nuc = S.j_end;                            // whole material (j_begin = 0)
        const int ci = C.first[S.mat];
        double acc[kPerLane][5];
#pragma unroll
        for (int w = 0; w < kPerLane; w++)
#pragma unroll
            for (int k = 0; k < 5; k++) acc[w][k] = 0.0;

        for (int c0 = 0; c0 < n_nuc; c0 += kChunk) {
            const int jn = min(kChunk, n_nuc - c0);
            const int n_steps = (jn + 1) & ~1;               // an odd tail is padded: concentration 0
            const int *nucs = s_nuc + S.first + c0;
            if constexpr (kStaged) {
                // record numbers of 32*kPerLane lookups x 32 steps: one coalesced index-row segment
                // per lookup, transposed through shared memory.  Columns >= jn resolve to a valid
                // record of nuclide 0 (never an out-of-range address, never used with conc != 0).
                __syncwarp();
                // Short chunks (small materials, the tail of fuel) pack several lookups into one
                // load instruction: COLS = 4 / 8 / 16 columns x 8 / 4 / 2 lookups per warp-load.
                auto stage_packed = [&](auto shift_tag) {
                    constexpr int kShift = decltype(shift_tag)::value;
                    constexpr int kCols = 1 << kShift, kRows = 32 >> kShift;
                    const int colj = lane & (kCols - 1), sub = lane >> kShift;
                    const int nuc_l = colj < jn ? nucs[colj] : 0;
                    const int *col = P.index_grid + nuc_l;
                    const uint32_t base_l = (uint32_t)nuc_l * (uint32_t)P.n_gp;
#pragma unroll
                    for (int k = 0; k < kSortedGroup / kRows; k++) {
                        const int i = k * kRows + sub;                       // lookup of the group
                        const int owner = i / kPerLane, w = i % kPerLane;
                        uint32_t w_i = 0;
#pragma unroll
                        for (int ww = 0; ww < kPerLane; ww++) {
                            const uint32_t cand = __shfl_sync(kFullMask, where32[ww], owner);
                            if (ww == w) w_i = cand;
                        }
                        const uint32_t no = base_l + (uint32_t)ldg_index_stream(col + (size_t)w_i * (uint32_t)P.n_iso);
                        warp_rec[owner * kLaneWords + w * kChunk + colj] = no;
                        if (XS_SORTED_STAGE_PF && (i == 0 || i == kSortedGroup - 1)) prefetch_l2(P.pairs + 8 * (size_t)no);
                    }
                };
                if (jn <= 4)                      stage_packed(std::integral_constant<int, 2>());
                else if (jn <= 8)                 stage_packed(std::integral_constant<int, 3>());
                else if (jn <= 16 || kChunk < 32) stage_packed(std::integral_constant<int, (kChunkShift < 4 ? kChunkShift : 4)>());
                else if constexpr (kChunk == 32) {
                const int nuc_l = lane < jn ? nucs[lane] : 0;
                const int *col = P.index_grid + nuc_l;
                const uint32_t base_l = (uint32_t)nuc_l * (uint32_t)P.n_gp;
#pragma unroll kStageUnroll
                for (int l = 0; l < 32; l++) {
#pragma unroll
                    for (int w = 0; w < kPerLane; w++) {
                        const uint32_t w_i = __shfl_sync(kFullMask, where32[w], l);
                        const uint32_t no = base_l + (uint32_t)ldg_index_stream(col + (size_t)w_i * (uint32_t)P.n_iso);
                        warp_rec[l * kLaneWords + w * kChunk + lane] = no;
                        // A record is used by ~one block only (neighbouring energies), so its first
                        // touch comes from DRAM: request the records of the group's first and last
                        // lookup (the others lie in between) for all 32 steps at once, now.
                        if (XS_SORTED_STAGE_PF && ((l == 0 && w == 0) || (l == 31 && w == kPerLane - 1)))
                            prefetch_l2(P.pairs + 8 * (size_t)no);
                    }
                }
                }
                __syncwarp();
                if (c0 + kChunk < n_nuc) {
                    // index-row segments of the next chunk: start their trip from DRAM now
                    const int last_col = min(2 * kChunk - 1, n_nuc - c0 - 1);
#pragma unroll
                    for (int w = 0; w < kPerLane; w++) {
                        const int *row = P.index_grid + (size_t)where32[w] * (uint32_t)P.n_iso;
                        prefetch_l2(row + nucs[kChunk]);
                        prefetch_l2(row + nucs[last_col]);
                    }
                }

                // ---- gather through the ring: kSpan consecutive records per nuclide, starting at
                // the group's first lookup's, are copied to shared memory kRing steps ahead
                // (cp.async, 16 B per lane) and read back as broadcasts; a lookup further ahead
                // than that (a group spanning more than kSpan grid intervals of a nuclide: sparse
                // materials, XL grids) loads its record directly.
                auto issue = [&](int s) {                    // steps s, s+1 (s even)
                    // The lookups of a group are sorted, so the records they need of one nuclide
                    // are CONSECUTIVE: copy kSpan of them starting at the first lookup's.
                    constexpr int kLanesPerStep = 8 * kSpan;                 // 16-byte pieces of a slot
                    constexpr int kStepsPerInstr = kLanesPerStep >= 32 ? 1 : 32 / kLanesPerStep;
#pragma unroll
                    for (int q = 0; q < 2 / kStepsPerInstr; q++) {
#pragma unroll
                        for (int part = 0; part < (kLanesPerStep + 31) / 32; part++) {
                            const int step = s + q * kStepsPerInstr + (kStepsPerInstr == 2 ? lane >> 4 : 0);
                            const int piece = part * 32 + (kStepsPerInstr == 2 ? lane & 15 : lane);
                            if (step < n_steps)
                                cp_async_16(ring + (uint32_t)((step % kRing) * kSlotBytes + piece * 16),
                                            P.pairs + 8 * (size_t)first_rec[step] + piece);
                        }
                    }
                    cp_async_commit();
                };
#pragma unroll
                for (int s = 0; s < kRing; s += 2) issue(s);
                for (int j = 0; j < n_steps; j += 2) {
                    cp_async_wait_group<kRing / 2 - 1>();
                    __syncwarp();
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int step = j + h;
                        const double conc = C.v[ci + c0 + step];
                        const uint32_t no_first = first_rec[step];
                        const uint32_t slot = ring + (uint32_t)((step % kRing) * kSlotBytes);
                        PairRecord r;
                        uint32_t have = 0xffffffffu;
#pragma unroll
                        for (int w = 0; w < kPerLane; w++) {
                            const uint32_t no = my_rec[w * kChunk + step];
                            if (no != have) {
                                const uint32_t ahead = no - no_first;        // 0, 1, ... in a sorted group
                                if (ahead < (uint32_t)kSpan) r = lds_record(slot + ahead * 128);
                                else                         r = ldg_record(P.pairs + 8 * (size_t)no);
                                have = no;
                            }
                            record_step(r, e[w], conc, acc[w]);
                        }
                    }
                    __syncwarp();                            // everyone is done with these two slots
                    issue(j + kRing);
                }
                cp_async_wait_group<0>();
            } else {
                auto record_no = [&](int j, int w) -> uint32_t {
                    const int nuc = nucs[min(j, jn - 1)];
                    return (uint32_t)nuc * (uint32_t)P.n_gp + (uint32_t)nuclide_low<GRID, false>(P, e[w], (long)where32[w], nuc);
                };
                // one step of all the lane's lookups; `r` holds the record of lookup 0 on entry
                auto lane_step = [&](PairRecord &r, uint32_t no, int j) {
                    const double conc = C.v[ci + c0 + j];
                    record_step(r, e[0], conc, acc[0]);
#pragma unroll
                    for (int w = 1; w < kPerLane; w++) {
                        const uint32_t no_w = record_no(j, w);
                        if (no_w != no) { r = ldg_record(P.pairs + 8 * (size_t)no_w); no = no_w; }
                        record_step(r, e[w], conc, acc[w]);
                    }
                };
                // two records in flight per lane, ping-pong (no register copies)
                uint32_t no_a = record_no(0, 0), no_b;
                PairRecord ra = ldg_record(P.pairs + 8 * (size_t)no_a), rb;
                for (int j = 0; j < n_steps; j += 2) {
                    no_b = record_no(j + 1, 0);
                    rb = ldg_record(P.pairs + 8 * (size_t)no_b);
                    lane_step(ra, no_a, j);
                    if (j + 2 < n_steps) {
                        no_a = record_no(j + 2, 0);
                        ra = ldg_record(P.pairs + 8 * (size_t)no_a);
                    }
                    lane_step(rb, no_b, j + 1);
                }
            }
        }

        finish_lane_lookups<kPerLane>(A, sink, t0, on, acc, my_sum);
    }
    warp_groups_done(sink, lane);
    finish_launch(T, sink, my_sum, s_part);
}

// (xs_dense_kernel, the kernel for dense segments, lives in xs_dense.cuh)

// Per-nuclide bucket tables for nuclide-grid mode (init only): bucket[i][b] = number of grid
// points of nuclide i whose energy maps to a bucket < b (same monotone map as the query).
__global__ void xs_build_nuclide_buckets_kernel(const double2 *grid, long n_iso, long n_gp, int n_buckets, uint32_t *bucket)
{
    const long total = n_iso * (n_gp + 1);
    const long stride = (long)gridDim.x * blockDim.x;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const long i = t / (n_gp + 1), k = t - i * (n_gp + 1);
        const double2 *g = grid + 3 * i * n_gp;
        const int b_prev = (k == 0) ? -1 : bucket_of(g[3 * (k - 1)].x, (double)n_buckets, n_buckets);
        const int b_here = (k == n_gp) ? n_buckets : bucket_of(g[3 * k].x, (double)n_buckets, n_buckets);
        uint32_t *row = bucket + i * (n_buckets + 1);
        for (int b = b_prev + 1; b <= b_here; b++) row[b] = (uint32_t)k;
    }
}

// Pair records for the sweep kernels (init only); record r = nuc*n_gp + k, k <= n_gp-2.  The
// last slot of every nuclide (k = n_gp-1) repeats k = n_gp-2: the reference clamps an index that
// points at the last grid point to the last interval (cuda/Simulation.cu:120-133, 159-162), and a gather
// that reads the slot directly gets that clamp for free.
__global__ void xs_build_pairs_kernel(const double2 *grid, long n_iso, long n_gp, double2 *pairs)
{
    const long total = n_iso * n_gp;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long r = (long)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += stride) {
        double2 out[8];
#pragma unroll
        for (int i = 0; i < 8; i++) out[i] = make_double2(0.0, 0.0);
        if (n_gp >= 2) {
            const long lo = (r % n_gp + 1 < n_gp) ? r : r - 1;
            const double2 l0 = grid[3 * lo], l1 = grid[3 * lo + 1], l2 = grid[3 * lo + 2];
            const double2 h0 = grid[3 * lo + 3], h1 = grid[3 * lo + 4], h2 = grid[3 * lo + 5];
            const double d = h0.x - l0.x;
            out[0] = make_double2(h0.y, h0.y - l0.y);        // total
            out[1] = make_double2(h1.x, h1.x - l1.x);        // elastic
            out[2] = make_double2(h1.y, h1.y - l1.y);        // absorbtion
            out[3] = make_double2(h2.x, h2.x - l2.x);        // fission
            out[4] = make_double2(h2.y, h2.y - l2.y);        // nu_fission
            out[5] = make_double2(l0.x, 0.0);                // lo.E (probed by the record searches)
            out[6] = make_double2(h0.x, d);                  // hi.E, d
            out[7] = make_double2(1.0 / d, 0.0);             // correctly rounded reciprocal
        }
#pragma unroll
        for (int i = 0; i < 8; i++) pairs[8 * r + i] = out[i];
    }
}

// ---------------------------------------------------------------------------------------
// Sampling kernel (variants 1..6): writes energy, material, optionally the UEG row / hash bin
// ("where"), a sort key, and counts lookups per material (replaces 12 x thrust::count).
// key = (material << 28) | top 28 bits of the 63-bit LCG state behind the energy (monotone in
// the energy).
// ---------------------------------------------------------------------------------------
#ifndef XS_SAMPLE_BLOCKS
#define XS_SAMPLE_BLOCKS 8            // 32 registers, 64 warps per SM: the kernel waits on dependent L2 round trips (0.23 -> 0.21 ms per 17 M against 37 registers / 48 warps)
#endif
__global__ void __launch_bounds__(256, XS_SAMPLE_BLOCKS)
xs_sample_kernel(const Problem P, int grid_type, long first_id, long count, double *energy, int *mat,
                 uint32_t *where, uint32_t *key, unsigned int *mat_histogram, unsigned int *bin_count, int bin_shift,
                 uint32_t row_begin, uint32_t row_end, double2 *pack, const DigitSpec digits)
{
    __shared__ unsigned int s_hist[kNumMaterials];
    __shared__ unsigned int s_digits[kMaxSortPasses][kRadix];   // the sort's digit counts, taken while the key is in a register
    if (threadIdx.x < kNumMaterials) s_hist[threadIdx.x] = 0;
    if (digits.digit_hist) digit_count_zero(s_digits);
    // the sort's top digit is (material << 4 | 4 energy bits): its counts add up to the material histogram
    // (one shared-memory atomic per lookup less -- the one with 12 hot addresses)
    const bool hist_from_digits = digit_top_is_material(digits);
    __syncthreads();
    const long stride = (long)gridDim.x * blockDim.x;
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < count) {
        // one full skip-ahead, then a fixed-stride affine jump per further element
        uint64_t s = lcg_skip(kStartSeed, 2ULL * (uint64_t)(first_id + t));
        const Affine hop = lcg_jump(2ULL * (uint64_t)stride);
        for (; t < count; t += stride) {
            const uint64_t s1 = lcg_step(s), s2 = lcg_step(s1);
            const int m = pick_material(P, lcg_to_double(s2));
            const double e = lcg_to_double(s1);
            if (energy) energy[t] = e;
            if (mat) mat[t] = m;
            bool mine = true;
            if (where || pack) {
                // energy-band sharding (unionized grid too large for one GPU): every device draws
                // every lookup and keeps those whose row lies in its band [row_begin, row_end);
                // the others get material 15 in the key (sorted to the end, never looked up)
                const uint32_t w = (uint32_t)locate_rt(P, grid_type, e);
                if (where) where[t] = w;
                if (pack) pack[t] = make_double2(e, __longlong_as_double((long long)w));
                mine = w >= row_begin && w < row_end;
            }
            const uint32_t k32 = ((mine ? (uint32_t)m : 15u) << 28) | (uint32_t)(s1 >> 35);
            if (key) key[t] = k32;
            if (digits.digit_hist) digit_count_key(s_digits, digits, k32);
            if (bin_count) atomicAdd(bin_count + (k32 >> bin_shift), 1u);
            if (mat_histogram && mine && !hist_from_digits) atomicAdd(&s_hist[m], 1u);
            s = apply(hop, s);
        }
    }
    __syncthreads();
    if (hist_from_digits && threadIdx.x < kNumMaterials) s_hist[threadIdx.x] = digit_material_count(s_digits, digits, threadIdx.x);
    if (mat_histogram && threadIdx.x < kNumMaterials && s_hist[threadIdx.x])
        atomicAdd(mat_histogram + threadIdx.x, s_hist[threadIdx.x]);
    if (digits.digit_hist) digit_count_flush(s_digits, digits);
}

// History mode, one generation for all particles (openmp-threading/Simulation.c:163-171, 225-233):
// generation 0 seeds particle p at fast_forward_LCG(1070, p*lookups*2*5); later generations
// first skip n_forward states (the feedback of the previous lookup), then draw energy and
// material.  Output goes to the same sample arrays the event-mode sampler fills, so the
// regrouping + windowed sweep that follow are shared with event mode.
__global__ void __launch_bounds__(256)
xs_history_step_kernel(const Problem P, int grid_type, long first_particle, long n_particles, int lookups,
                       int generation, uint64_t *seeds, const unsigned char *fwd, double *energy, int *mat,
                       uint32_t *where, unsigned int *mat_histogram, uint32_t *key, double2 *pack,
                       uint32_t row_begin, uint32_t row_end)
{
    __shared__ unsigned int s_hist[kNumMaterials];
    if (threadIdx.x < kNumMaterials) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const long stride = (long)gridDim.x * blockDim.x;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n_particles; t += stride) {
        uint64_t s;
        if (generation == 0) {
            s = lcg_skip(kStartSeed, (uint64_t)(first_particle + t) * (uint64_t)lookups * 10ULL);
        } else {
            s = seeds[t];
            for (int k = fwd[t]; k > 0; k--) s = lcg_step(s);
        }
        s = lcg_step(s);
        const double e = lcg_to_double(s);
        s = lcg_step(s);
        const int m = pick_material(P, lcg_to_double(s));
        seeds[t] = s;
        energy[t] = e;
        mat[t] = m;
        const uint32_t w = (uint32_t)locate_rt(P, grid_type, e);
        where[t] = w;
        const bool mine = w >= row_begin && w < row_end;      // (energy-band sharding, like the event sampler)
        if (pack) pack[t] = make_double2(e, __longlong_as_double((long long)w));
        // the sorted pipeline's key: material, then 28 bits monotone in the energy (like xs_locate_kernel)
        if (key) key[t] = ((mine ? (uint32_t)m : 15u) << 28) | (uint32_t)fmin(e * 268435456.0, 268435455.0);
        if (mine) atomicAdd(&s_hist[m], 1u);
    }
    __syncthreads();
    if (threadIdx.x < kNumMaterials && s_hist[threadIdx.x])
        atomicAdd(mat_histogram + threadIdx.x, s_hist[threadIdx.x]);
}

// Same bookkeeping for samples that already exist (host-provided): where + histogram.  The samples
// come from outside the library, so they are validated here: a material outside [0, 12) or an energy
// outside [0, 1] (NaN included) would index shared-memory histograms, material tables and the
// hash / unionized grids out of bounds further down.  Such a sample is counted in *bad (the call
// then fails with XS_ERR_ARG) and replaced in place by a harmless one, so nothing downstream can fault.
XS_DEV bool sanitize_sample(double &e, int &m)
{
    const bool ok = (unsigned)m < (unsigned)kNumMaterials && e >= 0.0 && e <= 1.0;
    if (!ok) { e = 0.5; m = 0; }
    return ok;
}

__global__ void __launch_bounds__(256)
xs_locate_kernel(const Problem P, int grid_type, long count, double *energy, int *mat, uint8_t *mat8,
                 uint32_t *where, uint32_t *key, unsigned int *mat_histogram, double2 *pack, unsigned long long *bad,
                 uint32_t row_begin, uint32_t row_end, const DigitSpec digits)
{
    // materials: ints as the caller holds them, or (mat8) the bytes the host narrowed them to (xs_hostpack.h)
    __shared__ unsigned int s_hist[kNumMaterials];
    __shared__ unsigned int s_digits[kMaxSortPasses][kRadix];
    if (threadIdx.x < kNumMaterials) s_hist[threadIdx.x] = 0;
    if (digits.digit_hist) digit_count_zero(s_digits);
    const bool hist_from_digits = key && digit_top_is_material(digits);      // (see xs_sample_kernel)
    __syncthreads();
    const long stride = (long)gridDim.x * blockDim.x;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        double e = energy[t];
        int m = mat8 ? (int)mat8[t] : mat[t];
        if (!sanitize_sample(e, m)) {
            atomicAdd(bad, 1ULL);
            energy[t] = e;
            if (mat8) mat8[t] = (uint8_t)m; else mat[t] = m;
        }
        const uint32_t w = (uint32_t)locate_rt(P, grid_type, e);
        if (where) where[t] = w;
        if (pack) pack[t] = make_double2(e, __longlong_as_double((long long)w));
        // energy-band sharding: a lookup whose row another device holds gets material 15 in the key (sorted
        // behind the last segment, never looked up) and is not counted
        const bool mine = w >= row_begin && w < row_end;
        if (key) {    // same layout as the sampler's key: material, then 28 bits monotone in the energy
            const double scaled = fmin(e * 268435456.0, 268435455.0);
            const uint32_t k32 = ((mine ? (uint32_t)m : 15u) << 28) | (uint32_t)scaled;
            key[t] = k32;
            if (digits.digit_hist) digit_count_key(s_digits, digits, k32);
        }
        if (mine && !hist_from_digits) atomicAdd(&s_hist[m], 1u);
    }
    __syncthreads();
    if (hist_from_digits && threadIdx.x < kNumMaterials) s_hist[threadIdx.x] = digit_material_count(s_digits, digits, threadIdx.x);
    if (threadIdx.x < kNumMaterials && s_hist[threadIdx.x])
        atomicAdd(mat_histogram + threadIdx.x, s_hist[threadIdx.x]);
    if (digits.digit_hist) digit_count_flush(s_digits, digits);
}

// Validation only (the in-order kernel reads host samples directly: XSB200_SWEEP=0).
__global__ void __launch_bounds__(256)
xs_validate_samples_kernel(long count, double *energy, int *mat, unsigned long long *bad)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        double e = energy[t];
        int m = mat[t];
        if (!sanitize_sample(e, m)) {
            atomicAdd(bad, 1ULL);
            energy[t] = e;
            mat[t] = m;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Partition by material (-k 4) or fuel / not-fuel (-k 5): replaces thrust::count x12 +
// thrust::sort_by_key (cuda/Simulation.cu:792-797) and thrust::partition (:936).  The group
// sizes are already known (histogram from the sampler), so one pass suffices: a block counts
// its tile per group in shared memory, reserves a range per group with one atomicAdd on the
// group cursor, and scatters energy / where / material / sample id.  Order inside a group is
// arbitrary (the checksum is order independent; per-lookup outputs are addressed by id).
// ---------------------------------------------------------------------------------------
constexpr int kPartItems = 8;
__global__ void __launch_bounds__(256)
xs_partition_kernel(const double *energy, const int *mat, const uint32_t *where, long count,
                    const unsigned int *mat_histogram, unsigned int *cursor, int fuel_or_not,
                    double *out_energy, uint32_t *out_where, int *out_mat, uint32_t *out_id)
{
    __shared__ unsigned int s_cnt[kNumMaterials], s_base[kNumMaterials];
    if (threadIdx.x < kNumMaterials) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const long base = (long)blockIdx.x * (256 * kPartItems);
    int grp[kPartItems];
    unsigned int rank[kPartItems];
#pragma unroll
    for (int i = 0; i < kPartItems; i++) {
        const long t = base + i * 256 + threadIdx.x;
        grp[i] = -1;
        if (t < count) {
            const int m = mat[t];
            grp[i] = fuel_or_not ? (m != 0) : m;
            rank[i] = atomicAdd(&s_cnt[grp[i]], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < kNumMaterials && s_cnt[threadIdx.x]) {
        // start of this group = sizes of the groups before it
        unsigned int start = 0;
        if (fuel_or_not) { if (threadIdx.x == 1) start = mat_histogram[0]; }
        else for (int m = 0; m < (int)threadIdx.x; m++) start += mat_histogram[m];
        s_base[threadIdx.x] = start + atomicAdd(cursor + threadIdx.x, s_cnt[threadIdx.x]);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kPartItems; i++) {
        const long t = base + i * 256 + threadIdx.x;
        if (grp[i] >= 0) {
            const unsigned int pos = s_base[grp[i]] + rank[i];
            out_energy[pos] = energy[t];
            out_where[pos] = where[t];
            if (out_mat) out_mat[pos] = mat[t];
            out_id[pos] = (uint32_t)t;
        }
    }
}

// Bin sort of -k 6 (default): the key (material, energy) is uniformly distributed inside a
// material, so ONE counting pass over fine value bins (material x 2^16 energy bins, counted by the
// sampler, exclusive-scanned in place) orders the lookups as well as the lane-per-lookup kernel
// needs -- 36 fuel lookups per bin at "large", a sixth of a grid interval.  Every lookup takes
// the next free slot of its bin (atomicAdd on the scanned table) and carries its payload along,
// so the three radix passes and the random-access gather behind them (1.6 ms) become one pass.
// The order inside a bin depends on atomic arrival order; results do not (every lookup is
// independent, the verification is a sum).
__global__ void __launch_bounds__(256)
xs_bin_scatter_kernel(const uint32_t *key, const double *energy, const uint32_t *where, long count,
                      unsigned int *bin_cursor, int bin_shift, double *out_energy, uint32_t *out_where,
                      uint32_t *out_id)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        const unsigned int at = atomicAdd(bin_cursor + (key[t] >> bin_shift), 1u);
        out_energy[at] = energy[t];
        out_where[at] = where[t];
        out_id[at] = (uint32_t)t;
    }
}

// Apply a permutation (from the radix sort, XSB200_BIN_BITS=0): grouped copies of energy / where.
__global__ void __launch_bounds__(256)
xs_gather_kernel(const uint32_t *perm, const double *energy, const uint32_t *where, long count,
                 double *out_energy, uint32_t *out_where)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        const uint32_t at = perm[t];
        out_energy[t] = energy[at];
        out_where[t] = where[at];
    }
}

// ---------------------------------------------------------------------------------------
// History mode: one warp per particle, `lookups` dependent lookups each
// (openmp-threading/Simulation.c:116-238).
// ---------------------------------------------------------------------------------------
template <int GRID, int GATHER>
__global__ void __launch_bounds__(kBlockThreads, 2)
xs_history_kernel(const Problem P, long first_particle, long n_particles, int lookups,
                  const BatchSink sink)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SharedTables &T = *reinterpret_cast<SharedTables *>(smem_raw);
    double *s_conc = reinterpret_cast<double *>(smem_raw + sizeof(SharedTables));
    int *s_nuc = reinterpret_cast<int *>(s_conc + P.mat_total);
    __shared__ unsigned long long s_part[kWarpsPerBlock];
    stage_tables(P, T, s_nuc, s_conc);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *exchange = T.exchange[warp];
    unsigned long long my_sum = 0, my_count = 0;     // kept on lane 0

    for (;;) {
        long t = 0;
        if (lane == 0) t = atomicAdd(sink.batch_counter, 1u);
        t = __shfl_sync(kFullMask, t, 0);
        if (t >= n_particles) break;
        const uint64_t p = (uint64_t)(first_particle + t);
        // all lanes carry the same state (warp-uniform control flow)
        uint64_t s = lcg_skip(kStartSeed, p * (uint64_t)lookups * 10ULL);
        s = lcg_step(s); double e = lcg_to_double(s);
        s = lcg_step(s); int mat = pick_material(P, lcg_to_double(s));
        for (int i = 0; i < lookups; i++) {
            const long where = locate<GRID>(P, e);
            const int first = T.first[mat], n = T.first[mat + 1] - first;
            double out[5];
            warp_macro<GRID, GATHER>(P, s_nuc, s_conc, first, n, e, where, lane, exchange, T.stage[warp], out);
            double gap;
            int am = argmax5(out, gap);
            bool redo = gap <= kTieGuard;
#pragma unroll
            for (int k = 0; k < 5; k++) redo |= fabs(out[k] - 1.0) <= kTieGuard;
            if (redo) {                          // warp-uniform (all lanes hold equal sums)
                macro_xs_serial<GRID>(P, e, mat, out);
                am = argmax5(out, gap);
            }
            int fwd = 0;
#pragma unroll
            for (int k = 0; k < 5; k++) fwd += out[k] > 1.0;
            my_sum += (unsigned long long)(am + 1);
            my_count += 1;
            for (int k = 0; k < fwd; k++) s = lcg_step(s);
            s = lcg_step(s); e = lcg_to_double(s);
            s = lcg_step(s); mat = pick_material(P, lcg_to_double(s));
        }
    }
    if (lane != 0) { my_sum = 0; my_count = 0; }
    const unsigned long long bs = block_sum(my_sum, s_part);
    const unsigned long long bc = block_sum(my_count, s_part);
    if (threadIdx.x == 0 && bc) {
        atomicAdd(sink.accum, bs);
        atomicAdd(sink.accum + 1, bc);
    }
}

// ---------------------------------------------------------------------------------------
// Self-test of the Newton-Markstein quotient (xs_gpu_selftest_division): the kernels' f against IEEE division.
// ---------------------------------------------------------------------------------------
XS_DEV uint64_t mix64(uint64_t z)          // splitmix64 finaliser: a counter-based generator
{
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
xs_division_selftest_kernel(unsigned long long seed, long n_pairs, int mode, unsigned long long *mismatches)
{
    const long stride = (long)gridDim.x * blockDim.x;
    unsigned int bad = 0;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n_pairs; t += stride) {
        const uint64_t a = mix64(seed + 3ULL * (uint64_t)t), b = mix64(seed + 3ULL * (uint64_t)t + 1), c = mix64(seed + 3ULL * (uint64_t)t + 2);
        double n, d;
        if (mode == 0) {
            // as a lookup forms them: grid energies lo < hi and lo < E <= hi, all in [0, 1)
            const double u1 = (double)(a >> 11) * 0x1p-53, u2 = (double)(b >> 11) * 0x1p-53, u3 = (double)(c >> 11) * 0x1p-53;
            const double lo = fmin(u1, u2), hi = fmax(u1, u2);
            double e = lo + u3 * (hi - lo);
            e = fmin(fmax(e, lo), hi);
            n = hi - e;
            d = hi - lo;
        } else {
            const int ed = -(int)(b >> 58);                                    // exponent of d in [-63, 0]
            uint64_t md = a & 0x000fffffffffffffULL;
            if (mode == 2) {
                const int k = (int)((c >> 52) & 63) % 52;
                switch ((int)(c >> 60) & 3) {
                    case 0: md = 0x000fffffffffffffULL >> (k % 8); break;                     // (nearly) all ones
                    case 1: md = 1ULL << k; break;                                              // a single bit
                    case 2: md = 0x000fffffffffffffULL & ~(1ULL << k); break;                   // all ones but one
                    default: md = (a & 0xffULL) << (k % 44); break;                            // a short pattern somewhere
                }
            }
            d = __longlong_as_double((long long)(((uint64_t)(1023 + ed) << 52) | md));
            n = d * ((double)(c >> 11) * 0x1p-53);                                              // 0 <= n <= d
            if (mode == 2 && (c & 1)) n = __longlong_as_double(__double_as_longlong(n) | (long long)(b & 0x7ULL));
        }
        if (!(d > 0.0) || !(n >= 0.0) || n > d) continue;
        const double inv = 1.0 / d;                                                             // IEEE: what xs_build_pairs_kernel stores
        const double q = n * inv;
        const double rem = __fma_rn(-d, q, n);
        const double f = __fma_rn(rem, inv, q);
        if (__double_as_longlong(f) != __double_as_longlong(n / d)) bad++;
    }
    if (bad) atomicAdd(mismatches, (unsigned long long)bad);
}

// ---------------------------------------------------------------------------------------
// Bucket table over the unionized grid (built once at init).
//   bucket[b] = number of rows whose energy maps to a bucket < b
// ---------------------------------------------------------------------------------------
__global__ void xs_build_buckets_kernel(const double *ueg, long n_ueg, double scale, int n_buckets,
                                        uint32_t *bucket)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long r = (long)blockIdx.x * blockDim.x + threadIdx.x; r <= n_ueg; r += stride) {
        const int b_prev = (r == 0) ? -1 : bucket_of(ueg[r - 1], scale, n_buckets);
        const int b_here = (r == n_ueg) ? n_buckets : bucket_of(ueg[r], scale, n_buckets);
        for (int b = b_prev + 1; b <= b_here; b++) bucket[b] = (uint32_t)r;
    }
}

// entry[b] = first[b] << 4 | min(first[b + 1] - first[b], 15)   (see ueg_row_t)
__global__ void xs_pack_buckets_kernel(const uint32_t *first, int n_buckets, uint32_t *entry)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long b = (long)blockIdx.x * blockDim.x + threadIdx.x; b <= n_buckets; b += stride) {
        const uint32_t f = first[b];
        const uint32_t rows = b < n_buckets ? first[b + 1] - f : 0u;
        entry[b] = (f << 4) | (rows < 15u ? rows : 15u);
    }
}

}  // namespace xs
