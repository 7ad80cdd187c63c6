// xs_kernels.cuh -- the lookup kernels.
//
// One kernel body serves every event-mode variant (-k 0..6) and the host-sample entry
// point; what changes between variants is where a warp's batch of 32 lookups comes from
// (sampled in-kernel from the lookup id, or read from sample arrays, optionally through a
// permutation / material filter) -- see xs_gpu.cu for the dispatch.
//
// Work decomposition (B200: 148 SMs x 64 resident warps):
//   * persistent grid, warps fetch BATCHES of 32 lookups from a global atomic counter;
//   * inside a batch the per-lookup scalar work is lane-parallel: lane l draws the sample of
//     lookup l (LCG skip-ahead) and finds its unionized-grid row / hash bin;
//   * the gather is warp-cooperative: the warp walks the batch one lookup at a time and
//     spreads that lookup's nuclides over its lanes (two lane mappings, see Gather below);
//   * per-channel sums are combined with warp shuffles; lane l keeps lookup l's argmax;
//   * the verification sum is reduced warp -> block -> one atomicAdd(u64) per block, so the
//     reference's 136 MB verification[] array and its thrust::reduce pass do not exist.
#pragma once

#include "xs_device.cuh"

namespace xs {

constexpr int kBlockThreads = 256;
constexpr int kWarpsPerBlock = kBlockThreads / 32;
constexpr double kTieGuard = 1e-10;      // relative gap below which order of summation could
                                         // change an integer decision -> settle serially

// Lane mappings of the gather.
//   kLanePerNuclide : lane j handles nuclide j (+32, ...): 6 divergent 16-B loads per lane.
//   kTriple         : 3 adjacent lanes share one nuclide; lane c loads chunk c of the low and
//                     of the high grid point (2 loads per lane, 48 contiguous bytes per
//                     instruction and nuclide -> ~2.5x fewer L1 wavefronts per nuclide).
constexpr int kLanePerNuclide = 0, kTriple = 1;

// Where a batch's (energy, material) pairs come from.
struct BatchSource {
    const double *energy;      // sample arrays (nullptr => sample in-kernel from lookup ids)
    const int    *mat;
    const uint32_t *perm;      // optional permutation: slot t reads sample perm[t]
    long   first_id;           // in-kernel sampling: id of slot 0
    long   count;              // number of slots
    int    mat_lo, mat_hi;     // only lookups with mat in [mat_lo, mat_hi] are performed
};

struct BatchSink {
    unsigned long long *accum; // [0] += sum(argmax+1), [1] += lookups performed
    unsigned int *batch_counter;
    double *macro_out;         // optional [count*5]
    double *energy_out;        // optional [count]
    int    *mat_out;           // optional [count]
    int    *argmax_out;        // optional [count]
};

// Material tables staged in shared memory (compact CSR: 484 entries at "large").
struct SharedTables {
    int    first[kNumMaterials + 1];
    int    pad[3];
    double exchange[kWarpsPerBlock][8];   // per-warp scratch for the 3-lane reduction
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

template <int GRID>
XS_DEV void warp_macro_triple(const Problem &P, const int *s_nuc, const double *s_conc,
                              int first, int n, double e, long where, int lane,
                              double *exchange /* 8 doubles, per warp */, double out[5])
{
    const int slot = lane / 3;                 // 10 nuclides per step; lanes 30,31 idle
    const int chunk = lane - 3 * slot;         // which 16-byte chunk of a grid point
    const bool lane_on = lane < 30;
    const int f_src = lane - chunk;            // lane holding both energies of this slot
    double acc_x = 0.0, acc_y = 0.0;

#pragma unroll 2
    for (int j0 = 0; j0 < n; j0 += 10) {
        const int j = j0 + slot;
        const bool on = lane_on && j < n;
        double2 lo = make_double2(0.0, 0.0), hi = make_double2(1.0, 0.0);
        double conc = 0.0;
        if (on) {
            const int nuc = s_nuc[first + j];
            conc = s_conc[first + j];
            const int low = nuclide_low<GRID>(P, e, where, nuc);
            const double2 *p = P.grid + 3 * ((long)nuc * P.n_gp + low) + chunk;
            lo = ldg_grid(p);
            hi = ldg_grid(p + 3);
        }
        const double f_own = (hi.x - e) / (hi.x - lo.x);      // meaningful on chunk 0 only
        const double f = __shfl_sync(kFullMask, f_own, f_src);
        if (on) {
            acc_x += lerp_xs(lo.x, hi.x, f) * conc;
            acc_y += lerp_xs(lo.y, hi.y, f) * conc;
        }
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
                       double e, long where, int lane, double *exchange, double out[5])
{
    if (GATHER == kTriple) warp_macro_triple<GRID>(P, s_nuc, s_conc, first, n, e, where, lane, exchange, out);
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
                s = lcg_step(s); mat_l = pick_material(lcg_to_double(s));
            }
        }
        const bool want = have && mat_l >= src.mat_lo && mat_l <= src.mat_hi;
        long where_l = want ? locate<GRID>(P, e_l) : 0;
        unsigned todo = __ballot_sync(kFullMask, want);

        // ---- warp-cooperative part: one lookup of the batch at a time --------------------
        int my_argmax = 0;
        double mine[5] = {0, 0, 0, 0, 0};
        while (todo) {
            const int i = __ffs(todo) - 1;
            todo &= todo - 1;
            const double e = __shfl_sync(kFullMask, e_l, i);
            const int mat = __shfl_sync(kFullMask, mat_l, i);
            const long where = __shfl_sync(kFullMask, where_l, i);
            const int first = T.first[mat], n = T.first[mat + 1] - first;
            double out[5];
            warp_macro<GRID, GATHER>(P, s_nuc, s_conc, first, n, e, where, lane, exchange, out);
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
// Sampling kernel (variants 1..6): writes energy, material and -- for the sorted variants
// -- a sort key, and counts lookups per material.  key = (material << 28) | top 28 bits of
// the 63-bit LCG state behind the energy (monotone in the energy).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
xs_sample_kernel(long first_id, long count, double *energy, int *mat, uint32_t *key,
                 unsigned int *mat_histogram)
{
    __shared__ unsigned int s_hist[kNumMaterials];
    if (threadIdx.x < kNumMaterials) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const long stride = (long)gridDim.x * blockDim.x;
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < count) {
        // one full skip-ahead, then a fixed-stride affine jump per further element
        uint64_t s = lcg_skip(kStartSeed, 2ULL * (uint64_t)(first_id + t));
        const Affine hop = lcg_jump(2ULL * (uint64_t)stride);
        for (; t < count; t += stride) {
            const uint64_t s1 = lcg_step(s), s2 = lcg_step(s1);
            const int m = pick_material(lcg_to_double(s2));
            energy[t] = lcg_to_double(s1);
            mat[t] = m;
            if (key) key[t] = ((uint32_t)m << 28) | (uint32_t)(s1 >> 35);
            if (mat_histogram) atomicAdd(&s_hist[m], 1u);
            s = apply(hop, s);
        }
    }
    __syncthreads();
    if (mat_histogram && threadIdx.x < kNumMaterials && s_hist[threadIdx.x])
        atomicAdd(mat_histogram + threadIdx.x, s_hist[threadIdx.x]);
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
        s = lcg_step(s); int mat = pick_material(lcg_to_double(s));
        for (int i = 0; i < lookups; i++) {
            const long where = locate<GRID>(P, e);
            const int first = T.first[mat], n = T.first[mat + 1] - first;
            double out[5];
            warp_macro<GRID, GATHER>(P, s_nuc, s_conc, first, n, e, where, lane, exchange, out);
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
            s = lcg_step(s); mat = pick_material(lcg_to_double(s));
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

}  // namespace xs
