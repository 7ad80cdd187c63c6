// xs_generate.cuh -- device-side construction of the synthetic problem.
//
// Produces the same arrays as the host generator (xsbench_b200/host/gridinit.c, i.e. the
// reference's grid_init_do_not_profile, cuda/GridInit.cu:90-262) directly in HBM, so that the
// 5.7 GB (large) / 116 GB (XL) problem never has to exist in host memory or cross PCIe:
//
//   1. keys   s1[p] = the 63-bit LCG state behind the energy of raw point p.  The reference
//             draws 6 values per point from ONE stream seeded 42 (:99,123-131): point p owns
//             draws 6p+1..6p+6, reachable by skip-ahead, so every point is independent.
//   2. sort A radix sort of all keys (8 passes of 8 bits, payload = p).  The state is
//             monotone in the energy, so the sorted order IS the unionized energy grid
//             (:169-173: sorted copy of all energies).
//   3. sort B stable 2-pass sort of that order by nuclide id (p / n_gp): now every nuclide's
//             points are contiguous and ascending in energy (:134-135, qsort per nuclide).
//             The six values of a point are then re-drawn from p into their final place.
//   4. index grid: the reference's sweep (:187-206) run per (row tile, nuclide) with cursors
//             started from the closed form, 32 nuclides per warp so that rows
//             are written in coalesced 128-byte segments.
//   5. hash grid: one bounded search per (bin, nuclide) (:221-233).
//
// Equal to the host generator byte for byte unless one nuclide holds two exactly equal
// energies (probability ~7e-9 per nuclide: qsort's tie order and the sweep's lag are then
// implementation details); that case is detected and reported instead of guessed.
#pragma once

#include "xs_device.cuh"
#include "xs_sort.cuh"

namespace xs {

constexpr uint64_t kGridSeed = 42ULL;

// ---- 1. keys ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gen_keys_kernel(long n_points, uint64_t *keys)
{
    const long stride = (long)gridDim.x * blockDim.x;
    long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_points) return;
    uint64_t s = lcg_skip(kGridSeed, 6ULL * (uint64_t)p);         // state before point p's draws
    const Affine hop = lcg_jump(6ULL * (uint64_t)stride);
    for (; p < n_points; p += stride) {
        keys[p] = lcg_step(s);                                     // energy draw of point p
        s = apply(hop, s);
    }
}

// ---- 2. unionized energy grid = keys in sorted order -------------------------------------
__global__ void __launch_bounds__(256)
gen_ueg_kernel(const uint64_t *sorted_keys, long n, double *ueg)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        ueg[i] = lcg_to_double(sorted_keys[i]);
}

// nuclide id of every entry of the energy-sorted order (key of sort B)
__global__ void __launch_bounds__(256)
gen_nuclide_key_kernel(const uint32_t *perm, long n, uint32_t n_gp, uint32_t *nuc_key)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        nuc_key[i] = perm[i] / n_gp;
}

// ---- 3. final nuclide grid: slot q <- the six draws of raw point perm[q] ----------------------
__global__ void __launch_bounds__(256)
gen_points_kernel(const uint32_t *perm, long n_points, double2 *grid, int *duplicate_flag)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < n_points; q += stride) {
        uint64_t s = lcg_skip(kGridSeed, 6ULL * (uint64_t)perm[q]);
        double v[6];
#pragma unroll
        for (int k = 0; k < 6; k++) { s = lcg_step(s); v[k] = lcg_to_double(s); }
        grid[3 * q] = make_double2(v[0], v[1]);
        grid[3 * q + 1] = make_double2(v[2], v[3]);
        grid[3 * q + 2] = make_double2(v[4], v[5]);
    }
    (void)duplicate_flag;
}

__global__ void __launch_bounds__(256)
gen_check_duplicates_kernel(const double2 *grid, long n_iso, long n_gp, int *duplicate_flag)
{
    const long n = n_iso * n_gp, stride = (long)gridDim.x * blockDim.x;
    for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q + 1 < n; q += stride)
        if ((q + 1) % n_gp != 0 && grid[3 * q].x == grid[3 * (q + 1)].x) *duplicate_flag = 1;
}

// ---- 4. index grid ----------------------------------------------------------------------
// Block = one tile of kIndexTileRows rows; warp w handles nuclides [32w', 32w'+32) in turn.
constexpr int kIndexTileRows = 512;
__global__ void __launch_bounds__(256)
gen_index_kernel(const double *ueg, const double2 *grid, long n_iso, long n_gp, long row_begin, long row_end,
                 int *index_grid)
{
    // rows [row_begin, row_end) of the unionized grid (an energy band; the whole grid by default);
    // index_grid holds exactly those rows
    const long e_begin = row_begin + (long)blockIdx.x * kIndexTileRows;
    const long e_end = min(e_begin + (long)kIndexTileRows, row_end);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (long i0 = 32L * warp; i0 < n_iso; i0 += 32L * n_warps) {
        const long i = i0 + lane;
        const bool on = i < n_iso;
        const double2 *g = grid + 3 * (on ? i : 0) * n_gp;
        // cursor valid for row e_begin-1: min(n_gp-2, #{k >= 1 : g[k].energy <= ueg[e_begin-1]})
        int cursor = 0;
        if (on && e_begin > 0) {
            const double q = ueg[e_begin - 1];
            if (g[3].x <= q) {
                if (q >= g[3 * (n_gp - 1)].x) cursor = (int)n_gp - 1;
                else {
                    int lo = 1, hi = (int)n_gp - 1;
                    while (hi - lo > 1) {
                        const int mid = lo + (hi - lo) / 2;
                        if (g[3 * (long)mid].x > q) hi = mid; else lo = mid;
                    }
                    cursor = lo;
                }
                if (cursor > n_gp - 2) cursor = (int)n_gp - 2;
            }
        }
        double next_energy = on ? g[3 * (long)(cursor + 1)].x : 2.0;
        for (long e = e_begin; e < e_end; e++) {
            const double ue = ueg[e];                                  // warp-uniform, L1 hit
            if (on && ue >= next_energy && cursor != n_gp - 2) {
                cursor++;
                next_energy = g[3 * (long)(cursor + 1)].x;
            }
            if (on) index_grid[(e - row_begin) * n_iso + i] = cursor;    // 32 consecutive ints per warp
        }
    }
}

// ---- 5. hash grid -----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gen_hash_kernel(const double2 *grid, long n_iso, long n_gp, int hash_bins, int *index_grid)
{
    const long n = (long)hash_bins * n_iso, stride = (long)gridDim.x * blockDim.x;
    const double du = 1.0 / (double)hash_bins;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        const long bin = t / n_iso, i = t - bin * n_iso;
        const double energy = (double)bin * du;
        const double2 *g = grid + 3 * i * n_gp;
        int lo = 0, hi = (int)n_gp - 1;
        while (hi - lo > 1) {
            const int mid = lo + (hi - lo) / 2;
            if (g[3 * (long)mid].x > energy) hi = mid; else lo = mid;
        }
        index_grid[t] = lo;
    }
}

// Index rows [row_begin, row_end) of the unionized grid from the generated grid + ueg (phase 2 of the
// generator; separate so that the sort temporaries of phase 1 are gone before a large index band is
// allocated).
inline int generate_index_rows(const double *ueg, const double2 *grid, long n_iso, long n_gp, long row_begin, long row_end,
                               int *index_grid, cudaStream_t stream)
{
    const long rows = row_end - row_begin;
    if (rows <= 0) return 0;
    const int tiles = (int)((rows + kIndexTileRows - 1) / kIndexTileRows);
    gen_index_kernel<<<tiles, 256, 0, stream>>>(ueg, grid, n_iso, n_gp, row_begin, row_end, index_grid);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// Generate grid (+ ueg, hash grid) into pre-allocated device arrays; the unionized index grid is
// built by generate_index_rows afterwards (index_grid may be null for that grid type).
// Returns 0, -1 on CUDA/sort failure, -2 if a nuclide holds duplicate energies.
inline int generate_problem(int grid_type, long n_iso, long n_gp, int hash_bins, double2 *grid, double *ueg,
                            int *index_grid, int sm_count, cudaStream_t stream)
{
    const long n_points = n_iso * n_gp;
    if (n_points > 0xffffffffL) return -1;
    uint64_t *key[2] = {nullptr, nullptr};
    uint32_t *perm[2] = {nullptr, nullptr}, *nuc_key[2] = {nullptr, nullptr};
    int *d_flag = nullptr;
    SortScratch scratch{};
    int rc = 0;
    auto cleanup = [&]() {
        for (int i = 0; i < 2; i++) { cudaFree(key[i]); cudaFree(perm[i]); cudaFree(nuc_key[i]); }
        cudaFree(d_flag);
        sort_scratch_free(scratch);
    };
    for (int i = 0; i < 2 && rc == 0; i++) {
        if (cudaMalloc(&key[i], (size_t)n_points * sizeof(uint64_t)) != cudaSuccess) rc = -1;
        if (rc == 0 && cudaMalloc(&perm[i], (size_t)n_points * sizeof(uint32_t)) != cudaSuccess) rc = -1;
        if (rc == 0 && cudaMalloc(&nuc_key[i], (size_t)n_points * sizeof(uint32_t)) != cudaSuccess) rc = -1;
    }
    if (rc == 0 && cudaMalloc(&d_flag, sizeof(int)) != cudaSuccess) rc = -1;
    if (rc == 0 && sort_scratch_alloc(scratch, n_points) != 0) rc = -1;
    if (rc != 0) { cleanup(); return -1; }
    cudaMemsetAsync(d_flag, 0, sizeof(int), stream);

    const int blocks = (int)std::min<long>((n_points + 255) / 256, (long)sm_count * 32);
    gen_keys_kernel<<<blocks, 256, 0, stream>>>(n_points, key[0]);
    uint32_t *order = nullptr;
    uint64_t *sorted_keys = nullptr;
    rc = radix_sort<uint64_t>(scratch, key, perm, n_points, 0, 63, 0, false, stream, &order, &sorted_keys, nullptr);
    if (rc == 0 && grid_type == kUnionized) gen_ueg_kernel<<<blocks, 256, 0, stream>>>(sorted_keys, n_points, ueg);
    if (rc == 0) {
        // sort B: payload = the energy order, key = nuclide id; the payload buffers are reused,
        // so the order must sit in perm[0]
        if (order != perm[0]) cudaMemcpyAsync(perm[0], order, (size_t)n_points * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream);
        gen_nuclide_key_kernel<<<blocks, 256, 0, stream>>>(perm[0], n_points, (uint32_t)n_gp, nuc_key[0]);
        int bits = 1;
        while ((1L << bits) < n_iso) bits++;
        uint32_t *final_order = nullptr;
        rc = radix_sort<uint32_t>(scratch, nuc_key, perm, n_points, 0, bits, 0, true, stream, &final_order, nullptr, nullptr);
        if (rc == 0) {
            gen_points_kernel<<<blocks, 256, 0, stream>>>(final_order, n_points, grid, d_flag);
            gen_check_duplicates_kernel<<<blocks, 256, 0, stream>>>(grid, n_iso, n_gp, d_flag);
        }
    }
    if (rc == 0 && grid_type == kHash) {
        const long n = (long)hash_bins * n_iso;
        gen_hash_kernel<<<(int)std::min<long>((n + 255) / 256, (long)sm_count * 32), 256, 0, stream>>>(grid, n_iso, n_gp, hash_bins, index_grid);
    }
    int h_flag = 0;
    if (rc == 0 && (cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
                    cudaStreamSynchronize(stream) != cudaSuccess || cudaGetLastError() != cudaSuccess))
        rc = -1;
    cleanup();
    if (rc == 0 && h_flag) rc = -2;
    return rc;
}

}  // namespace xs
