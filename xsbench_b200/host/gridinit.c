/*
 * gridinit.c -- synthetic problem generator (host side, not timed).
 *
 * Produces arrays that are byte-identical to the reference generator
 * (cuda/GridInit.cu:90-262 == openmp-threading/GridInit.c:3-160) but is organised for
 * parallel construction, because at "large" the reference's serial sweep writes 1.42 G
 * index entries and dominates wall time:
 *
 *   nuclide grid   the reference draws 6 values per point from ONE sequential LCG stream
 *                  (seed 42).  Nuclide i therefore owns draws [6*n_gp*i, 6*n_gp*(i+1)), so
 *                  each nuclide is filled independently after an LCG skip-ahead, then
 *                  sorted by energy with libc qsort + the same comparator (tie behaviour
 *                  of qsort is libc-defined; using the same call keeps ties identical).
 *   unionized grid sorted multiset of all energies: chunk-sort + merge (any correct sort
 *                  gives the same array of doubles).
 *   index grid     index_grid[e][i] is a per-nuclide monotone counter over e.  A block of
 *                  rows starting at e0 can start from the closed form
 *                      cursor_i = min(n_gp-2, #{k>=1 : grid_i[k].energy <= UEG[e0-1]})
 *                  and then run the reference's sweep inside the block.  The closed form
 *                  equals the sweep unless one nuclide holds two exactly equal energies;
 *                  that case is detected and handled by the serial sweep.
 *   hash grid      one bounded search per (bin, nuclide) -- embarrassingly parallel.
 */
#include "xs_host.h"

#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int double_compare(const void *a, const void *b)
{
    const double x = *(const double *)a, y = *(const double *)b;
    if (x > y) return 1;
    return (x < y) ? -1 : 0;
}

int NGP_compare(const void *a, const void *b)
{
    const double x = ((const NuclideGridPoint *)a)->energy;
    const double y = ((const NuclideGridPoint *)b)->energy;
    if (x > y) return 1;
    return (x < y) ? -1 : 0;
}

double get_time(void)
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}

/* Same formula as the reference's estimate (cuda/XSutils.cu:47-66), in MiB. */
size_t estimate_mem_usage(Inputs in)
{
    const size_t points = (size_t)in.n_isotopes * (size_t)in.n_gridpoints;
    size_t bytes = points * sizeof(NuclideGridPoint);
    if (in.grid_type == XS_UNIONIZED)
        bytes += points * sizeof(double) + points * (size_t)in.n_isotopes * sizeof(int);
    else if (in.grid_type == XS_HASH)
        bytes += (size_t)in.hash_bins * (size_t)in.n_isotopes * sizeof(int);
    return (size_t)ceil((double)bytes / (1024.0 * 1024.0));
}

/* Largest k in [lo, hi] ... restated: index of the last point with energy <= q inside the
 * bracket, by bisection with the invariant A[lo].energy <= q < A[hi].energy
 * (cuda/Simulation.cu:264-284). */
static long bracket_search(const NuclideGridPoint *A, double q, long lo, long hi)
{
    while (hi - lo > 1) {
        const long mid = lo + (hi - lo) / 2;
        if (A[mid].energy > q) hi = mid;
        else                   lo = mid;
    }
    return lo;
}

/* ---- nuclide grid ------------------------------------------------------------------ */
static int fill_nuclide_grid(NuclideGridPoint *grid, long n_iso, long n_gp)
{
    int has_duplicate = 0;
    #pragma omp parallel for schedule(dynamic, 1) reduction(|:has_duplicate)
    for (long i = 0; i < n_iso; i++) {
        NuclideGridPoint *g = grid + i * n_gp;
        uint64_t seed = fast_forward_LCG(42ULL, 6ULL * (uint64_t)n_gp * (uint64_t)i);
        for (long k = 0; k < n_gp; k++) {
            g[k].energy        = LCG_random_double(&seed);
            g[k].total_xs      = LCG_random_double(&seed);
            g[k].elastic_xs    = LCG_random_double(&seed);
            g[k].absorbtion_xs = LCG_random_double(&seed);
            g[k].fission_xs    = LCG_random_double(&seed);
            g[k].nu_fission_xs = LCG_random_double(&seed);
        }
        qsort(g, (size_t)n_gp, sizeof *g, NGP_compare);
        for (long k = 0; k + 1 < n_gp; k++)
            if (g[k].energy == g[k + 1].energy) has_duplicate = 1;
    }
    return has_duplicate;
}

/* ---- unionized energy grid --------------------------------------------------------- */
static void merge_runs(const double *a, long na, const double *b, long nb, double *out)
{
    long i = 0, j = 0, o = 0;
    while (i < na && j < nb) out[o++] = (b[j] < a[i]) ? b[j++] : a[i++];
    while (i < na) out[o++] = a[i++];
    while (j < nb) out[o++] = b[j++];
}

static void sort_doubles(double *v, long n)
{
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    int runs = 1;
    while (runs < nthreads && runs < 64) runs *= 2;
    if (n < 1 << 16) runs = 1;
    if (runs == 1) { qsort(v, (size_t)n, sizeof(double), double_compare); return; }

    long *cut = (long *)malloc((size_t)(runs + 1) * sizeof(long));
    for (int r = 0; r <= runs; r++) cut[r] = n * r / runs;
    #pragma omp parallel for schedule(dynamic, 1)
    for (int r = 0; r < runs; r++)
        qsort(v + cut[r], (size_t)(cut[r + 1] - cut[r]), sizeof(double), double_compare);

    double *tmp = (double *)malloc((size_t)n * sizeof(double));
    assert(tmp != NULL);
    double *src = v, *dst = tmp;
    for (int width = 1; width < runs; width *= 2) {
        #pragma omp parallel for schedule(dynamic, 1)
        for (int r = 0; r < runs; r += 2 * width) {
            const long lo = cut[r], mid = cut[r + width], hi = cut[r + 2 * width];
            merge_runs(src + lo, mid - lo, src + mid, hi - mid, dst + lo);
        }
        double *t = src; src = dst; dst = t;
    }
    if (src != v) memcpy(v, src, (size_t)n * sizeof(double));
    free(tmp);
    free(cut);
}

/* ---- index grid -------------------------------------------------------------------- */
/* The reference's sweep over rows [e_begin, e_end), given cursors valid for row e_begin-1. */
static void sweep_rows(const SimulationData *sd, long n_iso, long n_gp, long e_begin, long e_end,
                       int *cursor, double *next_energy)
{
    const NuclideGridPoint *grid = sd->nuclide_grid;
    for (long e = e_begin; e < e_end; e++) {
        const double ue = sd->unionized_energy_array[e];
        int *row = sd->index_grid + e * n_iso;
        for (long i = 0; i < n_iso; i++) {
            if (ue >= next_energy[i] && cursor[i] != n_gp - 2) {
                cursor[i]++;
                next_energy[i] = grid[i * n_gp + cursor[i] + 1].energy;
            }
            row[i] = cursor[i];
        }
    }
}

static void build_index_grid(const SimulationData *sd, long n_iso, long n_gp, int serial_only)
{
    const long n_rows = n_iso * n_gp;
    int blocks = 1;
#ifdef _OPENMP
    if (!serial_only && n_rows >= 4096) blocks = 8 * omp_get_max_threads();
#endif
    (void)serial_only;
    #pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < blocks; b++) {
        const long e_begin = n_rows * b / blocks, e_end = n_rows * (b + 1) / blocks;
        int    *cursor = (int *)malloc((size_t)n_iso * sizeof(int));
        double *next_energy = (double *)malloc((size_t)n_iso * sizeof(double));
        for (long i = 0; i < n_iso; i++) {
            const NuclideGridPoint *g = sd->nuclide_grid + i * n_gp;
            long c = 0;
            if (e_begin > 0) {
                /* #{k >= 1 : g[k].energy <= q}, capped at n_gp-2 */
                const double q = sd->unionized_energy_array[e_begin - 1];
                if (n_gp >= 2 && g[1].energy <= q) {
                    c = (q >= g[n_gp - 1].energy) ? n_gp - 1 : bracket_search(g, q, 1, n_gp - 1);
                    if (c > n_gp - 2) c = n_gp - 2;
                }
            }
            cursor[i] = (int)c;
            next_energy[i] = g[c + 1].energy;
        }
        sweep_rows(sd, n_iso, n_gp, e_begin, e_end, cursor, next_energy);
        free(cursor);
        free(next_energy);
    }
}

/* ---- public ------------------------------------------------------------------------ */
SimulationData grid_init_do_not_profile(Inputs in, int mype)
{
    SimulationData sd;
    memset(&sd, 0, sizeof sd);
    const long n_iso = in.n_isotopes, n_gp = in.n_gridpoints;
    const long n_points = n_iso * n_gp;
    size_t nbytes = 0;
#ifdef _OPENMP
    if (in.nthreads > 0) omp_set_num_threads(in.nthreads);
#endif

    if (mype == 0) printf("Intializing nuclide grids...\n");
    sd.length_nuclide_grid = (int)n_points;
    sd.nuclide_grid = (NuclideGridPoint *)malloc((size_t)n_points * sizeof(NuclideGridPoint));
    assert(sd.nuclide_grid != NULL);
    nbytes += (size_t)n_points * sizeof(NuclideGridPoint);
    const int has_duplicate = fill_nuclide_grid(sd.nuclide_grid, n_iso, n_gp);

    if (in.grid_type == XS_UNIONIZED) {
        if (mype == 0) printf("Intializing unionized grid...\n");
        sd.length_unionized_energy_array = (int)n_points;
        sd.unionized_energy_array = (double *)malloc((size_t)n_points * sizeof(double));
        assert(sd.unionized_energy_array != NULL);
        nbytes += (size_t)n_points * sizeof(double);
        #pragma omp parallel for schedule(static)
        for (long k = 0; k < n_points; k++)
            sd.unionized_energy_array[k] = sd.nuclide_grid[k].energy;
        sort_doubles(sd.unionized_energy_array, n_points);

        sd.length_index_grid = n_points * n_iso;
        sd.index_grid = (int *)malloc((size_t)sd.length_index_grid * sizeof(int));
        assert(sd.index_grid != NULL);
        nbytes += (size_t)sd.length_index_grid * sizeof(int);
        build_index_grid(&sd, n_iso, n_gp, has_duplicate);
    } else if (in.grid_type == XS_HASH) {
        if (mype == 0) printf("Intializing hash grid...\n");
        sd.length_index_grid = (long)in.hash_bins * n_iso;
        sd.index_grid = (int *)malloc((size_t)sd.length_index_grid * sizeof(int));
        assert(sd.index_grid != NULL);
        nbytes += (size_t)sd.length_index_grid * sizeof(int);
        const double du = 1.0 / in.hash_bins;
        #pragma omp parallel for schedule(static)
        for (long bin = 0; bin < in.hash_bins; bin++) {
            const double energy = bin * du;
            for (long i = 0; i < n_iso; i++)
                sd.index_grid[bin * n_iso + i] =
                    (int)bracket_search(sd.nuclide_grid + i * n_gp, energy, 0, n_gp - 1);
        }
    }

    if (mype == 0) printf("Intializing material data...\n");
    sd.num_nucs = load_num_nucs(n_iso);
    sd.length_num_nucs = XS_NUM_MATERIALS;
    sd.mats = load_mats(sd.num_nucs, n_iso, &sd.max_num_nucs);
    sd.length_mats = sd.length_num_nucs * sd.max_num_nucs;
    sd.concs = load_concs(sd.num_nucs, sd.max_num_nucs);
    sd.length_concs = sd.length_mats;

    if (mype == 0)
        printf("Intialization complete. Allocated %.0lf MB of data.\n", nbytes / 1024.0 / 1024.0);
    return sd;
}

SimulationData xs_materials_only(Inputs in)
{
    SimulationData sd;
    memset(&sd, 0, sizeof sd);
    sd.num_nucs = load_num_nucs(in.n_isotopes);
    sd.length_num_nucs = XS_NUM_MATERIALS;
    sd.mats = load_mats(sd.num_nucs, in.n_isotopes, &sd.max_num_nucs);
    sd.length_mats = sd.length_num_nucs * sd.max_num_nucs;
    sd.concs = load_concs(sd.num_nucs, sd.max_num_nucs);
    sd.length_concs = sd.length_mats;
    return sd;
}

void xs_free_simulation_data(SimulationData *sd)
{
    free(sd->num_nucs); free(sd->concs); free(sd->mats);
    free(sd->unionized_energy_array); free(sd->index_grid); free(sd->nuclide_grid);
    memset(sd, 0, sizeof *sd);
}
