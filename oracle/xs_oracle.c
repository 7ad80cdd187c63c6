/*
 * xs_oracle.c -- CPU ORACLE (test infrastructure only; see xs_oracle.h).
 *
 * Plain-C restatement of the XSBench v20 reference algorithm for the macroscopic
 * cross-section lookup path.  Citations are relative to /root/reference/.
 * Compiled with -ffp-contract=off so that no FMA is formed: the reference is built for
 * baseline x86-64 (no FMA ISA), so every product and sum below rounds like the reference.
 */
#include "xs_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------ */
/* LCG:  x <- (a x + 1) mod 2^63       openmp-threading/Simulation.c:462-470             */
/* ------------------------------------------------------------------------------------ */
#define XO_A     2806196910506780709ULL
#define XO_MASK  0x7FFFFFFFFFFFFFFFULL          /* mod 2^63 == keep the low 63 bits */
#define XO_SEED  1070ULL                        /* STARTING_SEED, XSbench_header.h:51 */

double xo_lcg_next(uint64_t *state)
{
    uint64_t s = (XO_A * (*state) + 1ULL) & XO_MASK;
    *state = s;
    /* (double)s / (double)2^63 : one int->f64 rounding, then an exact power-of-two scale. */
    return (double)s / 9223372036854775808.0;
}

/* Skip-ahead by n steps: openmp-threading/Simulation.c:472-499.  The reference builds the
 * affine map x -> A x + C of n steps by binary decomposition of n; restated here as
 * repeated squaring of the one-step map (a, 1) composed into the accumulator. */
uint64_t xo_lcg_skip(uint64_t state, uint64_t n)
{
    uint64_t mul = XO_A, add = 1ULL;            /* current 2^k-step map         */
    uint64_t acc_mul = 1ULL, acc_add = 0ULL;    /* identity                      */
    n &= XO_MASK;
    for (; n; n >>= 1) {
        if (n & 1ULL) {                          /* acc <- step_k o acc           */
            acc_mul = acc_mul * mul;
            acc_add = acc_add * mul + add;
        }
        add = add * (mul + 1ULL);                /* step_k o step_k               */
        mul = mul * mul;
    }
    return (acc_mul * state + acc_add) & XO_MASK;
}

/* ------------------------------------------------------------------------------------ */
/* pick_mat: openmp-threading/Simulation.c:426-460                                       */
/* The reference recomputes, for each candidate i, running = dist[i]+dist[i-1]+...+dist[1]
 * (that order) and returns the first i with roll < running; dist[0] is never added and
 * material 0 (fuel) is the fall-through.  Thresholds restated with the same add order. */
/* ------------------------------------------------------------------------------------ */
static const double xo_dist[12] = { 0.140, 0.052, 0.275, 0.134, 0.154, 0.064,
                                    0.066, 0.055, 0.008, 0.015, 0.025, 0.013 };

void xo_mat_thresholds(double thr[12])
{
    for (int i = 0; i < 12; i++) {
        double run = 0.0;
        for (int j = i; j > 0; j--) run += xo_dist[j];
        thr[i] = run;
    }
}

int xo_pick_mat(uint64_t *state)
{
    double thr[12];
    xo_mat_thresholds(thr);
    double roll = xo_lcg_next(state);
    for (int i = 0; i < 12; i++)
        if (roll < thr[i]) return i;
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Binary searches: openmp-threading/Simulation.c:380-400 (UEG) and :403-423 (nuclide).  */
/* Invariant a[lo] <= q < a[hi]; stops when hi - lo <= 1; returns lo.                    */
/* ------------------------------------------------------------------------------------ */
long xo_search_ueg(long n, double q, const double *a)
{
    long lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        long mid = lo + (hi - lo) / 2;
        if (a[mid] > q) hi = mid; else lo = mid;
    }
    return lo;
}

long xo_search_nuclide(double q, const xo_point *a, long lo, long hi)
{
    while (hi - lo > 1) {
        long mid = lo + (hi - lo) / 2;
        if (a[mid].e > q) hi = mid; else lo = mid;
    }
    return lo;
}

/* ------------------------------------------------------------------------------------ */
/* calculate_micro_xs: openmp-threading/Simulation.c:241-323                              */
/* ------------------------------------------------------------------------------------ */
void xo_micro_xs(const xo_data *d, double e, int nuc, long idx, double out[5])
{
    const xo_point *g = d->nuclide_grid + (long)nuc * d->n_gp;
    const long last = d->n_gp - 1;
    long low;

    if (d->grid_type == XO_GRID_NUCLIDE) {                       /* :252-263 */
        low = xo_search_nuclide(e, g, 0, last);
    } else if (d->grid_type == XO_GRID_UNIONIZED) {              /* :264-272 */
        low = d->index_grid[idx * d->n_iso + nuc];
    } else {                                                     /* hash, :273-302 */
        long u_lo = d->index_grid[idx * d->n_iso + nuc];
        long u_hi = (idx == d->hash_bins - 1) ? last
                                              : d->index_grid[(idx + 1) * d->n_iso + nuc] + 1;
        double e_lo = g[u_lo].e, e_hi = g[u_hi].e;
        if (e <= e_lo)      low = 0;
        else if (e >= e_hi) low = last;
        else                low = xo_search_nuclide(e, g, u_lo, u_hi);
    }
    if (low == last) low = last - 1;          /* never read past the nuclide's grid */

    const xo_point *p0 = g + low, *p1 = p0 + 1;
    double f = (p1->e - e) / (p1->e - p0->e);                    /* :307 */
    for (int k = 0; k < 5; k++)                                  /* :310-322 */
        out[k] = p1->xs[k] - f * (p1->xs[k] - p0->xs[k]);
}

/* ------------------------------------------------------------------------------------ */
/* calculate_macro_xs: openmp-threading/Simulation.c:326-375                              */
/* ------------------------------------------------------------------------------------ */
void xo_macro_xs(const xo_data *d, double e, int mat, double out[5])
{
    long idx = -1;
    for (int k = 0; k < 5; k++) out[k] = 0.0;

    if (d->grid_type == XO_GRID_UNIONIZED) {
        idx = xo_search_ueg(d->n_iso * d->n_gp, e, d->ueg);      /* :347 */
    } else if (d->grid_type == XO_GRID_HASH) {
        double du = 1.0 / d->hash_bins;                          /* :350-351: two roundings */
        idx = (long)(e / du);
    }
    const int    *nucs = d->mats  + (long)mat * d->max_num_nucs;
    const double *conc = d->concs + (long)mat * d->max_num_nucs;
    for (int j = 0; j < d->num_nucs[mat]; j++) {                 /* :364-374, in j order */
        double xs[5];
        xo_micro_xs(d, e, nucs[j], idx, xs);
        for (int k = 0; k < 5; k++) out[k] += xs[k] * conc[j];
    }
}

/* first index of the strict maximum, start value -1.0: Simulation.c:100-110 */
static inline int xo_argmax5(const double v[5])
{
    double best = -1.0; int at = 0;
    for (int k = 0; k < 5; k++) if (v[k] > best) { best = v[k]; at = k; }
    return at;
}

/* ------------------------------------------------------------------------------------ */
/* Data generation                                                                       */
/* ------------------------------------------------------------------------------------ */
static int xo_cmp_double(const void *a, const void *b)          /* XSutils.c:3-14 */
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}
static int xo_cmp_point(const void *a, const void *b)           /* XSutils.c:16-27 */
{
    double x = ((const xo_point *)a)->e, y = ((const xo_point *)b)->e;
    return (x > y) - (x < y);
}

/* Materials.c:7-31 (counts), :34-96 (nuclide id lists), :99-116 (concentrations). */
static void xo_materials(xo_data *d)
{
    static const int fuel_head[34] = { 58, 59, 60, 61, 40, 42, 43, 44, 45, 46, 1, 2, 3, 7,
                                       8, 9, 10, 29, 57, 47, 48, 0, 62, 15, 33, 34, 52, 53,
                                       54, 55, 56, 18, 23, 41 };
    static const int clad[5]   = { 63, 64, 65, 66, 67 };
    static const int water[4]  = { 24, 41, 4, 5 };
    static const int rpv[27]   = { 19, 20, 21, 22, 35, 36, 37, 38, 39, 25, 27, 28, 29,
                                   30, 31, 32, 26, 49, 50, 51, 11, 12, 13, 14, 6, 16, 17 };
    static const int refl[21]  = { 24, 41, 4, 5, 19, 20, 21, 22, 35, 36, 37, 38, 39, 25,
                                   49, 50, 51, 11, 12, 13, 14 };
    static const int fa[9]     = { 24, 41, 4, 5, 63, 64, 65, 66, 67 };
    static const int counts[12] = { 0, 5, 4, 4, 27, 21, 21, 21, 21, 21, 9, 9 };
    const int *lists[12] = { 0, clad, water, water, rpv, refl, refl, refl, refl, refl, fa, fa };

    memcpy(d->num_nucs, counts, sizeof counts);
    d->num_nucs[0] = (d->n_iso == 68) ? 34 : 321;               /* Materials.c:13-16 */
    d->max_num_nucs = 0;
    for (int m = 0; m < 12; m++)
        if (d->num_nucs[m] > d->max_num_nucs) d->max_num_nucs = d->num_nucs[m];

    const int w = d->max_num_nucs;
    d->mats  = (int *)calloc((size_t)12 * w, sizeof(int));
    d->concs = (double *)calloc((size_t)12 * w, sizeof(double));
    /* Note: the reference mallocs (uninitialised padding); only the first num_nucs[m]
     * entries of each row are defined, and only those are compared by the tests. */
    for (int j = 0; j < d->num_nucs[0]; j++)                    /* Materials.c:45-52 */
        d->mats[j] = (j < 34) ? fuel_head[j] : 68 + (j - 34);
    for (int m = 1; m < 12; m++)
        memcpy(d->mats + (size_t)m * w, lists[m], (size_t)d->num_nucs[m] * sizeof(int));

    uint64_t s = XO_SEED * XO_SEED;                              /* Materials.c:101 */
    for (int m = 0; m < 12; m++)
        for (int j = 0; j < d->num_nucs[m]; j++)
            d->concs[(size_t)m * w + j] = xo_lcg_next(&s);
}

/* grid_init_do_not_profile: GridInit.c:3-160 */
xo_data *xo_generate(long n_iso, long n_gp, int grid_type, int hash_bins)
{
    xo_data *d = (xo_data *)calloc(1, sizeof *d);
    d->n_iso = n_iso; d->n_gp = n_gp; d->grid_type = grid_type; d->hash_bins = hash_bins;
    const long n_pts = n_iso * n_gp;

    /* nuclide grid: 6 sequential draws per point from seed 42 (GridInit.c:12,39-47) ... */
    d->nuclide_grid = (xo_point *)malloc((size_t)n_pts * sizeof(xo_point));
    uint64_t s = 42;
    for (long i = 0; i < n_pts; i++) {
        d->nuclide_grid[i].e = xo_lcg_next(&s);
        for (int k = 0; k < 5; k++) d->nuclide_grid[i].xs[k] = xo_lcg_next(&s);
    }
    /* ... then each nuclide sorted by energy with libc qsort (GridInit.c:50-51) */
    for (long i = 0; i < n_iso; i++)
        qsort(d->nuclide_grid + i * n_gp, (size_t)n_gp, sizeof(xo_point), xo_cmp_point);

    if (grid_type == XO_GRID_UNIONIZED) {
        /* UEG = sorted copy of every energy (GridInit.c:79-89) */
        d->n_ueg = n_pts;
        d->ueg = (double *)malloc((size_t)n_pts * sizeof(double));
        for (long i = 0; i < n_pts; i++) d->ueg[i] = d->nuclide_grid[i].e;
        qsort(d->ueg, (size_t)n_pts, sizeof(double), xo_cmp_double);

        /* index grid sweep (GridInit.c:92-122): one monotone cursor per nuclide */
        d->n_index = n_pts * n_iso;
        d->index_grid = (int *)malloc((size_t)d->n_index * sizeof(int));
        int    *cur  = (int *)calloc((size_t)n_iso, sizeof(int));
        double *next = (double *)malloc((size_t)n_iso * sizeof(double));
        for (long i = 0; i < n_iso; i++) next[i] = d->nuclide_grid[i * n_gp + 1].e;
        for (long e = 0; e < n_pts; e++) {
            const double ue = d->ueg[e];
            int *row = d->index_grid + e * n_iso;
            for (long i = 0; i < n_iso; i++) {
                if (!(ue < next[i]) && cur[i] != n_gp - 2) {
                    cur[i]++;
                    next[i] = d->nuclide_grid[i * n_gp + cur[i] + 1].e;
                }
                row[i] = cur[i];
            }
        }
        free(cur); free(next);
    } else if (grid_type == XO_GRID_HASH) {
        /* hash grid (GridInit.c:125-149): bin e -> search of e*du in every nuclide */
        d->n_index = (long)hash_bins * n_iso;
        d->index_grid = (int *)malloc((size_t)d->n_index * sizeof(int));
        const double du = 1.0 / hash_bins;
        #pragma omp parallel for
        for (long e = 0; e < hash_bins; e++) {
            const double energy = e * du;
            for (long i = 0; i < n_iso; i++)
                d->index_grid[e * n_iso + i] =
                    (int)xo_search_nuclide(energy, d->nuclide_grid + i * n_gp, 0, n_gp - 1);
        }
    }
    xo_materials(d);                                             /* GridInit.c:152-175 */
    return d;
}

void xo_free(xo_data *d)
{
    if (!d) return;
    free(d->mats); free(d->concs); free(d->nuclide_grid); free(d->ueg); free(d->index_grid);
    free(d);
}

/* ------------------------------------------------------------------------------------ */
/* Drivers                                                                               */
/* ------------------------------------------------------------------------------------ */
static void xo_set_threads(int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
}

void xo_sample(long first, long n, double *e_out, int *mat_out)
{
    #pragma omp parallel for schedule(static)
    for (long t = 0; t < n; t++) {
        uint64_t s = xo_lcg_skip(XO_SEED, 2ULL * (uint64_t)(first + t));
        double e = xo_lcg_next(&s);
        int mat = xo_pick_mat(&s);
        if (e_out)   e_out[t] = e;
        if (mat_out) mat_out[t] = mat;
    }
}

/* run_event_based_simulation: Simulation.c:15-114 (lookup i uses LCG outputs 2i+1, 2i+2) */
unsigned long long xo_event_dump(const xo_data *d, long first, long n,
                                 double *e_out, int *mat_out, double *macro_out, int *argmax_out)
{
    unsigned long long sum = 0;
    #pragma omp parallel for schedule(dynamic, 100) reduction(+:sum)
    for (long t = 0; t < n; t++) {
        uint64_t s = xo_lcg_skip(XO_SEED, 2ULL * (uint64_t)(first + t));   /* :63-66 */
        double e = xo_lcg_next(&s);                                        /* :69 */
        int mat = xo_pick_mat(&s);                                         /* :70 */
        double macro[5];
        xo_macro_xs(d, e, mat, macro);
        int am = xo_argmax5(macro);
        sum += (unsigned long long)(am + 1);                               /* :110 */
        if (e_out)      e_out[t] = e;
        if (mat_out)    mat_out[t] = mat;
        if (macro_out)  memcpy(macro_out + 5 * t, macro, sizeof macro);
        if (argmax_out) argmax_out[t] = am;
    }
    return sum;
}

unsigned long long xo_event(const xo_data *d, long first, long n, int nthreads)
{
    xo_set_threads(nthreads);
    return xo_event_dump(d, first, n, 0, 0, 0, 0);
}

unsigned long long xo_lookup_samples(const xo_data *d, long n, const double *e, const int *mat,
                                     double *macro_out, int nthreads)
{
    xo_set_threads(nthreads);
    unsigned long long sum = 0;
    #pragma omp parallel for schedule(dynamic, 100) reduction(+:sum)
    for (long t = 0; t < n; t++) {
        double macro[5];
        xo_macro_xs(d, e[t], mat[t], macro);
        sum += (unsigned long long)(xo_argmax5(macro) + 1);
        if (macro_out) memcpy(macro_out + 5 * t, macro, sizeof macro);
    }
    return sum;
}

/* run_history_based_simulation: Simulation.c:116-238 */
unsigned long long xo_history(const xo_data *d, long first, long n, int lookups, int nthreads)
{
    xo_set_threads(nthreads);
    unsigned long long sum = 0;
    #pragma omp parallel for schedule(dynamic, 100) reduction(+:sum)
    for (long t = 0; t < n; t++) {
        const uint64_t p = (uint64_t)(first + t);
        uint64_t s = xo_lcg_skip(XO_SEED, p * (uint64_t)lookups * 2ULL * 5ULL);  /* :167 */
        double e = xo_lcg_next(&s);
        int mat = xo_pick_mat(&s);
        for (int i = 0; i < lookups; i++) {                                       /* :176 */
            double macro[5];
            xo_macro_xs(d, e, mat, macro);
            sum += (unsigned long long)(xo_argmax5(macro) + 1);
            uint64_t fwd = 0;                                                     /* :225-230 */
            for (int k = 0; k < 5; k++) if (macro[k] > 1.0) fwd++;
            if (fwd) s = xo_lcg_skip(s, fwd);
            e = xo_lcg_next(&s);                                                  /* :232-233 */
            mat = xo_pick_mat(&s);
        }
    }
    return sum;
}
