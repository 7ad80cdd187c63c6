/*
 * xs_oracle.h -- CPU ORACLE for the XSBench macroscopic cross-section lookup path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference algorithm
 * (ANL-CESAR/XSBench v20, openmp-threading/).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it, and only as the checker.
 * The product (xsbench_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (1) the reference's own 4-entry checksum table (openmp-threading/io.c:85-96),
 *   (2) the unmodified reference compiled from /root/reference (oracle/_ref/libxsref.so):
 *       byte-identical generated data and bit-identical macro_xs vectors,
 *   (3) committed golden vectors under tests/golden/ generated from (2).
 *
 * Every function cites the reference file:line it restates.
 */
#ifndef XS_ORACLE_H
#define XS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XO_GRID_UNIONIZED 0
#define XO_GRID_NUCLIDE   1
#define XO_GRID_HASH      2

/* One energy point of one nuclide: openmp-threading/XSbench_header.h:54-61 (6 x f64 = 48 B). */
typedef struct {
    double e;        /* energy                                  */
    double xs[5];    /* total, elastic, absorbtion, fission, nu_fission */
} xo_point;

/* The generated problem (the six arrays of SimulationData, XSbench_header.h:77-102). */
typedef struct {
    long      n_iso;          /* Inputs.n_isotopes   */
    long      n_gp;           /* Inputs.n_gridpoints */
    int       grid_type;
    int       hash_bins;
    int       max_num_nucs;
    int       num_nucs[12];
    int      *mats;           /* [12 * max_num_nucs] */
    double   *concs;          /* [12 * max_num_nucs] */
    xo_point *nuclide_grid;   /* [n_iso * n_gp]      */
    double   *ueg;            /* [n_iso * n_gp]  (unionized only) */
    int      *index_grid;     /* [n_ueg * n_iso] (unionized) or [hash_bins * n_iso] (hash) */
    long      n_ueg;
    long      n_index;
} xo_data;

/* --- LCG (Simulation.c:462-499) ------------------------------------------------------ */
double   xo_lcg_next(uint64_t *state);
uint64_t xo_lcg_skip(uint64_t state, uint64_t n);
/* --- material sampling (Simulation.c:426-460) ---------------------------------------- */
int      xo_pick_mat(uint64_t *state);
void     xo_mat_thresholds(double thr[12]);
/* --- searches (Simulation.c:380-423) ------------------------------------------------- */
long     xo_search_ueg(long n, double q, const double *a);
long     xo_search_nuclide(double q, const xo_point *a, long lo, long hi);
/* --- lookup arithmetic (Simulation.c:241-375) ---------------------------------------- */
void     xo_micro_xs(const xo_data *d, double e, int nuc, long idx, double out[5]);
void     xo_macro_xs(const xo_data *d, double e, int mat, double out[5]);
/* --- data generation (GridInit.c:3-160, Materials.c:7-117) --------------------------- */
xo_data *xo_generate(long n_iso, long n_gp, int grid_type, int hash_bins);
void     xo_free(xo_data *d);
/* --- drivers ------------------------------------------------------------------------- */
/* Event mode, lookup ids [first, first+n): Simulation.c:15-114. Returns un-modded sum. */
unsigned long long xo_event(const xo_data *d, long first, long n, int nthreads);
/* Same, but also writes per-lookup sample + result (any pointer may be NULL). */
unsigned long long xo_event_dump(const xo_data *d, long first, long n,
                                 double *e_out, int *mat_out, double *macro_out /* n*5 */,
                                 int *argmax_out);
/* Lookups on caller-provided samples (the split sample/lookup form, Simulation.c:698-760). */
unsigned long long xo_lookup_samples(const xo_data *d, long n, const double *e, const int *mat,
                                     double *macro_out /* n*5 or NULL */, int nthreads);
/* History mode, particles [first, first+n) of `lookups` dependent lookups each:
 * Simulation.c:116-238. */
unsigned long long xo_history(const xo_data *d, long first, long n, int lookups, int nthreads);
/* Samples only (energy, material) for ids [first, first+n). */
void xo_sample(long first, long n, double *e_out, int *mat_out);

#ifdef __cplusplus
}
#endif
#endif
