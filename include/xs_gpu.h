/*
 * xs_gpu.h -- C ABI of the B200-native XSBench lookup engine (libxsb200.so).
 *
 * This header is the drop-in boundary for the reference's event-based GPU path.  The
 * reference has no plugin/FFI layer: cuda/Main.cu calls three groups of functions directly
 * and passes `Inputs` / `SimulationData` by value.  Each entry point below replaces one of
 * those groups (citations are relative to the reference repository, ANL-CESAR/XSBench v20):
 *
 *   xs_gpu_init      <- move_simulation_data_to_device()          cuda/GridInit.cu:4-78
 *   xs_gpu_run       <- run_event_based_simulation_baseline()     cuda/Simulation.cu:15-39
 *                       run_event_based_simulation_optimization_1..6
 *                                                                 cuda/Simulation.cu:388,521,637,754,895,1024
 *                       run_history_based_simulation()            openmp-threading/Simulation.c:116-238
 *                       (history mode is CPU-only in the reference: cuda/Main.cu:84-88)
 *   xs_gpu_finalize  <- release_device_memory()                   cuda/GridInit.cu:81-88
 *
 * Plain C: pointers and sizes only, no C++/torch types.  The structs keep the reference's
 * field order and types so that sizeof(Inputs) == 64, sizeof(NuclideGridPoint) == 48 and
 * sizeof(SimulationData) == 128 with identical offsets (cuda/XSbench_header.cuh:42-85); a
 * reference `main` can pass its own objects by address.
 *
 * Error behaviour: every function returns XS_OK (0) or a negative XS_ERR_* code and never
 * calls exit() (the reference's gpuErrchk does: cuda/XSbench_header.cuh:31-39).  The message
 * for the last failure on the calling thread is returned by xs_gpu_last_error().
 * There is no CPU fallback: without a CUDA device xs_gpu_init fails with XS_ERR_CUDA.
 */
#ifndef XS_GPU_H
#define XS_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants: cuda/XSbench_header.cuh:14-29 -------------------------------------- */
#define XS_UNIONIZED      0
#define XS_NUCLIDE        1
#define XS_HASH           2
#define XS_HISTORY_BASED  1
#define XS_EVENT_BASED    2
#define XS_BINARY_NONE    0
#define XS_BINARY_READ    1
#define XS_BINARY_WRITE   2
#define XS_STARTING_SEED  1070
#define XS_NUM_MATERIALS  12
#define XS_HASH_MODULUS   999983ULL          /* cuda/Main.cu:103 */

#ifndef XS_GPU_NO_REFERENCE_TYPES
/* ---- data model: cuda/XSbench_header.cuh:42-85 ------------------------------------- */
typedef struct {
    double energy;
    double total_xs;
    double elastic_xs;
    double absorbtion_xs;
    double fission_xs;
    double nu_fission_xs;
} NuclideGridPoint;                            /* 48 B */

typedef struct {
    int    nthreads;
    long   n_isotopes;
    long   n_gridpoints;
    int    lookups;
    char  *HM;
    int    grid_type;
    int    hash_bins;
    int    particles;
    int    simulation_method;
    int    binary_mode;
    int    kernel_id;
} Inputs;                                      /* 64 B */

typedef struct {
    int              *num_nucs;                /* [length_num_nucs] = 12 */
    double           *concs;                   /* [length_concs]   = 12 * max_num_nucs */
    int              *mats;                    /* [length_mats]    = 12 * max_num_nucs */
    double           *unionized_energy_array;  /* [length_unionized_energy_array] */
    int              *index_grid;              /* [length_index_grid] */
    NuclideGridPoint *nuclide_grid;            /* [length_nuclide_grid] */
    int               length_num_nucs;
    int               length_concs;
    int               length_mats;
    int               length_unionized_energy_array;
    long              length_index_grid;
    int               length_nuclide_grid;
    int               max_num_nucs;
    unsigned long    *verification;            /* unused by this library */
    int               length_verification;
    double           *p_energy_samples;        /* unused by this library */
    int               length_p_energy_samples;
    int              *mat_samples;             /* unused by this library */
    int               length_mat_samples;
} SimulationData;                              /* 128 B */
#endif

/* ---- status codes ------------------------------------------------------------------ */
#define XS_OK             0
#define XS_ERR_ARG      (-1)   /* NULL / out-of-range argument, unknown kernel id ...   */
#define XS_ERR_CUDA     (-2)   /* CUDA runtime error (no device, OOM, launch failure)  */
#define XS_ERR_UNSUPP   (-3)   /* valid request this build cannot serve                */
#define XS_ERR_NCCL     (-4)   /* NCCL could not be loaded / collective failed         */

/* ---- phases of one run, for per-phase device timing -------------------------------- */
enum { XS_PHASE_SAMPLE = 0, XS_PHASE_SORT = 1, XS_PHASE_LOOKUP = 2, XS_PHASE_REDUCE = 3,
       XS_N_PHASES = 4 };

typedef struct xs_gpu_ctx xs_gpu_ctx;

/* Result of one xs_gpu_run / xs_gpu_lookup_samples call. */
typedef struct {
    unsigned long long verification;     /* un-modded sum of (argmax+1); the caller applies
                                            % XS_HASH_MODULUS like cuda/Main.cu:103          */
    unsigned long long n_lookups;        /* macroscopic lookups performed                     */
    double device_seconds;               /* cudaEvent time, first kernel start -> result ready
                                            (max over devices when n_gpus > 1)                */
    double phase_seconds[XS_N_PHASES];   /* sample / sort / lookup / reduce(+all-reduce)      */
    double host_seconds;                 /* wall clock around the same region incl. copies   */
    unsigned long long h2d_bytes;        /* bytes copied host->device inside the call        */
    unsigned long long d2h_bytes;        /* bytes copied device->host inside the call        */
    int    gpu_launches;                 /* kernels of this library launched by the call     */
    int    n_gpus;
} xs_gpu_result;

/* Static facts about the device-resident problem (for roofline accounting). */
typedef struct {
    int    device;                       /* CUDA ordinal of GPU 0 of this context            */
    int    sm_count;
    long   l2_bytes;
    long   resident_bytes;               /* device bytes held by the context (per GPU)       */
    long   n_isotopes, n_gridpoints;
    int    grid_type, hash_bins, max_num_nucs;
    long   n_ueg;                        /* rows of the unionized grid (0 otherwise)         */
    int    fp64_ops_per_pair;            /* FP64 operations per (lookup, nuclide) of the -k 6 dense kernel: 24 = the reference's
                                            roundings (default), 12 = fused (XSB200_ARITH=fused)                              */
} xs_gpu_info;

/*
 * Upload the problem.  Replaces move_simulation_data_to_device (cuda/GridInit.cu:4-78).
 *   in       : run parameters (n_isotopes, n_gridpoints, grid_type, hash_bins are read)
 *   host_sd  : the six host arrays + lengths, exactly as grid_init_do_not_profile builds them
 *              (cuda/GridInit.cu:90-262).  They stay owned by the caller and may be freed
 *              once this returns.
 *   n_gpus   : 1..8 GPUs of this node driven from the calling thread; the grid is replicated
 *              on each (peer copy from GPU 0), lookups are partitioned, and the only
 *              collective is an all-reduce of {verification, n_lookups}.
 * Device-side generation: if host_sd->nuclide_grid, ->unionized_energy_array and ->index_grid
 * are all NULL (num_nucs / mats / concs / max_num_nucs still given), the synthetic problem of
 * grid_init_do_not_profile is built directly in device memory, byte-identical to the host
 * generator -- nothing large exists on the host or crosses PCIe (XL: 116 GB).
 * Also pre-allocates every scratch buffer (the reference allocates its sample buffers inside
 * the timed region: cuda/Simulation.cu:401-409), builds search-acceleration tables and sets
 * the L2 persistence window.  Device used: the calling thread's current CUDA device, then
 * the following ordinals.
 * A unionized grid that one GPU cannot hold (XXL: 253 GB of index rows) is sharded by energy
 * band when n_gpus > 1: GPU g holds the index rows of band g, every GPU draws every lookup id
 * and keeps the lookups of its band; xs_gpu_run adds the bands up (no data exchange).  In that
 * mode every event variant runs the sorted pipeline; history mode, xs_gpu_lookup_samples and
 * xs_gpu_dump return XS_ERR_UNSUPP.
 */
int xs_gpu_init(const Inputs *in, const SimulationData *host_sd, int n_gpus, xs_gpu_ctx **out);

/*
 * Run one simulation.  Replaces run_event_based_simulation_{baseline,optimization_1..6}
 * (cuda/Simulation.cu:15,388,521,637,754,895,1024) and run_history_based_simulation
 * (openmp-threading/Simulation.c:116-238).  Dispatches on in->simulation_method,
 * in->kernel_id (0..6, event mode) and the grid type fixed at init.  in->lookups (event) or
 * in->particles x in->lookups (history) gives the amount of work.  Synchronous; re-entrant on
 * the same context (warm-up + timed repetitions).  Returns XS_ERR_ARG for an unknown
 * kernel id (the reference prints "No kernel ID" and exits: cuda/Main.cu:78-82).
 */
int xs_gpu_run(xs_gpu_ctx *ctx, const Inputs *in, xs_gpu_result *res);

/*
 * Same as xs_gpu_run, restricted to lookup ids (event) or particle ids (history)
 * [first_id, first_id + count).  This is how one-process-per-GPU callers shard the work:
 * rank r runs its own range and the caller all-reduces {verification, n_lookups}.
 * Lookup i depends only on i (cuda/Simulation.cu:53-56), so any partition is exact.
 */
int xs_gpu_run_range(xs_gpu_ctx *ctx, const Inputs *in, long first_id, long count,
                     xs_gpu_result *res);

/*
 * Lookups on caller-provided samples held in HOST memory: the split form of optimization 1
 * (sampling_kernel -> p_energy_samples / mat_samples -> lookup kernel; cuda/Simulation.cu:
 * 441-509).  Copies the samples host->device, performs n macroscopic lookups, reduces the
 * verification sum and copies the result back; if h_macro_xs_out != NULL also returns the
 * n x 5 macro_xs vectors (total, elastic, absorbtion, fission, nu_fission).
 */
int xs_gpu_lookup_samples(xs_gpu_ctx *ctx, const double *h_energy, const int *h_mat, long n,
                          double *h_macro_xs_out, xs_gpu_result *res);

/*
 * Parity dump: for event-mode lookup ids [first_id, first_id+n) return what the device
 * sampled and computed: energy[n], mat[n], macro_xs[n*5], argmax[n] (any may be NULL).
 */
int xs_gpu_dump(xs_gpu_ctx *ctx, long first_id, long n, double *h_energy_out, int *h_mat_out,
                double *h_macro_xs_out, int *h_argmax_out);

/*
 * Indirect (index-carrying) sort, the building block of the sorted variants: sorts bits
 * [lo_bit, hi_bit) of n 32-bit keys (host) with the library's device radix sort and returns the
 * permutation (host, n entries): keys[perm[0]] <= keys[perm[1]] <= ... on those bits, stable.
 * This is the "sort ids, not particle payloads" scheme the reference recommends for real
 * transport codes (cuda/Simulation.cu:1018-1022) and what -k 6 uses internally with
 * key = material << 28 | energy prefix.
 */
int xs_gpu_sort_keys(xs_gpu_ctx *ctx, const uint32_t *h_keys, long n, int lo_bit, int hi_bit,
                     uint32_t *h_perm_out);

/*
 * Copy a range of one of the device-resident problem arrays (reference layout) back to the
 * host: the device-generated problem can be checked against / saved like a host-generated one
 * (cuda/io.cu:443-470 binary_write).  `which` is one of XS_ARRAY_*.
 */
#define XS_ARRAY_NUCLIDE_GRID      0
#define XS_ARRAY_UNIONIZED_ENERGY  1
#define XS_ARRAY_INDEX_GRID        2
int xs_gpu_read_array(xs_gpu_ctx *ctx, int which, long offset_bytes, long n_bytes, void *h_dst);

/*
 * Self-test of the one place where the kernels do NOT use the reference's own operation: the
 * interpolation factor f = (hi.E - E) / (hi.E - lo.E) is formed from a stored, correctly rounded
 * reciprocal by one Newton-Markstein step (q = n * inv; r = fma(-d, q, n); f = fma(r, inv, q)) instead of
 * an IEEE division (cuda/Simulation.cu:168: 14 instructions instead of 4 on this GPU).  The step yields the
 * correctly rounded quotient whenever q is a faithful approximation, which RN(n * RN(1/d)) is not
 * guaranteed to be in general; this entry point compares the two on `n_pairs` generated (n, d) pairs and
 * returns the number of bit mismatches (expected: 0).  mode 0: pairs as the lookups form them (lo < E <= hi
 * uniform in [0, 1)); mode 1: random mantissas, n <= d, exponents down to 2^-60; mode 2: adversarial
 * divisors (mantissas of all ones / one bit / near powers of two) against random numerators.
 */
int xs_gpu_selftest_division(xs_gpu_ctx *ctx, unsigned long long seed, long n_pairs, int mode,
                             unsigned long long *mismatches);

/*
 * Host-side helper of xs_gpu_lookup_samples, exposed for tests (needs no GPU and no context): the caller's int
 * materials narrowed to bytes by `threads` host threads, as they then cross PCIe (17 MB instead of 68 MB per
 * 17 M samples; a value outside [0, 255] becomes 255, which the device-side validation rejects like any other
 * material outside [0, 12): the call fails with XS_ERR_ARG).  XSB200_HOST_PACK=0 turns the narrowing off, =1 forces
 * it; by default it is on when this process has at least 16 hardware threads per GPU to itself (processes sharing the
 * host are counted from LOCAL_WORLD_SIZE and its MPI / Slurm counterparts); with less, the narrowing is slower than the
 * copy it shortens and the ints travel as they are.
 */
int xs_gpu_narrow_materials(const int *mat, unsigned char *out, long n, int threads);

/* Use `cuda_stream` (a cudaStream_t) for all work of GPU 0 of this context; NULL = default. */
int xs_gpu_set_stream(xs_gpu_ctx *ctx, void *cuda_stream);

/* Free everything (including index_grid, which the reference leaks: cuda/GridInit.cu:81-88). */
int xs_gpu_finalize(xs_gpu_ctx *ctx);

int xs_gpu_get_info(const xs_gpu_ctx *ctx, xs_gpu_info *info);
const char *xs_gpu_last_error(void);
const char *xs_gpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* XS_GPU_H */
