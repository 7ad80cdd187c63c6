// xs_gpu.cu -- C-ABI implementation (include/xs_gpu.h): context, upload, dispatch, timing.
//
// Replaces, behind xs_gpu_init / xs_gpu_run / xs_gpu_finalize, the reference's
// move_simulation_data_to_device (cuda/GridInit.cu:4-78), the seven
// run_event_based_simulation_* drivers (cuda/Simulation.cu:15,388,521,637,754,895,1024),
// run_history_based_simulation (openmp-threading/Simulation.c:116-238) and
// release_device_memory (cuda/GridInit.cu:81-88).
//
// There is no CPU fallback anywhere in this file: every entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "xs_gpu.h"
#include "xs_kernels.cuh"
#include "xs_dense.cuh"
#include "xs_tile.cuh"
#include "xs_sort.cuh"
#include "xs_generate.cuh"
#include "xs_hostpack.h"

namespace {

thread_local char g_error[512] = "";

int set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
    return code;
}

// Every entry point leaves the calling thread's current device as it found it (the multi-GPU paths
// switch devices internally).
struct DeviceGuard {
    int saved = -1;
    DeviceGuard() { if (cudaGetDevice(&saved) != cudaSuccess) saved = -1; }
    ~DeviceGuard() { if (saved >= 0) cudaSetDevice(saved); }
};

#define CUDA_TRY(call)                                                                         \
    do {                                                                                       \
        cudaError_t err__ = (call);                                                            \
        if (err__ != cudaSuccess)                                                              \
            return set_error(XS_ERR_CUDA, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call,   \
                             cudaGetErrorString(err__));                                       \
    } while (0)

double wall_seconds()
{
    using clk = std::chrono::steady_clock;
    return std::chrono::duration<double>(clk::now().time_since_epoch()).count();
}

int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// device counters: [0,16) batch counters of the event kernel, then 16 partition cursors per chunk;
// histogram: 16 material counters per chunk (a run uses chunk 0; host-sample calls are pipelined in chunks)
enum { kMaxChunks = 8, kCursorBase = 16, kNumCounters = kCursorBase + 16 * kMaxChunks, kNumHist = 16 * kMaxChunks };
enum { EV_START = 0, EV_SAMPLED, EV_SORTED, EV_LOOKED_UP, EV_DONE, EV_COUNT };

struct DeviceState {
    int device = -1;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    long l2_bytes = 0;
    // problem
    unsigned char *hot_slab = nullptr;     // [bucket table | UEG] or hash index: L2-persisting
    size_t hot_bytes = 0;
    int *index_grid = nullptr;             // unionized index grid (streamed); in band mode only rows [row0, row1)
    long row0 = 0, row1 = 0;               // this device's energy band of the unionized grid (whole grid: [0, n_ueg))
    int band = 0;                          // band index of this device (band mode)
    double2 *grid = nullptr;
    int *mat_first = nullptr, *mat_nuc = nullptr;
    double *mat_conc = nullptr;
    xs::Problem P{};
    xs::ConcTable conc{};                  // zero-padded concentrations per material: a kernel PARAMETER, so it is this context's own
    size_t resident_bytes = 0;
    // per-run scratch
    unsigned long long *accum = nullptr;   // device [3]: verification sum, lookups done, rejected host samples
    unsigned int *counters = nullptr;      // device [kNumCounters]
    unsigned int *histogram = nullptr;     // device [16]
    unsigned int *dense_counter = nullptr; // lane-per-lookup kernels: [0] next warp-group, [1] warps done (the kernel re-arms both)
    xs::SegTable *seg_tables = nullptr;    // [kMaxChunks][2]: device-built segment tables (sparse, dense) of the host-sample pipeline
    unsigned long long *h_accum = nullptr; // pinned host [3]
    unsigned int *h_hist = nullptr;        // pinned host [16]
    int h_mat_first[XS_NUM_MATERIALS + 1] = {};
    long sample_capacity = 0;
    double *samp_e = nullptr;
    int *samp_mat = nullptr;
    uint32_t *key[2] = {nullptr, nullptr}, *perm[2] = {nullptr, nullptr};
    unsigned int *bin_count = nullptr;     // [12 << bin_bits] fine (material, energy) bins of the -k 6 bin sort
    unsigned int *bin_chunk_sum = nullptr; // scan scratch
    long digits_for = -1;                  // > 0: the kernel that wrote key[0][base..] counted the sort's digits for this many keys (consumed by the next sort)
    long digits_base = -1;
    bool bins_ready = false;               // the event sampler has just counted bin_count for the batch being regrouped
    xs::SortScratch sort{};
    xs::OnesweepScratch sweep{};           // the one-launch-per-digit sort of the lookup keys
    double2 *pairs = nullptr;              // pair records for the window kernel (128 B per grid point)
    uint32_t *nuc_bucket = nullptr;        // nuclide-grid mode: per-nuclide search tables
    uint32_t *samp_where = nullptr;        // [sample_capacity] UEG row / hash bin per sample
    double2 *samp_pack = nullptr;          // [sample_capacity] (energy, row) packed: what the -k 6 kernel reads through the permutation
    double *grp_e = nullptr;               // samples grouped by material: energy,
    uint32_t *grp_where = nullptr;         //   row / bin,
    int *grp_mat = nullptr;                //   material (fuel/other partition only),
    uint32_t *grp_id = nullptr;            //   original sample index
    double2 *sweep_partial = nullptr;      // [sample_capacity*3] partial sums between windows
    uint64_t *hist_seed = nullptr;         // history mode: LCG state per particle
    unsigned char *hist_fwd = nullptr;     // history mode: feedback of the previous lookup
    double *dump_macro = nullptr;          // staging for macro_xs output
    long dump_capacity = 0;
    cudaEvent_t ev[EV_COUNT] = {};
    cudaStream_t copy_stream = nullptr;    // host->device copies of a host-sample call overlap its compute
    uint8_t *h_mat8 = nullptr;             // pinned staging: the caller's materials narrowed to bytes (xs_hostpack.h)
    long h_mat8_capacity = 0;
    cudaEvent_t ev_copy[kMaxChunks] = {}, ev_ready = nullptr;
    cudaEvent_t ev_ecopy[kMaxChunks] = {}; // XSB200_E2E_TRACE=1: end of every chunk's energy copy
    cudaEvent_t ev_chunk[kMaxChunks] = {}; // XSB200_E2E_TRACE=1: end of every chunk's compute (timeline of a host-sample call on stderr)
    int launches = 0;
    // a grouped batch whose histogram read-back has been enqueued but not yet consumed
    struct Pending { bool active = false; int kernel_id = 0; long base = 0, count = 0; const uint32_t *id = nullptr;
                     bool indirect = false; xs::BatchSink sink{}; } pending;
};

}  // namespace

struct xs_gpu_ctx {
    std::vector<DeviceState> dev;
    int grid_type = 0;
    long n_iso = 0, n_gp = 0;
    int hash_bins = 0, max_num_nucs = 0;
    long n_ueg = 0;
    int gather = xs::kTriple;
    int blocks_per_sm = 0;                 // 0 = from occupancy
    int sweep = 1;                         // sorted variants use the windowed nuclide sweep kernel
    int e2e_chunks = 0;                    // host-sample pipeline depth (0 = by size)
    int host_pack = 1;                     // host-sample calls: materials cross PCIe as bytes (XSB200_HOST_PACK=0: as the caller's ints)
    int pack_threads = 4;                  // ... narrowed by this many host threads (XSB200_PACK_THREADS; the caller's thread is one of them)
    xs::PackPool *pack_pool = nullptr;
    long max_pass = 1L << 26;              // lookups materialised at once by -k >= 1 (7.5 GB of buffers)
    int bin_bits = 0;                      // -k 6: energy bits of the one-pass bin sort; 0 (default) = three-pass radix sort,
                                           // which measured the same total (5.49 vs 5.56 ms) and keeps a deterministic order
    int n_bands = 1;                       // > 1: energy-band sharding of the unionized index grid (SURVEY 8e option 2):
                                           // device g holds rows of band g, samples every lookup id and keeps those in its band
    int pack_samples = 1;                  // -k 6: samples also stored as 16-byte (energy, row) records for the permuted reads
    int fuse_gather = 1;                   // -k 6: the lookup kernel reads its samples through the sort's permutation
    int e2e_kernel = 6;                    // xs_gpu_lookup_samples: 6 = sort + lane-per-lookup kernel, 4 = partition + windowed sweep
    int sorted_kernel = 1;                 // -k 6: lane-per-lookup kernel on the energy-sorted batch (0 = windowed sweep)
    int window = 32;                       // nuclides per window (x 1.45 MB of pair records each at n_gp = 11303)
    int tile = 1;                          // -k 0..3 / xs_gpu_dump: xs_tile_kernel (tiles grouped in shared memory, windowed sweep, grid
                                           // barrier per round); 0 = round 1's xs_event_kernel (XSB200_TILE=0)
    int tile_barrier = -1;                 // XSB200_TILE_BARRIER: grid barrier between the rounds of xs_tile_kernel (cooperative launch); -1 = by
                                           // grid type: measured (-k 0, 17 M, ms with / without) unionized 13.6 / 12.9, hash 41.9 / 59.6, nuclide 35.8 / 44.4
    int fuse_digits = 1;                   // ... whose digit counts the sampler / locate kernels take while they write the keys (XSB200_FUSE_DIGITS=0: a pass of its own)
    int onesweep = 1;                      // the lookup sort: one launch per digit (XSB200_ONESWEEP=0: round 1's three kernels per pass)
    int device_segments = 1;               // -k 6 / host-sample pipeline: segment tables built on the device, no histogram read-back
    int e2e_split[kMaxChunks] = {};        // XSB200_E2E_SPLIT: chunk sizes of a host-sample call, in percent (0 = built-in schedule)
    int exact_arith = 1;                   // xs_dense_kernel: 1 = the reference's roundings (24 FP64 operations per (lookup, nuclide), macro_xs
                                           // bit-identical), 0 = fused (12 operations, within 1e-12, integers guarded): XSB200_ARITH=fused
    int dense_min = 64;                    // -k 6: materials with >= this many lookups per grid interval go to xs_dense_kernel (0 = never)
    int key_lo_bit = 8;                    // -k 6 sorts key bits [key_lo_bit, 32): material + 20 energy bits
    int num_nucs[XS_NUM_MATERIALS] = {};
    size_t smem_bytes = 0;
    void *nccl = nullptr;                  // multi-GPU collective state (xs_multi.cuh)
};

#include "xs_multi.cuh"

namespace {

// ---------------------------------------------------------------------------------------
// kernel selection
// ---------------------------------------------------------------------------------------
typedef void (*EventKernel)(const xs::Problem, const xs::BatchSource, const xs::BatchSink);
typedef void (*HistoryKernel)(const xs::Problem, long, long, int, const xs::BatchSink);

EventKernel event_kernel(int grid, int gather)
{
    using namespace xs;
    static const EventKernel table[3][2] = {
        { xs_event_kernel<kUnionized, kLanePerNuclide>, xs_event_kernel<kUnionized, kTriple> },
        { xs_event_kernel<kNuclide,   kLanePerNuclide>, xs_event_kernel<kNuclide,   kTriple> },
        { xs_event_kernel<kHash,      kLanePerNuclide>, xs_event_kernel<kHash,      kTriple> },
    };
    return table[grid][gather];
}

typedef void (*WindowKernel)(const xs::Problem, const xs::WindowArgs, const xs::BatchSink, const xs::ConcTable);
WindowKernel window_kernel(int grid)
{
    using namespace xs;
    static const WindowKernel table[3] = { xs_window_kernel<kUnionized>, xs_window_kernel<kNuclide>, xs_window_kernel<kHash> };
    return table[grid];
}

HistoryKernel history_kernel(int grid, int gather)
{
    using namespace xs;
    static const HistoryKernel table[3][2] = {
        { xs_history_kernel<kUnionized, kLanePerNuclide>, xs_history_kernel<kUnionized, kTriple> },
        { xs_history_kernel<kNuclide,   kLanePerNuclide>, xs_history_kernel<kNuclide,   kTriple> },
        { xs_history_kernel<kHash,      kLanePerNuclide>, xs_history_kernel<kHash,      kTriple> },
    };
    return table[grid][gather];
}

int persistent_grid(const xs_gpu_ctx *ctx, const DeviceState &d, const void *kernel, int *blocks, long smem = -1,
                    int threads = xs::kBlockThreads)
{
    const size_t dyn_smem = smem < 0 ? ctx->smem_bytes : (size_t)smem;
    int per_sm = ctx->blocks_per_sm;
    if (per_sm <= 0) {
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem));
        if (per_sm < 1) per_sm = 1;
    }
    *blocks = per_sm * d.sm_count;
    return XS_OK;
}

// ---------------------------------------------------------------------------------------
// upload
// ---------------------------------------------------------------------------------------
int upload_device(xs_gpu_ctx *ctx, DeviceState &d, const Inputs *in, const SimulationData *sd,
                  const DeviceState *peer)
{
    CUDA_TRY(cudaSetDevice(d.device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, d.device));
    d.sm_count = prop.multiProcessorCount;
    d.l2_bytes = prop.l2CacheSize;
    CUDA_TRY(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    d.own_stream = true;
    for (int i = 0; i < EV_COUNT; i++) CUDA_TRY(cudaEventCreate(&d.ev[i]));

    const long n_iso = in->n_isotopes, n_gp = in->n_gridpoints;
    const long n_points = n_iso * n_gp;
    xs::Problem &P = d.P;
    P.n_iso = (int)n_iso;
    P.n_gp = (int)n_gp;
    P.hash_bins = in->hash_bins;

    auto copy_in = [&](void *dst, const void *host_src, const void *peer_src, size_t bytes) -> cudaError_t {
        if (peer) return cudaMemcpyPeerAsync(dst, d.device, peer_src, peer->device, bytes, d.stream);
        return cudaMemcpyAsync(dst, host_src, bytes, cudaMemcpyHostToDevice, d.stream);
    };

    const bool generate = !peer && !sd->nuclide_grid;          // build the problem on the device

    // nuclide grid (48-byte points, viewed as 16-byte chunks)
    const size_t grid_bytes = (size_t)n_points * sizeof(NuclideGridPoint);
    CUDA_TRY(cudaMalloc(&d.grid, grid_bytes));
    if (!generate) CUDA_TRY(copy_in(d.grid, sd->nuclide_grid, peer ? peer->grid : nullptr, grid_bytes));
    d.resident_bytes += grid_bytes;
    P.grid = d.grid;

    // search structures
    double *ueg = nullptr;
    uint32_t *bucket = nullptr;
    long n_buckets = 0;
    if (ctx->grid_type == XS_UNIONIZED) {
        const long n_ueg = n_points;
        // one bucket per row on average: the <= 4 rows of a bucket are probed at once, and a warp's 32 lanes
        // search 32 buckets -- with 2 rows per bucket 83 % of the warps had a lane whose bucket held more than 4
        // and took a dependent binary-search step first, now 11 % (sampler 0.22 -> 0.20 ms; a table with the
        // bucket's energies inline was slower: 64 MB of cold sectors per step, profiles/r02_notes.md)
        n_buckets = n_ueg;
        if (n_buckets < 1) n_buckets = 1;
        if (n_buckets > (1L << 24)) n_buckets = 1L << 24;
        n_buckets = env_int("XSB200_BUCKETS", (int)n_buckets);
        const size_t bucket_bytes = (((size_t)n_buckets + 1) * sizeof(uint32_t) + 255) / 256 * 256;
        const size_t ueg_bytes = (size_t)n_ueg * sizeof(double);
        d.hot_bytes = bucket_bytes + ueg_bytes;
        CUDA_TRY(cudaMalloc(&d.hot_slab, d.hot_bytes));
        bucket = reinterpret_cast<uint32_t *>(d.hot_slab);
        ueg = reinterpret_cast<double *>(d.hot_slab + bucket_bytes);
        d.row0 = n_ueg * d.band / ctx->n_bands;
        d.row1 = n_ueg * (d.band + 1) / ctx->n_bands;
        const size_t index_bytes = (size_t)(d.row1 - d.row0) * (size_t)n_iso * sizeof(int);
        if (!generate) {
            // (band mode never copies from a peer: every device holds different rows)
            CUDA_TRY(cudaMalloc(&d.index_grid, index_bytes));
            CUDA_TRY(copy_in(ueg, sd->unionized_energy_array, peer ? peer->hot_slab + bucket_bytes : nullptr, ueg_bytes));
            CUDA_TRY(copy_in(d.index_grid, sd->index_grid ? sd->index_grid + d.row0 * n_iso : nullptr,
                             peer ? peer->index_grid : nullptr, index_bytes));
        }
        d.resident_bytes += d.hot_bytes + index_bytes;
        P.ueg = ueg;
        P.ueg_bucket = bucket;
        P.n_ueg = n_ueg;
        P.n_buckets = (int)n_buckets;
        P.bucket_scale = (double)n_buckets;
        P.bucket_shift = (n_ueg < (1L << 28) && env_int("XSB200_BUCKET_PACK", 1)) ? 4 : 0;
    } else if (ctx->grid_type == XS_HASH) {
        d.hot_bytes = (size_t)in->hash_bins * (size_t)n_iso * sizeof(int);
        CUDA_TRY(cudaMalloc(&d.hot_slab, d.hot_bytes));
        if (!generate) CUDA_TRY(copy_in(d.hot_slab, sd->index_grid, peer ? peer->hot_slab : nullptr, d.hot_bytes));
        d.resident_bytes += d.hot_bytes;
        P.index_grid = reinterpret_cast<const int *>(d.hot_slab);
    }
    if (generate) {
        int *index_dst = ctx->grid_type == XS_UNIONIZED ? nullptr : reinterpret_cast<int *>(d.hot_slab);
        int g = xs::generate_problem(ctx->grid_type, n_iso, n_gp, in->hash_bins, d.grid, ueg, index_dst, d.sm_count, d.stream);
        if (g == -2)
            return set_error(XS_ERR_UNSUPP, "device generator: a nuclide holds two equal energies; generate on the host instead");
        if (g == 0 && ctx->grid_type == XS_UNIONIZED) {
            // the generator's sort temporaries are gone: now the (possibly very large) index rows
            const size_t index_bytes = (size_t)(d.row1 - d.row0) * (size_t)n_iso * sizeof(int);
            CUDA_TRY(cudaMalloc(&d.index_grid, index_bytes));
            g = xs::generate_index_rows(ueg, d.grid, n_iso, n_gp, d.row0, d.row1, d.index_grid, d.stream);
        }
        if (g != 0)
            return set_error(XS_ERR_CUDA, "device generator failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    // kernels index the grid by global row number: bias the pointer by the band's first row
    if (ctx->grid_type == XS_UNIONIZED) P.index_grid = d.index_grid - d.row0 * n_iso;
    if (ctx->grid_type == XS_UNIONIZED) {
        if (P.bucket_shift) {
            // first rows into a temporary, then entries with the bucket's row count packed in (xs_pack_buckets_kernel)
            uint32_t *first = nullptr;
            CUDA_TRY(cudaMalloc(&first, ((size_t)n_buckets + 1) * sizeof(uint32_t)));
            xs::xs_build_buckets_kernel<<<d.sm_count * 8, 256, 0, d.stream>>>(ueg, n_points, (double)n_buckets, (int)n_buckets, first);
            xs::xs_pack_buckets_kernel<<<d.sm_count * 8, 256, 0, d.stream>>>(first, (int)n_buckets, bucket);
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
            cudaFree(first);
            CUDA_TRY(e);
        } else {
            xs::xs_build_buckets_kernel<<<d.sm_count * 8, 256, 0, d.stream>>>(ueg, n_points, (double)n_buckets, (int)n_buckets, bucket);
            CUDA_TRY(cudaGetLastError());
        }
    }
    if (ctx->grid_type == XS_NUCLIDE && env_int("XSB200_NUCLIDE_BUCKETS", 1)) {
        // nuclide-grid mode: a bucket table per nuclide replaces the top ~12 of the 14 (large)
        // dependent probes of grid_search_nuclide; ~2 grid points per bucket
        long nb = 1;
        while (nb * 2 <= n_gp / 2 && nb < 65536) nb *= 2;
        const size_t nb_bytes = (size_t)n_iso * (size_t)(nb + 1) * sizeof(uint32_t);
        CUDA_TRY(cudaMalloc(&d.nuc_bucket, nb_bytes));
        xs::xs_build_nuclide_buckets_kernel<<<d.sm_count * 8, 256, 0, d.stream>>>(d.grid, n_iso, n_gp, (int)nb, d.nuc_bucket);
        CUDA_TRY(cudaGetLastError());
        d.resident_bytes += nb_bytes;
        P.nuc_bucket = d.nuc_bucket;
        P.nuc_buckets = (int)nb;
    }
    // pair records (B200 layout for the windowed sweep): one 128-byte line per (nuclide, k)
    // (+ kSpan records of padding: the sorted kernel copies kSpan consecutive records at a time)
    const size_t pair_bytes = ((size_t)n_points + xs::kSpan) * 8 * sizeof(double2);
    CUDA_TRY(cudaMalloc(&d.pairs, pair_bytes));
    CUDA_TRY(cudaMemsetAsync(d.pairs + (size_t)n_points * 8, 0, (size_t)xs::kSpan * 8 * sizeof(double2), d.stream));
    xs::xs_build_pairs_kernel<<<d.sm_count * 8, 256, 0, d.stream>>>(d.grid, n_iso, n_gp, d.pairs);
    CUDA_TRY(cudaGetLastError());
    d.resident_bytes += pair_bytes;
    P.pairs = d.pairs;

    // compact (CSR) material tables
    int first[XS_NUM_MATERIALS + 1];
    first[0] = 0;
    for (int m = 0; m < XS_NUM_MATERIALS; m++) first[m + 1] = first[m] + sd->num_nucs[m];
    const int total = first[XS_NUM_MATERIALS];
    std::vector<int> nuc(total);
    std::vector<double> conc(total);
    for (int m = 0; m < XS_NUM_MATERIALS; m++)
        for (int j = 0; j < sd->num_nucs[m]; j++) {
            nuc[first[m] + j] = sd->mats[(size_t)m * sd->max_num_nucs + j];
            conc[first[m] + j] = sd->concs[(size_t)m * sd->max_num_nucs + j];
        }
    memcpy(d.h_mat_first, first, sizeof first);
    CUDA_TRY(cudaMalloc(&d.mat_first, sizeof first));
    CUDA_TRY(cudaMalloc(&d.mat_nuc, (size_t)total * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d.mat_conc, (size_t)total * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(d.mat_first, first, sizeof first, cudaMemcpyHostToDevice, d.stream));
    CUDA_TRY(cudaMemcpyAsync(d.mat_nuc, nuc.data(), (size_t)total * sizeof(int), cudaMemcpyHostToDevice, d.stream));
    CUDA_TRY(cudaMemcpyAsync(d.mat_conc, conc.data(), (size_t)total * sizeof(double), cudaMemcpyHostToDevice, d.stream));
    P.mat_first = d.mat_first;
    P.mat_nuc = d.mat_nuc;
    P.mat_conc = d.mat_conc;
    P.mat_total = total;
    ctx->smem_bytes = sizeof(xs::SharedTables) + (size_t)total * (sizeof(double) + sizeof(int));

    // material thresholds: same summation order as pick_mat (cuda/Simulation.cu:314-321)
    static const double frac[XS_NUM_MATERIALS] = { 0.140, 0.052, 0.275, 0.134, 0.154, 0.064,
                                                   0.066, 0.055, 0.008, 0.015, 0.025, 0.013 };
    double thr[XS_NUM_MATERIALS];
    for (int i = 0; i < XS_NUM_MATERIALS; i++) {
        double acc = 0.0;
        for (int j = i; j >= 1; j--) acc += frac[j];
        thr[i] = acc;
    }
    memcpy(P.mat_threshold, thr, sizeof thr);
    // zero-padded concentration rows (kernel parameter, see xs::ConcTable)
    {
        int at = 0;
        memset(&d.conc, 0, sizeof d.conc);
        for (int m = 0; m < XS_NUM_MATERIALS; m++) {
            d.conc.first[m] = at;
            const int padded = (sd->num_nucs[m] + 7) / 8 * 8 + xs::kConcPad;
            if (at + padded > xs::kConcCap)
                return set_error(XS_ERR_UNSUPP, "the materials hold %d nuclide entries (padded), more than the %d this build supports",
                                 at + padded, xs::kConcCap);
            for (int j = 0; j < sd->num_nucs[m]; j++) d.conc.v[at + j] = sd->concs[(size_t)m * sd->max_num_nucs + j];
            at += padded;
        }
    }

    // run scratch
    CUDA_TRY(cudaMalloc(&d.accum, 3 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&d.counters, kNumCounters * sizeof(unsigned int)));
    CUDA_TRY(cudaMalloc(&d.histogram, kNumHist * sizeof(unsigned int)));
    CUDA_TRY(cudaMalloc(&d.dense_counter, 2 * sizeof(unsigned int)));
    CUDA_TRY(cudaMemsetAsync(d.dense_counter, 0, 2 * sizeof(unsigned int), d.stream));
    CUDA_TRY(cudaMalloc(&d.seg_tables, (size_t)kMaxChunks * 2 * sizeof(xs::SegTable)));
    CUDA_TRY(cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < kMaxChunks; i++) CUDA_TRY(cudaEventCreate(&d.ev_copy[i]));
    for (int i = 0; i < kMaxChunks; i++) CUDA_TRY(cudaEventCreate(&d.ev_chunk[i]));
    for (int i = 0; i < kMaxChunks; i++) CUDA_TRY(cudaEventCreate(&d.ev_ecopy[i]));
    CUDA_TRY(cudaEventCreateWithFlags(&d.ev_ready, cudaEventDisableTiming));
    CUDA_TRY(cudaMallocHost(&d.h_accum, 3 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMallocHost(&d.h_hist, 16 * sizeof(unsigned int)));

    CUDA_TRY(cudaStreamSynchronize(d.stream));   // host vectors above go out of scope

    // Optional (XSB200_L2_WINDOW=1, off by default): persisting access-policy window over the
    // search structures (bucket table + UEG, or the hash grid).  Measured on B200 it does not
    // pay: the L2 is better spent on the nuclide grid (see the policy note in xs_device.cuh).
    if (d.hot_slab && env_int("XSB200_L2_WINDOW", 0)) {
        size_t want = std::min<size_t>(d.hot_bytes, (size_t)prop.persistingL2CacheMaxSize);
        if (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof attr);
            attr.accessPolicyWindow.base_ptr = d.hot_slab;
            attr.accessPolicyWindow.num_bytes = std::min<size_t>(d.hot_bytes, (size_t)prop.accessPolicyMaxWindowSize);
            attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)attr.accessPolicyWindow.num_bytes);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
            cudaStreamSetAttribute(d.stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        }
        cudaGetLastError();   // residency is an optimisation; never fatal
    }
    return XS_OK;
}

int ensure_sample_buffers(DeviceState &d, long n, bool need_sort, int bin_bits = 0)
{
    CUDA_TRY(cudaSetDevice(d.device));
    if (n > d.sample_capacity) {
        cudaFree(d.samp_e); cudaFree(d.samp_mat);
        for (int i = 0; i < 2; i++) { cudaFree(d.key[i]); cudaFree(d.perm[i]); d.key[i] = d.perm[i] = nullptr; }
        cudaFree(d.samp_where); cudaFree(d.samp_pack); cudaFree(d.grp_e); cudaFree(d.grp_where); cudaFree(d.grp_mat); cudaFree(d.grp_id);
        cudaFree(d.sweep_partial); cudaFree(d.hist_seed); cudaFree(d.hist_fwd);
        d.hist_seed = nullptr; d.hist_fwd = nullptr;
        d.samp_where = nullptr; d.samp_pack = nullptr; d.grp_e = nullptr; d.grp_where = nullptr; d.grp_mat = nullptr; d.grp_id = nullptr;
        d.sweep_partial = nullptr;
        d.samp_e = nullptr; d.samp_mat = nullptr;
        CUDA_TRY(cudaMalloc(&d.samp_e, (size_t)n * sizeof(double)));
        CUDA_TRY(cudaMalloc(&d.samp_mat, (size_t)n * sizeof(int)));
        d.sample_capacity = n;
    }
    if (need_sort && !d.key[0]) {
        for (int i = 0; i < 2; i++) {
            CUDA_TRY(cudaMalloc(&d.key[i], (size_t)d.sample_capacity * sizeof(uint32_t)));
            CUDA_TRY(cudaMalloc(&d.perm[i], (size_t)d.sample_capacity * sizeof(uint32_t)));
        }
        const size_t cap = (size_t)d.sample_capacity;
        CUDA_TRY(cudaMalloc(&d.samp_where, cap * sizeof(uint32_t)));
        CUDA_TRY(cudaMalloc(&d.samp_pack, cap * sizeof(double2)));
        CUDA_TRY(cudaMalloc(&d.grp_e, cap * sizeof(double)));
        CUDA_TRY(cudaMalloc(&d.grp_where, cap * sizeof(uint32_t)));
        CUDA_TRY(cudaMalloc(&d.grp_mat, cap * sizeof(int)));
        CUDA_TRY(cudaMalloc(&d.grp_id, cap * sizeof(uint32_t)));
        CUDA_TRY(cudaMalloc(&d.sweep_partial, cap * 3 * sizeof(double2)));
        CUDA_TRY(cudaMalloc(&d.hist_seed, cap * sizeof(uint64_t)));
        CUDA_TRY(cudaMalloc(&d.hist_fwd, cap));
        int rc = xs::sort_scratch_alloc(d.sort, d.sample_capacity);
        if (rc == 0) rc = xs::onesweep_scratch_alloc(d.sweep, d.sample_capacity);
        if (rc != 0) return set_error(XS_ERR_CUDA, "sort scratch allocation failed");
    }
    if (need_sort && bin_bits > 0 && !d.bin_count) {
        const size_t n_bins = (size_t)XS_NUM_MATERIALS << bin_bits;
        CUDA_TRY(cudaMalloc(&d.bin_count, n_bins * sizeof(unsigned int)));
        CUDA_TRY(cudaMalloc(&d.bin_chunk_sum, (n_bins / xs::kScanChunk + 1) * sizeof(unsigned int)));
    }
    return XS_OK;
}

int ensure_dump_buffer(DeviceState &d, long n)
{
    if (n > d.dump_capacity) {
        cudaFree(d.dump_macro);
        d.dump_macro = nullptr;
        CUDA_TRY(cudaMalloc(&d.dump_macro, (size_t)n * 5 * sizeof(double)));
        d.dump_capacity = n;
    }
    return XS_OK;
}

// ---------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------
typedef void (*TileKernel)(const xs::Problem, const xs::BatchSource, const xs::BatchSink, const xs::ConcTable, int, int);

// The in-order variants in one launch with deliberate L2 locality (xs_tile.cuh).  The grid barrier needs
// every block resident: cooperative launch, grid = occupancy x SMs.
int launch_tile(xs_gpu_ctx *ctx, DeviceState &d, const xs::BatchSource &src, xs::BatchSink sink, int counter_slot)
{
    static const TileKernel table[3] = { xs::xs_tile_kernel<xs::kUnionized>, xs::xs_tile_kernel<xs::kNuclide>, xs::xs_tile_kernel<xs::kHash> };
    TileKernel k = table[ctx->grid_type];
    const size_t smem = sizeof(xs::TileShared) + (size_t)d.P.mat_total * sizeof(int) + (XS_NUM_MATERIALS + 1) * sizeof(int);
    CUDA_TRY(cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = 0;
    int rc = persistent_grid(ctx, d, (const void *)k, &blocks, (long)smem);
    if (rc != XS_OK) return rc;
    const long n_tiles = (src.count + xs::kTile - 1) / xs::kTile;
    if (blocks > n_tiles) blocks = (int)n_tiles;
    sink.batch_counter = d.counters + counter_slot;           // the round barrier's arrival counter (zeroed per pass)
    const int quantum = 2 * xs::kSweepUnroll;
    int window = std::max(quantum, std::min(ctx->window, 32) / quantum * quantum);
    int use_barrier = ctx->tile_barrier >= 0 ? ctx->tile_barrier : ctx->grid_type != XS_UNIONIZED;
    if (use_barrier) {
        void *args[] = { (void *)&d.P, (void *)&src, (void *)&sink, (void *)&d.conc, (void *)&window, (void *)&use_barrier };
        const cudaError_t err = cudaLaunchCooperativeKernel((const void *)k, dim3(blocks), dim3(xs::kBlockThreads), args, smem, d.stream);
        if (err == cudaErrorCooperativeLaunchTooLarge || err == cudaErrorNotSupported) {
            // (a device or mode without cooperative launches: same kernel, no barrier -- results do not depend on it)
            cudaGetLastError();
            use_barrier = 0;
        } else {
            CUDA_TRY(err);
        }
    }
    if (!use_barrier) {
        k<<<blocks, xs::kBlockThreads, smem, d.stream>>>(d.P, src, sink, d.conc, window, use_barrier);
        CUDA_TRY(cudaGetLastError());
    }
    d.launches++;
    return XS_OK;
}

int launch_event(xs_gpu_ctx *ctx, DeviceState &d, const xs::BatchSource &src, xs::BatchSink sink,
                 int counter_slot)
{
    if (src.count <= 0) return XS_OK;
    if (ctx->tile && !sink.fwd_out) return launch_tile(ctx, d, src, sink, counter_slot);
    EventKernel k = event_kernel(ctx->grid_type, ctx->gather);
    int blocks = 0;
    int rc = persistent_grid(ctx, d, (const void *)k, &blocks);
    if (rc != XS_OK) return rc;
    const long n_batches = (src.count + 31) / 32;
    const long max_useful = (n_batches + xs::kWarpsPerBlock - 1) / xs::kWarpsPerBlock;
    if (blocks > max_useful) blocks = (int)max_useful;
    sink.batch_counter = d.counters + counter_slot;
    k<<<blocks, xs::kBlockThreads, ctx->smem_bytes, d.stream>>>(d.P, src, sink);
    CUDA_TRY(cudaGetLastError());
    d.launches++;
    return XS_OK;
}

// Buffers of one grouped batch of lookups (a whole run, or one chunk of a host-sample call).
struct GroupedBatch {
    const double *energy;          // grouped by material
    const uint32_t *where;
    const uint32_t *id;            // original sample index per slot (macro_xs dumps)
    bool indirect;                 // energy / where are still in sample order: the kernel reads them through id
    const double2 *pack;           // indirect mode: packed (energy, row) per sample, or null
    double2 *partial;
    long offset[XS_NUM_MATERIALS]; // first slot of each material (host copy of the histogram prefix)
    long count[XS_NUM_MATERIALS];  // lookups per material
    const unsigned int *hist;      // device histogram of this batch (device-built segment tables)
    long total;                    // lookups in the batch
};

// One launch of the window kernel over `a.n_seg` segments.
int launch_window(xs_gpu_ctx *ctx, DeviceState &d, xs::WindowArgs &a, const GroupedBatch &b, xs::BatchSink sink)
{
    long groups = 0;
    for (int i = 0; i < a.n_seg; i++) {
        const int m = a.seg[i].mat;
        a.seg[i].offset = b.offset[m];
        a.seg[i].count = (int)b.count[m];
        a.seg[i].group_begin = (int)groups;
        groups += (b.count[m] + xs::kSweepSlots - 1) / xs::kSweepSlots;
    }
    if (groups == 0) return XS_OK;
    a.n_groups = (int)groups;
    a.energy = b.energy;
    a.where = b.where;
    a.sample_id = b.id;
    a.partial = b.partial;
    WindowKernel k = window_kernel(ctx->grid_type);
    int blocks = 0;
    const size_t smem = (size_t)d.P.mat_total * sizeof(int);
    int rc = persistent_grid(ctx, d, (const void *)k, &blocks, (long)smem);
    if (rc != XS_OK) return rc;
    const long max_useful = (groups + xs::kWarpsPerBlock - 1) / xs::kWarpsPerBlock;
    if (blocks > max_useful) blocks = (int)max_useful;
    k<<<blocks, xs::kBlockThreads, smem, d.stream>>>(d.P, a, sink, d.conc);
    CUDA_TRY(cudaGetLastError());
    d.launches++;
    return XS_OK;
}

// Lane-per-lookup sweep over a batch sorted by (material, energy): the materials with many lookups
// per grid interval (>= ctx->dense_min: a warp-group's 96 lookups then fall into the first lookup's
// interval or the next) go to xs_dense_kernel, the others to xs_sorted_kernel -- two launches.
int launch_sorted(xs_gpu_ctx *ctx, DeviceState &d, const GroupedBatch &b, xs::BatchSink sink, xs::SegTable *dev_tables = nullptr)
{
    static const WindowKernel table[3][3] = {
        { xs::xs_sorted_kernel<xs::kUnionized>, xs::xs_sorted_kernel<xs::kNuclide>, xs::xs_sorted_kernel<xs::kHash> },
        { xs::xs_dense_kernel<xs::kUnionized, true>, xs::xs_dense_kernel<xs::kNuclide, true>, xs::xs_dense_kernel<xs::kHash, true> },
        { xs::xs_dense_kernel<xs::kUnionized, false>, xs::xs_dense_kernel<xs::kNuclide, false>, xs::xs_dense_kernel<xs::kHash, false> } };
    if (dev_tables) {
        // the segments of both launches from the histogram, on the device: the host does not wait for it
        xs::MatShape shape;
        for (int m = 0; m <= XS_NUM_MATERIALS; m++) shape.first[m] = d.h_mat_first[m];
        for (int m = 0; m < XS_NUM_MATERIALS; m++) shape.n_nuc[m] = ctx->num_nucs[m];
        xs::xs_build_segments_kernel<<<1, 32, 0, d.stream>>>(b.hist, shape, ctx->dense_min > 0 ? (long)ctx->dense_min * d.P.n_gp : 0L,
                                                             xs::kDenseGroup, xs::kSortedGroup, dev_tables + 1, dev_tables + 0);
        CUDA_TRY(cudaGetLastError());
        d.launches++;
    }
    for (int dense = 1; dense >= 0; dense--) {
        xs::WindowArgs a{};
        long groups = 0;
        if (!dev_tables) {
            for (int m = 0; m < XS_NUM_MATERIALS; m++) {
                if (b.count[m] <= 0) continue;
                const bool is_dense = ctx->dense_min > 0 && b.count[m] >= (long)ctx->dense_min * d.P.n_gp;
                if (is_dense != (dense == 1)) continue;
                xs::WindowSegment &sgm = a.seg[a.n_seg++];
                sgm.mat = m; sgm.first = d.h_mat_first[m]; sgm.j_begin = 0; sgm.j_end = ctx->num_nucs[m];
                sgm.offset = b.offset[m];
                sgm.count = (int)b.count[m];
                sgm.group_begin = (int)groups;
                const int group = dense ? xs::kDenseGroup : xs::kSortedGroup;
                groups += (b.count[m] + group - 1) / group;
            }
            if (groups == 0) continue;
        } else if (dense && ctx->dense_min <= 0) {
            continue;
        }
        a.dev_table = dev_tables ? dev_tables + dense : nullptr;
        a.n_groups = (int)groups;
        a.energy = b.energy;
        a.where = b.where;
        a.sample_id = b.id;
        a.indirect = b.indirect;
        a.pack = b.pack;
        a.first_window = a.last_window = 1;
        WindowKernel k = table[dense ? (ctx->exact_arith ? 1 : 2) : 0][ctx->grid_type];
        int blocks = 0;
        size_t staged = 0;
        const int threads = dense ? xs::kDenseThreads : xs::kBlockThreads, warps = threads / 32;
        if (dense)
            staged = (size_t)warps * (xs::kDenseRingBytes + xs::kDenseFirstWords * sizeof(uint32_t) + xs::kDenseStageBytes);
        else if (ctx->grid_type == XS_UNIONIZED)
            staged = (size_t)xs::kWarpsPerBlock * (32 * xs::kLaneWords * sizeof(uint32_t) + xs::kRingBytes);
        const size_t smem = staged + (size_t)d.P.mat_total * sizeof(int);
        CUDA_TRY(cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int rc = persistent_grid(ctx, d, (const void *)k, &blocks, (long)smem, threads);
        if (rc != XS_OK) return rc;
        if (!dev_tables) {                                    // (a device-built table: the group count is not known here)
            const long max_useful = (groups + warps - 1) / warps;
            if (blocks > max_useful) blocks = (int)max_useful;
        } else {
            const long most = (b.total + (dense ? xs::kDenseGroup : xs::kSortedGroup) - 1) / (dense ? xs::kDenseGroup : xs::kSortedGroup) + XS_NUM_MATERIALS;
            const long max_useful = (most + warps - 1) / warps;
            if (blocks > max_useful) blocks = (int)max_useful;
        }
        xs::BatchSink launch_sink = sink;
        launch_sink.batch_counter = d.dense_counter;          // group hand-out; the kernel leaves it zeroed
        k<<<blocks, threads, smem, d.stream>>>(d.P, a, launch_sink, d.conc);
        CUDA_TRY(cudaGetLastError());
        d.launches++;
    }
    return XS_OK;
}

// Windowed nuclide sweep over a grouped batch, for the materials in `mats`.  Materials that
// need several windows (fuel) get one launch per window; all single-window materials share one.
int launch_sweep(xs_gpu_ctx *ctx, DeviceState &d, const GroupedBatch &b, int n_mats, const int *mats,
                 xs::BatchSink sink)
{
    // windows of `width` nuclides (a multiple of the gather loop's step quantum, at most 32); a
    // remainder of up to one quantum is folded into the last window instead of costing a launch
    // of its own (fuel: 321 = 9 x 32 + 33)
    const int quantum = 2 * xs::kSweepUnroll;
    const int width = std::max(quantum, std::min(ctx->window, 32) / quantum * quantum);
    xs::WindowArgs small{};
    small.first_window = small.last_window = 1;
    int rc = XS_OK;
    for (int i = 0; i < n_mats && rc == XS_OK; i++) {
        const int m = mats[i], n = ctx->num_nucs[m];
        int passes = (n + width - 1) / width;
        const int folded = n - (passes - 2) * width;           // size of the last window if the remainder is folded in
        if (passes > 1 && (folded + quantum - 1) / quantum * quantum <= xs::kMaxWindow) passes--;
        if (b.count[m] <= 0) continue;
        if (passes == 1) {
            xs::WindowSegment &sgm = small.seg[small.n_seg++];
            sgm.mat = m; sgm.first = d.h_mat_first[m]; sgm.j_begin = 0; sgm.j_end = n;
            continue;
        }
        for (int p = 0; p < passes && rc == XS_OK; p++) {
            xs::WindowArgs a{};
            a.n_seg = 1;
            a.seg[0].mat = m; a.seg[0].first = d.h_mat_first[m];
            a.seg[0].j_begin = p * width;
            a.seg[0].j_end = (p == passes - 1) ? n : (p + 1) * width;
            a.first_window = p == 0;
            a.last_window = p == passes - 1;
            rc = launch_window(ctx, d, a, b, sink);
        }
    }
    if (rc == XS_OK && small.n_seg) rc = launch_window(ctx, d, small, b, sink);
    return rc;
}

int sort_lookup_keys(xs_gpu_ctx *ctx, DeviceState &d, uint32_t *key[2], uint32_t *perm[2], long count, int lo_bit, int hi_bit,
                     uint32_t **sorted_perm, int *launches);

// The kernel about to write `count` sort keys at key[0] + base counts their digits itself (xs_sort.cuh): zero the
// sort's scratch now, remember for which sort.  XSB200_FUSE_DIGITS=0: the sort reads the keys once more instead.
xs::DigitSpec begin_digit_count(xs_gpu_ctx *ctx, DeviceState &d, long base, long count)
{
    xs::DigitSpec spec;
    d.digits_for = -1;
    if (ctx->onesweep && ctx->fuse_digits && count > 0 && xs::onesweep_begin(d.sweep, count, ctx->key_lo_bit, 32, d.stream, &spec)) {
        d.digits_for = count;
        d.digits_base = base;
    }
    return spec;
}

int launch_sample(xs_gpu_ctx *ctx, DeviceState &d, long first_id, long count, bool with_where, bool with_key,
                  bool with_hist)
{
    int blocks = (int)std::min<long>((count + 255) / 256, (long)d.sm_count * 16);
    const bool with_bins = with_key && ctx->bin_bits > 0;
    if (with_bins)
        CUDA_TRY(cudaMemsetAsync(d.bin_count, 0, ((size_t)XS_NUM_MATERIALS << ctx->bin_bits) * sizeof(unsigned int), d.stream));
    // -k 6 on the default path reads its samples only as packed (energy, row) records through the
    // sort permutation: the separate arrays are not written then
    const bool pack_only = with_key && ctx->pack_samples && ctx->fuse_gather && ctx->sorted_kernel && !with_bins;
    const xs::DigitSpec digits = with_key && !with_bins ? begin_digit_count(ctx, d, 0, count) : xs::DigitSpec{};
    xs::xs_sample_kernel<<<blocks, 256, 0, d.stream>>>(d.P, ctx->grid_type, first_id, count,
                                                       pack_only ? nullptr : d.samp_e, pack_only ? nullptr : d.samp_mat,
                                                       with_where && !pack_only ? d.samp_where : nullptr,
                                                       with_key ? d.key[0] : nullptr,
                                                       with_hist ? d.histogram : nullptr,
                                                       with_bins ? d.bin_count : nullptr, 28 - ctx->bin_bits,
                                                       ctx->n_bands > 1 ? (uint32_t)d.row0 : 0u,
                                                       ctx->n_bands > 1 ? (uint32_t)d.row1 : 0xffffffffu,
                                                       with_key && ctx->pack_samples ? d.samp_pack : nullptr, digits);
    CUDA_TRY(cudaGetLastError());
    d.launches++;
    d.bins_ready = with_bins;
    return XS_OK;
}

// Sorted variants on device-resident samples (samp_e / samp_mat / samp_where, the material
// histogram and, for -k 6, key[0] are ready in device memory): regroup, then sweep material by
// material.  `base` offsets every per-sample buffer (chunks of a host-sample call); `hist` /
// `cursor` are this batch's device counters.
// Front half: regroup (radix sort or one-pass partition) and start the 64-byte histogram
// read-back.  Nothing here waits for the device, so several GPUs can be fed back to back.
int enqueue_grouped_front(xs_gpu_ctx *ctx, DeviceState &d, int kernel_id, long base, long count,
                          unsigned int *hist, unsigned int *cursor, xs::BatchSink sink, bool record_event)
{
    CUDA_TRY(cudaSetDevice(d.device));
    const uint32_t *id = nullptr;
    bool indirect = false;
    if (kernel_id == 6 && d.bins_ready) {   // (the bins are counted by the event sampler only)
        d.bins_ready = false;
        // optimization 6 (cuda/Simulation.cu:1024-1099): order by (material, energy) -- one-pass bin
        // sort on the fine histogram the sampler counted
        const long n_bins = (long)XS_NUM_MATERIALS << ctx->bin_bits;
        const int n_chunks = (int)((n_bins + xs::kScanChunk - 1) / xs::kScanChunk);
        xs::sort_chunk_sum_kernel<<<n_chunks, xs::kScanThreads, 0, d.stream>>>(d.bin_count, n_bins, d.bin_chunk_sum);
        xs::sort_scan_kernel<<<n_chunks, xs::kScanThreads, 0, d.stream>>>(d.bin_count, n_bins, d.bin_chunk_sum);
        const int blocks = (int)std::min<long>((count + 255) / 256, (long)d.sm_count * 16);
        xs::xs_bin_scatter_kernel<<<blocks, 256, 0, d.stream>>>(d.key[0] + base, d.samp_e + base, d.samp_where + base, count,
                                                               d.bin_count, 28 - ctx->bin_bits, d.grp_e + base,
                                                               d.grp_where + base, d.grp_id + base);
        CUDA_TRY(cudaGetLastError());
        d.launches += 3;
        id = d.grp_id + base;
    } else if (kernel_id == 6) {
        // the same order from a stable three-pass radix sort + gather (XSB200_BIN_BITS=0)
        uint32_t *sorted_perm = nullptr;
        uint32_t *key[2] = { d.key[0] + base, d.key[1] + base }, *perm[2] = { d.perm[0] + base, d.perm[1] + base };
        int rc = sort_lookup_keys(ctx, d, key, perm, count, ctx->key_lo_bit, 32, &sorted_perm, &d.launches);
        if (rc != 0) return set_error(XS_ERR_CUDA, "radix sort failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (ctx->sorted_kernel && ctx->fuse_gather) {
            indirect = true;          // the lane-per-lookup kernel applies the permutation itself
        } else {
            const int blocks = (int)std::min<long>((count + 255) / 256, (long)d.sm_count * 16);
            xs::xs_gather_kernel<<<blocks, 256, 0, d.stream>>>(sorted_perm, d.samp_e + base, d.samp_where + base, count,
                                                              d.grp_e + base, d.grp_where + base);
            CUDA_TRY(cudaGetLastError());
            d.launches++;
        }
        id = sorted_perm;
    } else {
        // optimization 4 (:754-821): group by material; optimization 5 (:895-958): fuel first
        const int tiles = (int)((count + 256 * xs::kPartItems - 1) / (256 * xs::kPartItems));
        xs::xs_partition_kernel<<<tiles, 256, 0, d.stream>>>(d.samp_e + base, d.samp_mat + base, d.samp_where + base, count,
                                                           hist, cursor, kernel_id == 5, d.grp_e + base, d.grp_where + base,
                                                           kernel_id == 5 ? d.grp_mat + base : nullptr, d.grp_id + base);
        CUDA_TRY(cudaGetLastError());
        d.launches++;
        id = d.grp_id + base;
    }
    if (record_event) CUDA_TRY(cudaEventRecord(d.ev[EV_SORTED], d.stream));
    // group sizes: the sampler's histogram (64 bytes device -> pinned host; the window launches
    // take slot ranges as kernel arguments, which measured 5 % faster than deriving them on
    // the device)
    CUDA_TRY(cudaMemcpyAsync(d.h_hist, hist, 16 * sizeof(unsigned int), cudaMemcpyDeviceToHost, d.stream));
    d.pending.active = true;
    d.pending.kernel_id = kernel_id;
    d.pending.base = base;
    d.pending.count = count;
    d.pending.id = id;
    d.pending.indirect = indirect;
    d.pending.sink = sink;
    return XS_OK;
}

// The lookup sort: one launch per digit (xs_sort.cuh, onesweep) or, XSB200_ONESWEEP=0 / very large batches,
// the three-kernel passes of round 1.  Both are stable and give the same permutation.
int sort_lookup_keys(xs_gpu_ctx *ctx, DeviceState &d, uint32_t *key[2], uint32_t *perm[2], long count, int lo_bit, int hi_bit,
                     uint32_t **sorted_perm, int *launches)
{
    // (digits counted by the kernel that wrote exactly these keys, for exactly this sort: see begin_digit_count)
    const bool counted = d.digits_for == count && key[0] == d.key[0] + d.digits_base && lo_bit == ctx->key_lo_bit && hi_bit == 32;
    d.digits_for = -1;
    if (ctx->onesweep) {
        const int r = xs::onesweep_sort(d.sweep, key, perm, count, lo_bit, hi_bit, d.sm_count, d.stream, sorted_perm, launches, counted);
        if (r == 0) return 0;
        if (r != -2) return r;                               // (-2: not applicable -> the three-kernel passes)
    }
    return xs::sort_lookups(d.sort, key, perm, count, lo_bit, hi_bit, 0, d.stream, sorted_perm, launches);
}

// -k 6 without a host round trip: sort, then both lane-per-lookup launches with segment tables a
// one-thread kernel derives from the histogram.  (The windowed sweep takes its slot ranges as kernel
// arguments -- measured 5 % faster there -- so -k 4 / 5 keep the read-back.)
int enqueue_sorted_nosync(xs_gpu_ctx *ctx, DeviceState &d, long base, long count, unsigned int *hist, xs::BatchSink sink,
                          bool record_event, int table_slot)
{
    CUDA_TRY(cudaSetDevice(d.device));
    uint32_t *sorted_perm = nullptr;
    uint32_t *key[2] = { d.key[0] + base, d.key[1] + base }, *perm[2] = { d.perm[0] + base, d.perm[1] + base };
    int rc = sort_lookup_keys(ctx, d, key, perm, count, ctx->key_lo_bit, 32, &sorted_perm, &d.launches);
    if (rc != 0) return set_error(XS_ERR_CUDA, "radix sort failed: %s", cudaGetErrorString(cudaGetLastError()));
    GroupedBatch b{};
    b.indirect = ctx->fuse_gather != 0;
    if (!b.indirect) {
        const int blocks = (int)std::min<long>((count + 255) / 256, (long)d.sm_count * 16);
        xs::xs_gather_kernel<<<blocks, 256, 0, d.stream>>>(sorted_perm, d.samp_e + base, d.samp_where + base, count,
                                                          d.grp_e + base, d.grp_where + base);
        CUDA_TRY(cudaGetLastError());
        d.launches++;
    }
    if (record_event) CUDA_TRY(cudaEventRecord(d.ev[EV_SORTED], d.stream));
    b.pack = b.indirect && ctx->pack_samples ? d.samp_pack + base : nullptr;
    b.energy = (b.indirect ? d.samp_e : d.grp_e) + base;
    b.where = (b.indirect ? d.samp_where : d.grp_where) + base;
    b.id = sorted_perm;
    b.hist = hist;
    b.total = count;
    return launch_sorted(ctx, d, b, sink, d.seg_tables + 2 * table_slot);
}

// Back half: wait for the histogram, then sweep material by material.
int enqueue_grouped_back(xs_gpu_ctx *ctx, DeviceState &d)
{
    if (!d.pending.active) return XS_OK;
    d.pending.active = false;
    const int kernel_id = d.pending.kernel_id;
    const long base = d.pending.base, count = d.pending.count;
    const xs::BatchSink sink = d.pending.sink;
    CUDA_TRY(cudaSetDevice(d.device));
    CUDA_TRY(cudaStreamSynchronize(d.stream));
    GroupedBatch b{};
    b.indirect = d.pending.indirect;
    b.pack = b.indirect && ctx->pack_samples ? d.samp_pack + base : nullptr;
    b.energy = (b.indirect ? d.samp_e : d.grp_e) + base;
    b.where = (b.indirect ? d.samp_where : d.grp_where) + base;
    b.partial = d.sweep_partial + 3 * base;
    b.id = d.pending.id;
    long offset = 0;
    for (int m = 0; m < XS_NUM_MATERIALS; m++) {
        b.offset[m] = offset;
        b.count[m] = d.h_hist[m];
        offset += d.h_hist[m];
    }
    int rc = XS_OK;
    if (kernel_id == 5) {
        // fuel -> windowed sweep; the other 11 materials stay mixed -> one in-order launch
        const int fuel = 0;
        rc = launch_sweep(ctx, d, b, 1, &fuel, sink);
        const long n_fuel = d.h_hist[0];
        if (rc == XS_OK && count > n_fuel) {
            xs::BatchSource rest{};
            rest.energy = d.grp_e + base + n_fuel;
            rest.mat = d.grp_mat + base + n_fuel;
            rest.count = count - n_fuel;
            rest.mat_lo = 0; rest.mat_hi = XS_NUM_MATERIALS - 1;
            rc = launch_event(ctx, d, rest, sink, 1);
        }
        return rc;
    }
    if (kernel_id == 6 && ctx->sorted_kernel) return launch_sorted(ctx, d, b, sink);
    int mats[XS_NUM_MATERIALS];
    for (int m = 0; m < XS_NUM_MATERIALS; m++) mats[m] = m;
    return launch_sweep(ctx, d, b, XS_NUM_MATERIALS, mats, sink);
}

// Sorted variants on device-resident samples (samp_e / samp_mat / samp_where, the material
// histogram and, for -k 6, key[0] are ready in device memory): regroup, then sweep material by
// material.  `base` offsets every per-sample buffer (chunks of a host-sample call); `hist` /
// `cursor` are this batch's device counters.
int enqueue_grouped_lookup(xs_gpu_ctx *ctx, DeviceState &d, int kernel_id, long base, long count,
                           unsigned int *hist, unsigned int *cursor, xs::BatchSink sink, bool record_event)
{
    int rc = enqueue_grouped_front(ctx, d, kernel_id, base, count, hist, cursor, sink, record_event);
    return rc == XS_OK ? enqueue_grouped_back(ctx, d) : rc;
}

// One device's share of an event-mode run: ids [first_id, first_id + count).
int enqueue_event_pass(xs_gpu_ctx *ctx, DeviceState &d, int kernel_id, long first_id, long count, bool first_pass);

// Variants that materialise samples (-k >= 1) work through the id range in passes of at most
// `max_pass` lookups, so the sample / grouping buffers stay bounded (112 B per lookup) however
// many lookups are requested (BASELINE config 5: 1e9+ lookups).  The verification sum simply
// accumulates over the passes.
int enqueue_event_all(xs_gpu_ctx *ctx, int kernel_id, long first_id, long count)
{
    const int n = (int)ctx->dev.size();
    long lo[8], cnt[8], done[8];
    long most = 0;
    for (int g = 0; g < n; g++) {
        DeviceState &d = ctx->dev[g];
        if (ctx->n_bands > 1) {          // band mode: every device draws every id, keeps its band
            lo[g] = first_id;
            cnt[g] = count;
        } else {
            lo[g] = first_id + count * g / n;
            cnt[g] = first_id + count * (g + 1) / n - lo[g];
        }
        done[g] = 0;
        most = std::max(most, cnt[g]);
        CUDA_TRY(cudaSetDevice(d.device));
        d.launches = 0;
        CUDA_TRY(cudaEventRecord(d.ev[EV_START], d.stream));
        CUDA_TRY(cudaMemsetAsync(d.accum, 0, 3 * sizeof(unsigned long long), d.stream));
    }
    const long max_pass = kernel_id == 0 ? std::max<long>(most, 1) : std::max<long>(ctx->max_pass, 1);
    bool more = true, first = true;
    while (more) {
        more = false;
        // front halves on every GPU first (nothing waits), then the back halves: the GPUs of
        // one process start their sweeps together instead of one histogram read-back apart
        for (int g = 0; g < n; g++) {
            const long todo = std::min(cnt[g] - done[g], max_pass);
            if (todo <= 0 && !(first && cnt[g] == 0)) continue;
            CUDA_TRY(cudaSetDevice(ctx->dev[g].device));
            int rc = enqueue_event_pass(ctx, ctx->dev[g], kernel_id, lo[g] + done[g], std::max<long>(todo, 0), first);
            if (rc != XS_OK) return rc;
            done[g] += std::max<long>(todo, 0);
            more |= done[g] < cnt[g];
        }
        for (int g = 0; g < n; g++) {
            int rc = enqueue_grouped_back(ctx, ctx->dev[g]);
            if (rc != XS_OK) return rc;
        }
        first = false;
    }
    for (int g = 0; g < n; g++) {
        CUDA_TRY(cudaSetDevice(ctx->dev[g].device));
        CUDA_TRY(cudaEventRecord(ctx->dev[g].ev[EV_LOOKED_UP], ctx->dev[g].stream));
    }
    return XS_OK;
}

int enqueue_event_pass(xs_gpu_ctx *ctx, DeviceState &d, int kernel_id, long first_id, long count, bool first_pass)
{
    if (count <= 0) {                      // an empty share still needs its phase marks
        if (first_pass) {
            CUDA_TRY(cudaEventRecord(d.ev[EV_SAMPLED], d.stream));
            CUDA_TRY(cudaEventRecord(d.ev[EV_SORTED], d.stream));
        }
        return XS_OK;
    }
    CUDA_TRY(cudaMemsetAsync(d.counters, 0, kNumCounters * sizeof(unsigned int), d.stream));
    CUDA_TRY(cudaMemsetAsync(d.histogram, 0, kNumHist * sizeof(unsigned int), d.stream));

    xs::BatchSink sink{};
    sink.accum = d.accum;
    xs::BatchSource src{};
    src.first_id = first_id;
    src.count = count;
    src.mat_lo = 0;
    src.mat_hi = XS_NUM_MATERIALS - 1;
    int rc = XS_OK;

    if (kernel_id == 0) {
        // baseline semantics (cuda/Simulation.cu:44-99): sample + lookup fused, one launch
        if (first_pass) {
            CUDA_TRY(cudaEventRecord(d.ev[EV_SAMPLED], d.stream));
            CUDA_TRY(cudaEventRecord(d.ev[EV_SORTED], d.stream));
        }
        rc = launch_event(ctx, d, src, sink, 0);
    } else {
        const bool sorted = kernel_id == 4 || kernel_id == 5 || kernel_id == 6;
        if ((rc = ensure_sample_buffers(d, count, sorted, ctx->bin_bits)) != XS_OK) return rc;
        if ((rc = launch_sample(ctx, d, first_id, count, sorted, kernel_id == 6, sorted)) != XS_OK) return rc;
        if (first_pass) CUDA_TRY(cudaEventRecord(d.ev[EV_SAMPLED], d.stream));
        src.energy = d.samp_e;
        src.mat = d.samp_mat;
        if (!sorted) {
            if (first_pass) CUDA_TRY(cudaEventRecord(d.ev[EV_SORTED], d.stream));
            if (kernel_id == 1) {
                // optimization 1 (cuda/Simulation.cu:388-439): split sample / lookup
                rc = launch_event(ctx, d, src, sink, 0);
            } else if (kernel_id == 2) {
                // optimization 2 (:521-574): one launch per material, unsorted samples
                for (int m = 0; m < XS_NUM_MATERIALS && rc == XS_OK; m++) {
                    src.mat_lo = src.mat_hi = m;
                    rc = launch_event(ctx, d, src, sink, m);
                }
            } else {
                // optimization 3 (:637-690): fuel launch + everything-else launch
                src.mat_lo = src.mat_hi = 0;
                rc = launch_event(ctx, d, src, sink, 0);
                src.mat_lo = 1; src.mat_hi = XS_NUM_MATERIALS - 1;
                if (rc == XS_OK) rc = launch_event(ctx, d, src, sink, 1);
            }
        } else if (kernel_id == 6 && ctx->sorted_kernel && ctx->device_segments && !d.bins_ready) {
            rc = enqueue_sorted_nosync(ctx, d, 0, count, d.histogram, sink, first_pass, 0);
        } else {
            rc = enqueue_grouped_front(ctx, d, kernel_id, 0, count, d.histogram, d.counters + kCursorBase, sink, first_pass);
        }
    }
    return rc;
}

// History mode by generations: every generation is an event-style batch (all live particles'
// current lookups); the feedback that makes the reference's inner loop dependent (n_forward) travels
// through one byte per particle.  A generation runs the sorted pipeline without a host round trip
// (xs_history_step_kernel -> radix sort -> segment tables -> lane-per-lookup kernels), or, with
// XSB200_SORTED_KERNEL=0 / XSB200_DEVICE_SEGMENTS=0, round 1's partition + windowed sweep.  Several GPUs
// of one process are fed generation by generation, device by device, so they run side by side.
// Energy-band sharding: every device steps every particle, looks up the ones whose row lies in its band
// and the per-particle feedback bytes are summed over the devices (NCCL all-reduce, uint8) before the
// next generation -- the one place where this path has a collective in the data path.
int xs_multi_allreduce_bytes(xs_gpu_ctx *ctx, size_t n_bytes);

int enqueue_history_all(xs_gpu_ctx *ctx, long first_particle, long n_particles, int lookups)
{
    const int n = (int)ctx->dev.size();
    const bool bands = ctx->n_bands > 1;
    if (bands && n != ctx->n_bands)
        return set_error(XS_ERR_UNSUPP, "xs_gpu_run: history mode on an energy-band-sharded grid needs all bands in one context "
                                        "(the bands exchange the particles' feedback every generation)");
    long lo[8], cnt[8];
    for (int g = 0; g < n; g++) {
        DeviceState &d = ctx->dev[g];
        lo[g] = bands ? first_particle : first_particle + n_particles * g / n;
        cnt[g] = bands ? n_particles : first_particle + n_particles * (g + 1) / n - lo[g];
        CUDA_TRY(cudaSetDevice(d.device));
        d.launches = 0;
        CUDA_TRY(cudaEventRecord(d.ev[EV_START], d.stream));
        CUDA_TRY(cudaMemsetAsync(d.accum, 0, 3 * sizeof(unsigned long long), d.stream));
        CUDA_TRY(cudaMemsetAsync(d.counters, 0, kNumCounters * sizeof(unsigned int), d.stream));
        CUDA_TRY(cudaEventRecord(d.ev[EV_SAMPLED], d.stream));
        CUDA_TRY(cudaEventRecord(d.ev[EV_SORTED], d.stream));
    }
    if (!ctx->sweep) {
        // warp per particle, all lookups of a particle in one go (round 1's first version; kept for comparison)
        for (int g = 0; g < n; g++) {
            DeviceState &d = ctx->dev[g];
            if (cnt[g] <= 0) continue;
            CUDA_TRY(cudaSetDevice(d.device));
            HistoryKernel k = history_kernel(ctx->grid_type, ctx->gather);
            int blocks = 0;
            int rc = persistent_grid(ctx, d, (const void *)k, &blocks);
            if (rc != XS_OK) return rc;
            const long max_useful = (cnt[g] + xs::kWarpsPerBlock - 1) / xs::kWarpsPerBlock;
            if (blocks > max_useful) blocks = (int)max_useful;
            xs::BatchSink sink{};
            sink.accum = d.accum;
            sink.batch_counter = d.counters;
            k<<<blocks, xs::kBlockThreads, ctx->smem_bytes, d.stream>>>(d.P, lo[g], cnt[g], lookups, sink);
            CUDA_TRY(cudaGetLastError());
            d.launches++;
        }
    } else {
        long most = 0;
        for (int g = 0; g < n; g++) most = std::max(most, cnt[g]);
        // The sorted pipeline pays when a generation is dense enough for xs_dense_kernel (fuel, 13.9 % of the
        // lookups, at >= dense_min lookups per grid interval: 5.2 M particles at "large"); below that the
        // partition + windowed sweep is faster (500 k particles x 34: 14.3 vs 18.8 ms).  Band sharding needs
        // the sorted pipeline (it is the one that drops the other bands' lookups).
        const bool dense_enough = ctx->dense_min > 0 && 0.139 * (double)std::min(most, ctx->max_pass) >= (double)ctx->dense_min * (double)ctx->n_gp;
        const bool nosync = ctx->sorted_kernel && ctx->device_segments && (bands || dense_enough || env_int("XSB200_HISTORY_SORTED", 0));
        if (bands && !nosync)
            return set_error(XS_ERR_UNSUPP, "xs_gpu_run: history mode on an energy-band-sharded grid needs the sorted pipeline (default knobs)");
        for (long done = 0; done < most; done += ctx->max_pass) {
            long pass[8];
            for (int g = 0; g < n; g++) {
                pass[g] = std::max<long>(0, std::min(cnt[g] - done, ctx->max_pass));
                if (pass[g] > 0) { int rc = ensure_sample_buffers(ctx->dev[g], pass[g], true); if (rc != XS_OK) return rc; }
            }
            for (int gen = 0; gen < lookups; gen++) {
                for (int g = 0; g < n; g++) {
                    DeviceState &d = ctx->dev[g];
                    const long np = pass[g];
                    if (np <= 0) continue;
                    CUDA_TRY(cudaSetDevice(d.device));
                    CUDA_TRY(cudaMemsetAsync(d.counters, 0, kNumCounters * sizeof(unsigned int), d.stream));
                    CUDA_TRY(cudaMemsetAsync(d.histogram, 0, kNumHist * sizeof(unsigned int), d.stream));
                    const int blocks = (int)std::min<long>((np + 255) / 256, (long)d.sm_count * 16);
                    d.digits_for = -1;                        // (this kernel does not count the sort's digits)
                    xs::xs_history_step_kernel<<<blocks, 256, 0, d.stream>>>(d.P, ctx->grid_type, lo[g] + done, np, lookups, gen,
                                                                           d.hist_seed, d.hist_fwd, d.samp_e, d.samp_mat, d.samp_where,
                                                                           d.histogram, nosync ? d.key[0] : nullptr,
                                                                           nosync && ctx->pack_samples ? d.samp_pack : nullptr,
                                                                           bands ? (uint32_t)d.row0 : 0u, bands ? (uint32_t)d.row1 : 0xffffffffu);
                    CUDA_TRY(cudaGetLastError());
                    d.launches++;
                    if (bands) CUDA_TRY(cudaMemsetAsync(d.hist_fwd, 0, (size_t)np, d.stream));   // each band fills in its own particles
                    xs::BatchSink sink{};
                    sink.accum = d.accum;
                    sink.fwd_out = d.hist_fwd;
                    int rc = nosync ? enqueue_sorted_nosync(ctx, d, 0, np, d.histogram, sink, false, 0)
                                    : enqueue_grouped_lookup(ctx, d, 4, 0, np, d.histogram, d.counters + kCursorBase, sink, false);
                    if (rc != XS_OK) return rc;
                }
                if (bands && gen + 1 < lookups) {
                    int rc = xs_multi_allreduce_bytes(ctx, (size_t)pass[0]);
                    if (rc != XS_OK) return rc;
                }
            }
        }
    }
    for (int g = 0; g < n; g++) {
        CUDA_TRY(cudaSetDevice(ctx->dev[g].device));
        CUDA_TRY(cudaEventRecord(ctx->dev[g].ev[EV_LOOKED_UP], ctx->dev[g].stream));
    }
    return XS_OK;
}

// Collect {verification, n_lookups}: all-reduce across devices (NCCL) when n_gpus > 1, then
// device -> pinned host, and fill the timing fields.
int finish_run(xs_gpu_ctx *ctx, xs_gpu_result *res, double host_t0)
{
    const int n = (int)ctx->dev.size();
    if (n > 1) {
        int rc = xs_multi_allreduce(ctx);
        if (rc != XS_OK) return rc;
    }
    for (int g = 0; g < n; g++) {
        DeviceState &d = ctx->dev[g];
        CUDA_TRY(cudaSetDevice(d.device));
        CUDA_TRY(cudaMemcpyAsync(d.h_accum, d.accum, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, d.stream));
        CUDA_TRY(cudaEventRecord(d.ev[EV_DONE], d.stream));
    }
    memset(res, 0, sizeof *res);
    for (int g = 0; g < n; g++) {
        DeviceState &d = ctx->dev[g];
        CUDA_TRY(cudaSetDevice(d.device));
        CUDA_TRY(cudaStreamSynchronize(d.stream));
        float ms[4], total = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&total, d.ev[EV_START], d.ev[EV_DONE]));
        for (int p = 0; p < 4; p++) CUDA_TRY(cudaEventElapsedTime(&ms[p], d.ev[p], d.ev[p + 1]));
        res->device_seconds = std::max(res->device_seconds, (double)total * 1e-3);
        for (int p = 0; p < 4; p++) res->phase_seconds[p] = std::max(res->phase_seconds[p], (double)ms[p] * 1e-3);
        res->gpu_launches += d.launches;
        res->d2h_bytes += 2 * sizeof(unsigned long long);
    }
    CUDA_TRY(cudaSetDevice(ctx->dev[0].device));
    res->verification = ctx->dev[0].h_accum[0];     // all-reduced: identical on every device
    res->n_lookups = ctx->dev[0].h_accum[1];
    res->host_seconds = wall_seconds() - host_t0;
    res->n_gpus = n;
    return XS_OK;
}

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

const char *xs_gpu_last_error(void) { return g_error; }
const char *xs_gpu_version(void) { return "xsbench_b200 0.1 (sm_100a)"; }

int xs_gpu_init(const Inputs *in, const SimulationData *sd, int n_gpus, xs_gpu_ctx **out)
{
    if (!in || !sd || !out) return set_error(XS_ERR_ARG, "xs_gpu_init: NULL argument");
    *out = nullptr;
    if (n_gpus < 1 || n_gpus > 8) return set_error(XS_ERR_ARG, "xs_gpu_init: n_gpus must be 1..8");
    if (in->grid_type < 0 || in->grid_type > 2) return set_error(XS_ERR_ARG, "xs_gpu_init: bad grid_type %d", in->grid_type);
    if (in->n_isotopes < 1 || in->n_gridpoints < 2) return set_error(XS_ERR_ARG, "xs_gpu_init: bad problem size");
    if (!sd->num_nucs || !sd->mats || !sd->concs)
        return set_error(XS_ERR_ARG, "xs_gpu_init: SimulationData has NULL material arrays");
    // All three big arrays NULL = build the problem on the device (xs_generate.cuh).
    const bool generate = !sd->nuclide_grid && !sd->unionized_energy_array && !sd->index_grid;
    if (!generate) {
        if (!sd->nuclide_grid) return set_error(XS_ERR_ARG, "xs_gpu_init: nuclide_grid is NULL");
        if (sd->length_nuclide_grid != in->n_isotopes * in->n_gridpoints)
            return set_error(XS_ERR_ARG, "xs_gpu_init: length_nuclide_grid does not match Inputs");
        if (in->grid_type == XS_UNIONIZED &&
            (!sd->unionized_energy_array || !sd->index_grid ||
             sd->length_index_grid != (long)sd->length_unionized_energy_array * in->n_isotopes))
            return set_error(XS_ERR_ARG, "xs_gpu_init: unionized grid arrays missing or inconsistent");
        if (in->grid_type == XS_HASH &&
            (!sd->index_grid || in->hash_bins < 1 || sd->length_index_grid != (long)in->hash_bins * in->n_isotopes))
            return set_error(XS_ERR_ARG, "xs_gpu_init: hash grid missing or inconsistent");
    } else if (in->grid_type == XS_HASH && in->hash_bins < 1) {
        return set_error(XS_ERR_ARG, "xs_gpu_init: hash_bins must be >= 1");
    }

    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev < 1)
        return set_error(XS_ERR_CUDA, "xs_gpu_init: no CUDA device (%s); this library has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    int dev0 = 0;
    CUDA_TRY(cudaGetDevice(&dev0));
    if (dev0 + n_gpus > n_dev)
        return set_error(XS_ERR_ARG, "xs_gpu_init: %d GPUs requested from device %d but only %d present", n_gpus, dev0, n_dev);

    xs_gpu_ctx *ctx = new (std::nothrow) xs_gpu_ctx;
    if (!ctx) return set_error(XS_ERR_ARG, "out of host memory");
    ctx->grid_type = in->grid_type;
    ctx->n_iso = in->n_isotopes;
    ctx->n_gp = in->n_gridpoints;
    ctx->hash_bins = in->hash_bins;
    ctx->max_num_nucs = sd->max_num_nucs;
    ctx->n_ueg = in->grid_type == XS_UNIONIZED ? in->n_isotopes * in->n_gridpoints : 0;
    ctx->gather = env_int("XSB200_GATHER", xs::kTriple) ? xs::kTriple : xs::kLanePerNuclide;
    ctx->blocks_per_sm = env_int("XSB200_BLOCKS_PER_SM", 0);
    ctx->sweep = env_int("XSB200_SWEEP", 1);
    ctx->max_pass = std::max(1024, env_int("XSB200_MAX_PASS", 1 << 26));
    ctx->e2e_chunks = std::min<int>(kMaxChunks, std::max(0, env_int("XSB200_E2E_CHUNKS", 0)));
    {
        // host threads that narrow the materials of a host-sample call: this process's share of the cores, between
        // 2 and 8.  The processes sharing the host are counted from the launcher's environment (torchrun / Open MPI /
        // MPICH / Slurm; one process otherwise) times the GPUs of this context.  With fewer than 8 cores' worth per
        // GPU the narrowing is slower than the copy it shortens -- measured on a 32-core host: 4 ranks 7.8 ms per call
        // with, 5.6 without; 8 ranks 13.0 / 10.7 -- and the materials travel as the caller's ints.
        int ranks = 1;
        for (const char *name : { "LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "MPI_LOCALNRANKS", "SLURM_NTASKS_PER_NODE" }) {
            const int v = env_int(name, 0);
            if (v > 0) { ranks = v; break; }
        }
        const int cores = (int)std::thread::hardware_concurrency();
        const int share = cores / (2 * ranks * std::max(1, n_gpus));
        const int dflt = std::min(8, std::max(2, share));
        ctx->pack_threads = std::min(64, std::max(1, env_int("XSB200_PACK_THREADS", dflt)));
        ctx->host_pack = env_int("XSB200_HOST_PACK", share >= 8 ? 1 : 0);
    }
    ctx->window = std::max(1, env_int("XSB200_WINDOW", 32));
    ctx->sorted_kernel = env_int("XSB200_SORTED_KERNEL", 1);
    ctx->e2e_kernel = env_int("XSB200_E2E_KERNEL", 6) == 4 ? 4 : 6;
    ctx->fuse_gather = env_int("XSB200_FUSE_GATHER", 1);
    ctx->pack_samples = env_int("XSB200_PACK_SAMPLES", 1);
    ctx->bin_bits = std::min(20, std::max(0, env_int("XSB200_BIN_BITS", 0)));
    ctx->dense_min = std::max(0, env_int("XSB200_DENSE_MIN", 64));
    ctx->device_segments = env_int("XSB200_DEVICE_SEGMENTS", 1);
    ctx->onesweep = env_int("XSB200_ONESWEEP", 1);
    ctx->fuse_digits = env_int("XSB200_FUSE_DIGITS", 1);
    ctx->tile = env_int("XSB200_TILE", 1);
    ctx->tile_barrier = env_int("XSB200_TILE_BARRIER", -1);
    if (const char *split = getenv("XSB200_E2E_SPLIT")) {
        int n = 0;
        for (const char *p = split; *p && n < kMaxChunks; ) {
            ctx->e2e_split[n++] = atoi(p);
            while (*p >= '0' && *p <= '9') p++;
            while (*p && !(*p >= '0' && *p <= '9')) p++;       // any separator
        }
    }
    {
        const char *arith = getenv("XSB200_ARITH");
        ctx->exact_arith = !(arith && (!strcmp(arith, "fused") || !strcmp(arith, "12")));
    }
    ctx->key_lo_bit = std::min(28, std::max(0, env_int("XSB200_KEY_LO_BIT", 8)));
    for (int m = 0; m < XS_NUM_MATERIALS; m++) ctx->num_nucs[m] = sd->num_nucs[m];
    ctx->dev.resize(n_gpus);
    for (int g = 0; g < n_gpus; g++) ctx->dev[g].device = dev0 + g;

    // Energy-band sharding of the unionized index grid (SURVEY 8e option 2).  XSB200_BANDS=N forces
    // it; by default it switches on when several GPUs are given and one GPU cannot hold the grid
    // (XXL: 253 GB of index rows).  A single-GPU context can hold one band (XSB200_BAND_INDEX): the
    // caller then sums the results of the N contexts, e.g. one rank per GPU.
    if (in->grid_type == XS_UNIONIZED) {
        const int forced = env_int("XSB200_BANDS", 0);
        if (forced > 1) {
            ctx->n_bands = forced;
        } else if (forced == 0 && n_gpus > 1) {
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            const double points = (double)in->n_isotopes * (double)in->n_gridpoints;
            const double need = points * (double)in->n_isotopes * 4.0 + points * (48.0 + 128.0 + 8.0 + 40.0);
            if (need > 0.92 * (double)total_b) ctx->n_bands = n_gpus;
        }
        if (ctx->n_bands > 1) {
            const int index = env_int("XSB200_BAND_INDEX", -1);
            if (n_gpus > 1 && ctx->n_bands != n_gpus) {
                delete ctx;
                return set_error(XS_ERR_ARG, "xs_gpu_init: %d energy bands need %d GPUs (got %d)", forced, forced, n_gpus);
            }
            if (n_gpus == 1 && (index < 0 || index >= ctx->n_bands)) {
                delete ctx;
                return set_error(XS_ERR_ARG, "xs_gpu_init: XSB200_BANDS=%d on one GPU needs XSB200_BAND_INDEX in [0, %d)", forced, forced);
            }
            for (int g = 0; g < n_gpus; g++) ctx->dev[g].band = n_gpus > 1 ? g : index;
            ctx->bin_bits = 0;            // (the bin sort has no bin for dropped lookups)
        }
    }

    int rc = XS_OK;
    for (int g = 0; g < n_gpus && rc == XS_OK; g++) {
        if (g > 0) {
            // replicate from GPU 0 over NVLink (peer copy), excluded from the FOM like the
            // reference's H2D (cuda/Main.cu:42 is before the timer)
            int can = 0;
            cudaDeviceCanAccessPeer(&can, ctx->dev[g].device, ctx->dev[0].device);
            if (can) {
                cudaSetDevice(ctx->dev[g].device);
                cudaDeviceEnablePeerAccess(ctx->dev[0].device, 0);
                cudaGetLastError();
            }
        }
        rc = upload_device(ctx, ctx->dev[g], in, sd, g > 0 && ctx->n_bands == 1 ? &ctx->dev[0] : nullptr);
    }
    if (rc == XS_OK && in->simulation_method == XS_EVENT_BASED && in->kernel_id != 0 && in->lookups > 0) {
        // pre-allocate the sample / sort buffers here, not inside the timed region
        const long per_gpu = ctx->n_bands > 1 ? (long)in->lookups : ((long)in->lookups + n_gpus - 1) / n_gpus;
        for (int g = 0; g < n_gpus && rc == XS_OK; g++)
            rc = ensure_sample_buffers(ctx->dev[g], std::min(per_gpu, ctx->max_pass), in->kernel_id >= 4 || ctx->n_bands > 1,
                                       ctx->bin_bits);
    }
    if (rc == XS_OK && in->simulation_method == XS_HISTORY_BASED && in->particles > 0 && ctx->sweep) {
        // history mode works on one generation (= all particles) at a time
        const long per_gpu = ctx->n_bands > 1 ? (long)in->particles : ((long)in->particles + n_gpus - 1) / n_gpus;
        for (int g = 0; g < n_gpus && rc == XS_OK; g++) rc = ensure_sample_buffers(ctx->dev[g], per_gpu, true);
    }
    if (rc == XS_OK && n_gpus > 1) rc = xs_multi_init(ctx);
    if (rc != XS_OK) {
        char keep[sizeof g_error];
        memcpy(keep, g_error, sizeof keep);
        xs_gpu_finalize(ctx);
        memcpy(g_error, keep, sizeof keep);
        return rc;
    }
    cudaSetDevice(dev0);
    *out = ctx;
    return XS_OK;
}

int xs_gpu_run_range(xs_gpu_ctx *ctx, const Inputs *in, long first_id, long count, xs_gpu_result *res)
{
    DeviceGuard restore_device;
    if (!ctx || !in || !res) return set_error(XS_ERR_ARG, "xs_gpu_run: NULL argument");
    if (count < 0 || first_id < 0) return set_error(XS_ERR_ARG, "xs_gpu_run: negative range");
    const bool event = in->simulation_method == XS_EVENT_BASED;
    if (!event && in->simulation_method != XS_HISTORY_BASED)
        return set_error(XS_ERR_ARG, "xs_gpu_run: unknown simulation_method %d", in->simulation_method);
    if (event && (in->kernel_id < 0 || in->kernel_id > 6))
        return set_error(XS_ERR_ARG, "xs_gpu_run: no kernel ID %d", in->kernel_id);
    if (!event && in->kernel_id != 0)
        return set_error(XS_ERR_ARG, "xs_gpu_run: no kernel ID %d for history mode", in->kernel_id);
    if (!event && in->lookups < 1) return set_error(XS_ERR_ARG, "xs_gpu_run: lookups per particle must be >= 1");
    if (in->grid_type != ctx->grid_type || in->n_isotopes != ctx->n_iso || in->n_gridpoints != ctx->n_gp)
        return set_error(XS_ERR_ARG, "xs_gpu_run: Inputs do not match the problem uploaded by xs_gpu_init");

    const double t0 = wall_seconds();
    const int n = (int)ctx->dev.size();
    if (event) {
        // band mode: every variant runs the sorted pipeline (the only one that can drop the
        // lookups of other bands)
        int rc = enqueue_event_all(ctx, ctx->n_bands > 1 ? 6 : in->kernel_id, first_id, count);
        if (rc != XS_OK) return rc;
    } else {
        int rc = enqueue_history_all(ctx, first_id, count, in->lookups);
        if (rc != XS_OK) return rc;
    }
    return finish_run(ctx, res, t0);
}

int xs_gpu_run(xs_gpu_ctx *ctx, const Inputs *in, xs_gpu_result *res)
{
    if (!in) return set_error(XS_ERR_ARG, "xs_gpu_run: NULL argument");
    const long count = in->simulation_method == XS_EVENT_BASED ? (long)in->lookups : (long)in->particles;
    return xs_gpu_run_range(ctx, in, 0, count, res);
}

int xs_gpu_lookup_samples(xs_gpu_ctx *ctx, const double *h_energy, const int *h_mat, long n,
                          double *h_macro_xs_out, xs_gpu_result *res)
{
    DeviceGuard restore_device;
    if (!ctx || !h_energy || !h_mat || !res || n < 0)
        return set_error(XS_ERR_ARG, "xs_gpu_lookup_samples: bad argument");
    // Energy-band sharding: every device takes ALL samples and looks up those whose unionized row lies in
    // its band (a single-GPU context holding one band returns that band's share: the caller adds the
    // contexts up, like for xs_gpu_run; macro_xs rows of other bands' lookups stay zero).
    const bool bands = ctx->n_bands > 1;
    if (bands && !(ctx->sweep && ctx->e2e_kernel == 6 && ctx->sorted_kernel && ctx->device_segments))
        return set_error(XS_ERR_UNSUPP, "xs_gpu_lookup_samples: an energy-band-sharded grid needs the sorted pipeline (default knobs)");
    const double t0 = wall_seconds();
    const int ng = (int)ctx->dev.size();
    unsigned long long mat_bytes_moved = 0;                  // what the materials cost on PCIe: 4 bytes each, or 1 narrowed (xs_hostpack.h)
    for (int g = 0; g < ng; g++) {
        DeviceState &d = ctx->dev[g];
        const long lo = bands ? 0 : n * g / ng, cnt = bands ? n : n * (g + 1) / ng - lo;
        mat_bytes_moved += (unsigned long long)cnt * sizeof(int);
        const uint32_t band_lo = bands ? (uint32_t)d.row0 : 0u, band_hi = bands ? (uint32_t)d.row1 : 0xffffffffu;
        int rc = ensure_sample_buffers(d, cnt, ctx->sweep != 0);
        if (rc == XS_OK && h_macro_xs_out) rc = ensure_dump_buffer(d, cnt);
        if (rc != XS_OK) return rc;
        CUDA_TRY(cudaSetDevice(d.device));
        d.launches = 0;
        CUDA_TRY(cudaEventRecord(d.ev[EV_START], d.stream));
        CUDA_TRY(cudaMemsetAsync(d.accum, 0, 3 * sizeof(unsigned long long), d.stream));
        CUDA_TRY(cudaMemsetAsync(d.counters, 0, kNumCounters * sizeof(unsigned int), d.stream));
        CUDA_TRY(cudaMemsetAsync(d.histogram, 0, kNumHist * sizeof(unsigned int), d.stream));
        xs::BatchSink sink{};
        sink.accum = d.accum;
        if (bands && h_macro_xs_out && cnt > 0)
            CUDA_TRY(cudaMemsetAsync(d.dump_macro, 0, (size_t)cnt * 5 * sizeof(double), d.stream));
        if (ctx->sweep && cnt > 0) {
            // Pipelined in chunks: the host->device copy of chunk c+1 (copy stream) overlaps the
            // row search, sort and lookup kernels of chunk c (compute stream).  Every chunk is a complete
            // -k 6 (or, XSB200_E2E_KERNEL=4, -k 4) pipeline on its own slice of the buffers.
            // Chunk schedule.  The call costs the copy of all samples plus whatever compute is not hidden
            // behind it: the first chunk's copy has nothing to overlap with, the last chunk's compute nothing
            // left to hide behind.  Many small chunks would shrink both ends -- but a chunk is a complete sorted
            // pipeline, and what makes that pipeline fast is the DENSITY of the sorted lookups (lookups per grid
            // interval: with a third of the samples fuel drops under the dense kernel's threshold and the lookup
            // phase costs 1.5 x as much per lookup).  Measured on B200, 17 M lookups, 204 MB (profiles/r02_notes.md):
            // 1 chunk 6.76 ms, 2 equal chunks 5.50, 3 chunks (28:40:32) 5.53, 4 chunks 5.9, 6 chunks 6.6, 8 chunks 7.2;
            // the copy alone is 3.9 ms.  Nothing on the host waits between chunks (segment tables are built on
            // the device).
            static const int schedule[kMaxChunks + 1][kMaxChunks] = {
                {}, {100}, {50, 50}, {28, 40, 32}, {22, 36, 27, 15}, {18, 27, 27, 17, 11}, {15, 25, 25, 17, 11, 7},
                {13, 21, 23, 18, 12, 8, 5}, {11, 18, 21, 18, 13, 9, 6, 4} };
            int n_chunks = ctx->e2e_chunks ? ctx->e2e_chunks : cnt >= 4000000 ? 2 : 1;
            if (ctx->e2e_split[0] > 0) { n_chunks = 0; while (n_chunks < kMaxChunks && ctx->e2e_split[n_chunks] > 0) n_chunks++; }
            const int *weights = ctx->e2e_split[0] > 0 ? ctx->e2e_split : schedule[n_chunks];
            long bound[kMaxChunks + 1] = { 0 };
            {
                long total_w = 0, run = 0;
                for (int c = 0; c < n_chunks; c++) total_w += weights[c];
                for (int c = 0; c < n_chunks; c++) { run += weights[c]; bound[c + 1] = c + 1 == n_chunks ? cnt : (cnt * run / total_w) & ~4095L; }    // (copies start on 4 KB boundaries, also as bytes)
            }
            const bool nosync = ctx->e2e_kernel == 6 && ctx->sorted_kernel && ctx->device_segments;
            // Materials as bytes (xs_hostpack.h): only where nothing downstream reads them as ints -- the sorted
            // pipeline on packed samples, whose key carries the material.
            const bool pack_mats = ctx->host_pack && nosync && ctx->pack_samples && ctx->fuse_gather;
            if (pack_mats) {
                mat_bytes_moved -= (unsigned long long)cnt * (sizeof(int) - 1);
                if (cnt > d.h_mat8_capacity) {
                    if (d.h_mat8) cudaFreeHost(d.h_mat8);
                    d.h_mat8 = nullptr; d.h_mat8_capacity = 0;
                    CUDA_TRY(cudaHostAlloc(&d.h_mat8, (size_t)cnt, cudaHostAllocDefault));
                    d.h_mat8_capacity = cnt;
                }
                if (!ctx->pack_pool) ctx->pack_pool = new (std::nothrow) xs::PackPool(ctx->pack_threads - 1);
                if (!ctx->pack_pool) return set_error(XS_ERR_ARG, "out of host memory");
            }
            uint8_t *d_mat8 = reinterpret_cast<uint8_t *>(d.samp_mat);       // (the byte view of the same device buffer)
            CUDA_TRY(cudaEventRecord(d.ev_ready, d.stream));
            CUDA_TRY(cudaStreamWaitEvent(d.copy_stream, d.ev_ready, 0));
            // A chunk's copies (one copy stream), then its kernels.  Narrowed materials: the host narrows chunk c
            // while the DMA engine moves chunk c's energies, and the next chunk's energies are enqueued before this
            // chunk's kernels so the engine does not idle meanwhile.  (Tried and slower, profiles/r02_notes.md: the
            // materials on a stream of their own -- copies run in submission order whatever their stream --, and all
            // materials narrowed up front in one copy.)
            auto copy_energies = [&](int c) {
                cudaError_t e = cudaMemcpyAsync(d.samp_e + bound[c], h_energy + lo + bound[c], (size_t)(bound[c + 1] - bound[c]) * sizeof(double),
                                                cudaMemcpyHostToDevice, d.copy_stream);
                if (e == cudaSuccess) e = cudaEventRecord(d.ev_ecopy[c], d.copy_stream);
                return e;
            };
            double pack_ms[kMaxChunks] = {};                                   // (host timeline, for XSB200_E2E_TRACE)
            CUDA_TRY(copy_energies(0));
            for (int c = 0; c < n_chunks && rc == XS_OK; c++) {
                const long c_lo = bound[c], c_n = bound[c + 1] - c_lo;
                if (pack_mats) {
                    const double tp = wall_seconds();
                    ctx->pack_pool->run(h_mat + lo + c_lo, d.h_mat8 + c_lo, c_n);
                    pack_ms[c] = 1e3 * (wall_seconds() - tp);
                    CUDA_TRY(cudaMemcpyAsync(d_mat8 + c_lo, d.h_mat8 + c_lo, (size_t)c_n, cudaMemcpyHostToDevice, d.copy_stream));
                } else {
                    CUDA_TRY(cudaMemcpyAsync(d.samp_mat + c_lo, h_mat + lo + c_lo, (size_t)c_n * sizeof(int), cudaMemcpyHostToDevice, d.copy_stream));
                }
                CUDA_TRY(cudaEventRecord(d.ev_copy[c], d.copy_stream));
                if (c + 1 < n_chunks) CUDA_TRY(copy_energies(c + 1));
                CUDA_TRY(cudaStreamWaitEvent(d.stream, d.ev_copy[c], 0));
                if (c_n <= 0) continue;
                const int blocks = (int)std::min<long>((c_n + 255) / 256, (long)d.sm_count * 16);
                const xs::DigitSpec digits = ctx->e2e_kernel == 6 && nosync ? begin_digit_count(ctx, d, c_lo, c_n) : xs::DigitSpec{};
                xs::xs_locate_kernel<<<blocks, 256, 0, d.stream>>>(d.P, ctx->grid_type, c_n, d.samp_e + c_lo, d.samp_mat + c_lo,
                                                                 pack_mats ? d_mat8 + c_lo : nullptr,
                                                                 pack_mats ? nullptr : d.samp_where + c_lo, ctx->e2e_kernel == 6 ? d.key[0] + c_lo : nullptr,
                                                                 d.histogram + 16 * c,
                                                                 ctx->e2e_kernel == 6 && ctx->pack_samples ? d.samp_pack + c_lo : nullptr,
                                                                 d.accum + 2, band_lo, band_hi, digits);
                CUDA_TRY(cudaGetLastError());
                d.launches++;
                if (c == 0) CUDA_TRY(cudaEventRecord(d.ev[EV_SAMPLED], d.stream));
                xs::BatchSink chunk_sink = sink;
                chunk_sink.macro_out = h_macro_xs_out ? d.dump_macro + 5 * c_lo : nullptr;
                if (nosync)
                    rc = enqueue_sorted_nosync(ctx, d, c_lo, c_n, d.histogram + 16 * c, chunk_sink, c == 0, c);
                else
                    rc = enqueue_grouped_lookup(ctx, d, ctx->e2e_kernel, c_lo, c_n, d.histogram + 16 * c, d.counters + kCursorBase + 16 * c,
                                                chunk_sink, c == 0);
                CUDA_TRY(cudaEventRecord(d.ev_chunk[c], d.stream));
            }
            if (env_int("XSB200_E2E_TRACE", 0) && rc == XS_OK) {
                CUDA_TRY(cudaStreamSynchronize(d.stream));
                for (int c = 0; c < n_chunks; c++) {
                    float t_copy = 0.f, t_done = 0.f, t_e = 0.f;
                    cudaEventElapsedTime(&t_e, d.ev[EV_START], d.ev_ecopy[c]);
                    fprintf(stderr, "[e2e trace] energies of chunk %d in at %.3f ms\n", c, t_e);
                    cudaEventElapsedTime(&t_copy, d.ev[EV_START], d.ev_copy[c]);
                    cudaEventElapsedTime(&t_done, d.ev[EV_START], d.ev_chunk[c]);
                    fprintf(stderr, "[e2e trace] gpu %d chunk %d: %ld samples, copy done %.3f ms, compute done %.3f ms; host: materials narrowed in %.3f ms "
                            "(%d threads)\n", g, c, bound[c + 1] - bound[c], t_copy, t_done, pack_ms[c], pack_mats ? ctx->pack_threads : 0);
                }
            }
        } else {
            CUDA_TRY(cudaMemcpyAsync(d.samp_e, h_energy + lo, (size_t)cnt * sizeof(double), cudaMemcpyHostToDevice, d.stream));
            CUDA_TRY(cudaMemcpyAsync(d.samp_mat, h_mat + lo, (size_t)cnt * sizeof(int), cudaMemcpyHostToDevice, d.stream));
            if (cnt > 0) {
                const int vblocks = (int)std::min<long>((cnt + 255) / 256, (long)d.sm_count * 16);
                xs::xs_validate_samples_kernel<<<vblocks, 256, 0, d.stream>>>(cnt, d.samp_e, d.samp_mat, d.accum + 2);
                CUDA_TRY(cudaGetLastError());
                d.launches++;
            }
            CUDA_TRY(cudaEventRecord(d.ev[EV_SAMPLED], d.stream));
            CUDA_TRY(cudaEventRecord(d.ev[EV_SORTED], d.stream));
            sink.macro_out = h_macro_xs_out ? d.dump_macro : nullptr;
            xs::BatchSource src{};
            src.energy = d.samp_e; src.mat = d.samp_mat; src.count = cnt;
            src.mat_lo = 0; src.mat_hi = XS_NUM_MATERIALS - 1;
            rc = launch_event(ctx, d, src, sink, 0);
        }
        if (rc != XS_OK) return rc;
        if (h_macro_xs_out && (!bands || g == 0))
            CUDA_TRY(cudaMemcpyAsync(h_macro_xs_out + 5 * lo, d.dump_macro, (size_t)cnt * 5 * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
        CUDA_TRY(cudaEventRecord(d.ev[EV_LOOKED_UP], d.stream));
    }
    int rc = finish_run(ctx, res, t0);
    if (rc != XS_OK) return rc;
    if (bands && h_macro_xs_out && ng > 1 && n > 0) {
        // every lookup was performed by exactly one band, the others left zeros: add the bands up (exact)
        std::vector<double> part((size_t)n * 5);
        for (int g = 1; g < ng; g++) {
            CUDA_TRY(cudaSetDevice(ctx->dev[g].device));
            CUDA_TRY(cudaMemcpy(part.data(), ctx->dev[g].dump_macro, part.size() * sizeof(double), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < part.size(); i++) h_macro_xs_out[i] += part[i];
        }
    }
    unsigned long long rejected = 0;
    for (int g = 0; g < ng; g++) rejected += ctx->dev[g].h_accum[2];
    if (rejected)
        return set_error(XS_ERR_ARG, "xs_gpu_lookup_samples: %llu sample(s) with a material outside [0, %d) or an energy outside [0, 1]",
                         rejected, XS_NUM_MATERIALS);
    res->h2d_bytes = (unsigned long long)n * sizeof(double) * (bands ? (unsigned long long)ng : 1ULL) + mat_bytes_moved;
    if (h_macro_xs_out) res->d2h_bytes += (unsigned long long)n * 5 * sizeof(double) * (bands ? (unsigned long long)ng : 1ULL);
    return XS_OK;
}

int xs_gpu_dump(xs_gpu_ctx *ctx, long first_id, long n, double *h_energy_out, int *h_mat_out,
                double *h_macro_xs_out, int *h_argmax_out)
{
    DeviceGuard restore_device;
    if (!ctx || n < 0 || first_id < 0) return set_error(XS_ERR_ARG, "xs_gpu_dump: bad argument");
    if (n == 0) return XS_OK;
    // (an energy-band-sharded grid: GPU 0 performs the lookups of ITS band; energies and materials are
    // complete, macro_xs / argmax of the other bands' lookups read 0 / -1)
    DeviceState &d = ctx->dev[0];
    CUDA_TRY(cudaSetDevice(d.device));
    double *d_e = nullptr, *d_macro = nullptr;
    int *d_mat = nullptr, *d_am = nullptr;
    CUDA_TRY(cudaMalloc(&d_e, (size_t)n * sizeof(double)));
    CUDA_TRY(cudaMalloc(&d_macro, (size_t)n * 5 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&d_mat, (size_t)n * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_am, (size_t)n * sizeof(int)));
    CUDA_TRY(cudaMemsetAsync(d.accum, 0, 3 * sizeof(unsigned long long), d.stream));
    CUDA_TRY(cudaMemsetAsync(d.counters, 0, kNumCounters * sizeof(unsigned int), d.stream));
    xs::BatchSource src{};
    src.first_id = first_id; src.count = n; src.mat_lo = 0; src.mat_hi = XS_NUM_MATERIALS - 1;
    if (ctx->n_bands > 1) {
        if (!ctx->tile) { cudaFree(d_e); cudaFree(d_macro); cudaFree(d_mat); cudaFree(d_am);
                          return set_error(XS_ERR_UNSUPP, "xs_gpu_dump: an energy-band-sharded grid needs the tile kernel (XSB200_TILE=1)"); }
        src.row_begin = (uint32_t)d.row0; src.row_end = (uint32_t)d.row1;
        CUDA_TRY(cudaMemsetAsync(d_macro, 0, (size_t)n * 5 * sizeof(double), d.stream));
        CUDA_TRY(cudaMemsetAsync(d_am, 0xff, (size_t)n * sizeof(int), d.stream));
    }
    xs::BatchSink sink{};
    sink.accum = d.accum; sink.macro_out = d_macro; sink.energy_out = d_e; sink.mat_out = d_mat; sink.argmax_out = d_am;
    int rc = launch_event(ctx, d, src, sink, 0);
    if (rc == XS_OK) {
        cudaError_t e = cudaStreamSynchronize(d.stream);
        if (e == cudaSuccess && h_energy_out) e = cudaMemcpy(h_energy_out, d_e, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && h_mat_out) e = cudaMemcpy(h_mat_out, d_mat, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && h_macro_xs_out) e = cudaMemcpy(h_macro_xs_out, d_macro, (size_t)n * 5 * sizeof(double), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && h_argmax_out) e = cudaMemcpy(h_argmax_out, d_am, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = set_error(XS_ERR_CUDA, "xs_gpu_dump: %s", cudaGetErrorString(e));
    }
    cudaFree(d_e); cudaFree(d_macro); cudaFree(d_mat); cudaFree(d_am);
    return rc;
}

int xs_gpu_sort_keys(xs_gpu_ctx *ctx, const uint32_t *h_keys, long n, int lo_bit, int hi_bit, uint32_t *h_perm_out)
{
    DeviceGuard restore_device;
    if (!ctx || !h_keys || !h_perm_out || n < 0 || lo_bit < 0 || hi_bit > 32 || lo_bit >= hi_bit)
        return set_error(XS_ERR_ARG, "xs_gpu_sort_keys: bad argument");
    if (n == 0) return XS_OK;
    if (n > 0xffffffffL) return set_error(XS_ERR_ARG, "xs_gpu_sort_keys: at most 2^32-1 keys");
    DeviceState &d = ctx->dev[0];
    int rc = ensure_sample_buffers(d, n, true);
    if (rc != XS_OK) return rc;
    CUDA_TRY(cudaSetDevice(d.device));
    CUDA_TRY(cudaMemcpyAsync(d.key[0], h_keys, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, d.stream));
    d.digits_for = -1;
    uint32_t *sorted_perm = nullptr;
    int launches = 0;
    if (sort_lookup_keys(ctx, d, d.key, d.perm, n, lo_bit, hi_bit, &sorted_perm, &launches) != 0)
        return set_error(XS_ERR_CUDA, "radix sort failed: %s", cudaGetErrorString(cudaGetLastError()));
    CUDA_TRY(cudaMemcpyAsync(h_perm_out, sorted_perm, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, d.stream));
    CUDA_TRY(cudaStreamSynchronize(d.stream));
    return XS_OK;
}

int xs_gpu_narrow_materials(const int *mat, unsigned char *out, long n, int threads)
{
    if (n < 0 || threads < 1 || threads > 64 || (n > 0 && (!mat || !out)))
        return set_error(XS_ERR_ARG, "xs_gpu_narrow_materials: bad argument");
    xs::PackPool pool(threads - 1);
    pool.run(mat, out, n);
    return XS_OK;
}

int xs_gpu_selftest_division(xs_gpu_ctx *ctx, unsigned long long seed, long n_pairs, int mode, unsigned long long *mismatches)
{
    DeviceGuard restore_device;
    if (!ctx || !mismatches || n_pairs < 0 || mode < 0 || mode > 2) return set_error(XS_ERR_ARG, "xs_gpu_selftest_division: bad argument");
    DeviceState &d = ctx->dev[0];
    CUDA_TRY(cudaSetDevice(d.device));
    CUDA_TRY(cudaMemsetAsync(d.accum, 0, 3 * sizeof(unsigned long long), d.stream));
    if (n_pairs > 0) {
        xs::xs_division_selftest_kernel<<<d.sm_count * 16, 256, 0, d.stream>>>(seed, n_pairs, mode, d.accum);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaMemcpyAsync(d.h_accum, d.accum, sizeof(unsigned long long), cudaMemcpyDeviceToHost, d.stream));
    CUDA_TRY(cudaStreamSynchronize(d.stream));
    *mismatches = d.h_accum[0];
    return XS_OK;
}

int xs_gpu_read_array(xs_gpu_ctx *ctx, int which, long offset_bytes, long n_bytes, void *h_dst)
{
    DeviceGuard restore_device;
    if (!ctx || !h_dst || offset_bytes < 0 || n_bytes < 0) return set_error(XS_ERR_ARG, "xs_gpu_read_array: bad argument");
    DeviceState &d = ctx->dev[0];
    const long n_points = ctx->n_iso * ctx->n_gp;
    const unsigned char *src = nullptr;
    long total = 0;
    if (which == XS_ARRAY_NUCLIDE_GRID) { src = reinterpret_cast<const unsigned char *>(d.grid); total = n_points * 48; }
    else if (which == XS_ARRAY_UNIONIZED_ENERGY && ctx->grid_type == XS_UNIONIZED) {
        src = reinterpret_cast<const unsigned char *>(d.P.ueg); total = n_points * 8;
    } else if (which == XS_ARRAY_INDEX_GRID && ctx->grid_type != XS_NUCLIDE) {
        src = reinterpret_cast<const unsigned char *>(d.P.index_grid);
        total = (ctx->grid_type == XS_UNIONIZED ? n_points : (long)ctx->hash_bins) * ctx->n_iso * 4;
        if (ctx->n_bands > 1 && (offset_bytes < d.row0 * ctx->n_iso * 4 || offset_bytes + n_bytes > d.row1 * ctx->n_iso * 4))
            return set_error(XS_ERR_ARG, "xs_gpu_read_array: GPU 0 holds index rows [%ld, %ld) only (energy-band sharding)", d.row0, d.row1);
    } else return set_error(XS_ERR_ARG, "xs_gpu_read_array: array %d does not exist for this grid type", which);
    if (offset_bytes + n_bytes > total) return set_error(XS_ERR_ARG, "xs_gpu_read_array: range beyond the array (%ld bytes)", total);
    CUDA_TRY(cudaSetDevice(d.device));
    CUDA_TRY(cudaStreamSynchronize(d.stream));
    CUDA_TRY(cudaMemcpy(h_dst, src + offset_bytes, (size_t)n_bytes, cudaMemcpyDeviceToHost));
    return XS_OK;
}

int xs_gpu_set_stream(xs_gpu_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return set_error(XS_ERR_ARG, "xs_gpu_set_stream: NULL context");
    DeviceState &d = ctx->dev[0];
    CUDA_TRY(cudaSetDevice(d.device));
    CUDA_TRY(cudaStreamSynchronize(d.stream));
    if (d.own_stream) cudaStreamDestroy(d.stream);
    d.stream = (cudaStream_t)cuda_stream;
    d.own_stream = false;
    return XS_OK;
}

int xs_gpu_get_info(const xs_gpu_ctx *ctx, xs_gpu_info *info)
{
    if (!ctx || !info) return set_error(XS_ERR_ARG, "xs_gpu_get_info: NULL argument");
    const DeviceState &d = ctx->dev[0];
    info->device = d.device;
    info->sm_count = d.sm_count;
    info->l2_bytes = d.l2_bytes;
    info->resident_bytes = (long)d.resident_bytes;
    info->n_isotopes = ctx->n_iso;
    info->n_gridpoints = ctx->n_gp;
    info->grid_type = ctx->grid_type;
    info->hash_bins = ctx->hash_bins;
    info->max_num_nucs = ctx->max_num_nucs;
    info->n_ueg = ctx->n_ueg;
    info->fp64_ops_per_pair = ctx->exact_arith ? 24 : 12;
    return XS_OK;
}

int xs_gpu_finalize(xs_gpu_ctx *ctx)
{
    DeviceGuard restore_device;
    if (!ctx) return XS_OK;
    xs_multi_destroy(ctx);
    delete ctx->pack_pool;
    ctx->pack_pool = nullptr;
    for (DeviceState &d : ctx->dev) {
        if (d.device < 0) continue;
        cudaSetDevice(d.device);
        if (d.stream) cudaStreamSynchronize(d.stream);
        cudaFree(d.hot_slab); cudaFree(d.index_grid); cudaFree(d.grid);
        cudaFree(d.mat_first); cudaFree(d.mat_nuc); cudaFree(d.mat_conc);
        cudaFree(d.accum); cudaFree(d.counters); cudaFree(d.histogram); cudaFree(d.dense_counter); cudaFree(d.seg_tables);
        cudaFree(d.samp_e); cudaFree(d.samp_mat);
        for (int i = 0; i < 2; i++) { cudaFree(d.key[i]); cudaFree(d.perm[i]); }
        cudaFree(d.bin_count); cudaFree(d.bin_chunk_sum);
        xs::sort_scratch_free(d.sort);
        xs::onesweep_scratch_free(d.sweep);
        cudaFree(d.samp_where); cudaFree(d.samp_pack); cudaFree(d.grp_e); cudaFree(d.grp_where); cudaFree(d.grp_mat); cudaFree(d.grp_id);
        cudaFree(d.sweep_partial); cudaFree(d.pairs); cudaFree(d.hist_seed); cudaFree(d.hist_fwd); cudaFree(d.nuc_bucket);
        cudaFree(d.dump_macro);
        if (d.h_accum) cudaFreeHost(d.h_accum);
        if (d.h_mat8) cudaFreeHost(d.h_mat8);
        if (d.h_hist) cudaFreeHost(d.h_hist);
        for (int i = 0; i < EV_COUNT; i++) if (d.ev[i]) cudaEventDestroy(d.ev[i]);
        for (int i = 0; i < kMaxChunks; i++) if (d.ev_copy[i]) cudaEventDestroy(d.ev_copy[i]);
        for (int i = 0; i < kMaxChunks; i++) if (d.ev_chunk[i]) cudaEventDestroy(d.ev_chunk[i]);
        for (int i = 0; i < kMaxChunks; i++) if (d.ev_ecopy[i]) cudaEventDestroy(d.ev_ecopy[i]);
        if (d.ev_ready) cudaEventDestroy(d.ev_ready);
        if (d.copy_stream) cudaStreamDestroy(d.copy_stream);
        if (d.own_stream && d.stream) cudaStreamDestroy(d.stream);
    }
    cudaGetLastError();
    delete ctx;
    return XS_OK;
}

}  // extern "C"
