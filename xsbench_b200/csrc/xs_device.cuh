// xs_device.cuh -- device-side building blocks of the sm_100a XS-lookup engine.
//
// Everything here is written for Blackwell (compile: -gencode arch=compute_100a,code=sm_100a
// -fmad=false).  -fmad=false matters: the reference CPU build targets baseline x86-64 (no
// FMA), so with contraction off every product/difference below rounds exactly like the
// reference and the only numerical difference left is the order in which the per-nuclide
// contributions are summed (warp reductions).
//
// Reference behaviour restated (citations relative to ANL-CESAR/XSBench v20):
//   lcg_*            cuda/Simulation.cu:326-362
//   pick_material    cuda/Simulation.cu:287-324
//   ueg_row          cuda/Simulation.cu:241-261   (grid_search; closed form in SURVEY A.2)
//   nuclide_low_*    cuda/Simulation.cu:102-165   (index selection part of calculate_micro_xs)
//   interpolate      cuda/Simulation.cu:168-183
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define XS_DEV __device__ __forceinline__

namespace xs {

constexpr int kUnionized = 0, kNuclide = 1, kHash = 2;
constexpr int kNumMaterials = 12;
constexpr unsigned kFullMask = 0xffffffffu;

// ---------------------------------------------------------------------------------------
// Device view of the problem.  Passed to kernels by value (like the reference passes
// SimulationData), lives in the kernel parameter constant bank.
// ---------------------------------------------------------------------------------------
struct Problem {
    const double  *ueg;          // [n_ueg]            unionized energy grid (unionized only)
    const int     *index_grid;   // [n_ueg*n_iso] or [hash_bins*n_iso]
    const double2 *grid;         // nuclide grid viewed as 16-byte chunks: point = 3 chunks
                                 //   chunk0 = (energy,total) chunk1 = (elastic,absorbtion)
                                 //   chunk2 = (fission,nu_fission)
    const double2 *pairs;        // pair records (see xs_build_pairs_kernel): 8 chunks = 128 B per
                                 //   (nuclide, k): low point k and high point k+1 interleaved
    const uint32_t *ueg_bucket;  // [n_buckets+1]      bucket -> first UEG row in it
    const uint32_t *nuc_bucket;  // [n_iso][nuc_buckets+1]  per-nuclide bucket -> first grid point in it
                                 //   (nuclide-grid mode only; replaces the top of grid_search_nuclide)
    const int     *mat_first;    // [13]  CSR offsets of the compact material table
    const int     *mat_nuc;      // [mat_first[12]]    nuclide id
    const double  *mat_conc;     // [mat_first[12]]    concentration
    long   n_ueg;
    double bucket_scale;         // (double)n_buckets
    int    n_buckets;
    int    bucket_shift;         // 4: entries are (first row << 4 | min(rows, 15)); 0: plain first rows
    int    nuc_buckets;          // buckets per nuclide (0 = table not built)
    int    n_iso;
    int    n_gp;
    int    hash_bins;
    int    mat_total;            // mat_first[12]
    double mat_threshold[kNumMaterials];   // pick_mat's cumulative thresholds (summed in the reference's order)
};

// Concentrations for the sweep / sorted / dense kernels: one zero-padded row per material, read
// through the uniform datapath (LDC / LDCU on the kernel-parameter bank) with no bounds logic --
// padded steps simply multiply by 0.  The table is a KERNEL PARAMETER (by value, 8 KB of the 32 KB
// sm_100 allows): it belongs to the context that launches, so two live contexts with different
// material tables cannot clobber each other (a __constant__ symbol is process-global -- round 1
// kept it there).  Row m starts at first[m] (even: 16-byte aligned), holds num_nucs[m]
// concentrations and at least kConcPad zeros.
constexpr int kConcCap = 1024;
constexpr int kConcPad = 8;
struct ConcTable {
    double v[kConcCap];
    int    first[kNumMaterials];
};

// ---------------------------------------------------------------------------------------
// LCG: x <- (a x + 1) mod 2^63
// ---------------------------------------------------------------------------------------
constexpr uint64_t kLcgA = 2806196910506780709ULL;
constexpr uint64_t kLcgMask = 0x7fffffffffffffffULL;
constexpr uint64_t kStartSeed = 1070ULL;

struct Affine { uint64_t m, c; };           // x -> m x + c   (mod 2^64, masked on use)

XS_DEV uint64_t lcg_step(uint64_t s) { return (kLcgA * s + 1ULL) & kLcgMask; }
XS_DEV double   lcg_to_double(uint64_t s) { return __ull2double_rn(s) * 0x1p-63; }
XS_DEV uint64_t apply(Affine f, uint64_t s) { return (f.m * s + f.c) & kLcgMask; }

// The n-step jump as an affine map (same binary decomposition as fast_forward_LCG).
XS_DEV Affine lcg_jump(uint64_t n)
{
    uint64_t sm = kLcgA, sc = 1ULL;
    Affine r{1ULL, 0ULL};
    for (n &= kLcgMask; n; n >>= 1) {
        if (n & 1ULL) { r.m *= sm; r.c = r.c * sm + sc; }
        sc *= sm + 1ULL;
        sm *= sm;
    }
    return r;
}
XS_DEV uint64_t lcg_skip(uint64_t s, uint64_t n) { return apply(lcg_jump(n), s); }

// First i with roll < thr[i] (thr[0] == 0 never matches), else 0 = fuel.
XS_DEV int pick_material(const Problem &P, double roll)
{
    int m = 0;
#pragma unroll
    for (int i = kNumMaterials - 1; i >= 1; i--)
        if (roll < P.mat_threshold[i]) m = i;     // descending scan keeps the FIRST match
    return m;
}

// ---------------------------------------------------------------------------------------
// Loads.  Read-only (.nc) path; eviction-priority hints keep the small, hot search
// structures in L2 while the 5.7 GB index grid streams through.
//
// Measured on B200 (profiles/r01_notes.md): ".L1::no_allocate" makes a load evict_first in
// L2 as well.  With it on the nuclide-grid loads the 192 MB grid never stayed in the 126 MB
// L2 (hit rate 27 %, 162 GB of DRAM reads per run, DRAM-bound); the grid loads therefore use
// the default policy (L1 allocate, L2 evict_normal) and only the index rows are streamed.
// ---------------------------------------------------------------------------------------
#ifndef XS_GRID_LOAD
#define XS_GRID_LOAD 0
#endif
XS_DEV uint64_t policy_keep_grid()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
XS_DEV double2 ldg_grid(const double2 *p)
{
    double2 v;
#if XS_GRID_LOAD == 0
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
#elif XS_GRID_LOAD == 1      // coherent path, default policy (what the reference's loads compile to)
    asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
#elif XS_GRID_LOAD == 2      // read-only path + explicit L2 evict_last
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(policy_keep_grid()));
#elif XS_GRID_LOAD == 3      // coherent path + explicit L2 evict_last
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(policy_keep_grid()));
#elif XS_GRID_LOAD == 4      // plain C++ load (compiler's choice)
    v = *p;
#endif
    return v;
}
// One 32-byte half-record: (low.chunk, high.chunk) of a pair record; 256-bit load (sm_100+).
XS_DEV void ldg_pair(const double2 *p, double2 &lo, double2 &hi)
{
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(lo.x), "=d"(lo.y), "=d"(hi.x), "=d"(hi.y) : "l"(p));
}
XS_DEV double ldg_grid_energy(const double2 *p)
{
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
// L2 eviction-priority policy (createpolicy).  On sm_100a the bare ".L2::evict_*" qualifier
// exists only for 256-bit loads; narrower loads take a policy operand via ".L2::cache_hint".
//
// What is worth keeping in the 126 MB L2 is the 192 MB nuclide grid (96 B per nuclide and
// lookup, ~200 sectors per lookup).  Everything else is touched once per lookup -- one
// index-grid row segment, one bucket entry, 1-2 UEG entries -- and is loaded evict_first so it
// does not displace grid lines.  (Measured: pinning bucket table + UEG (40 MB) with evict_last
// left so little L2 for the grid that its hit rate fell to ~25 % and the kernel became
// DRAM-bound at 122 GB per run; see profiles/r01_notes.md.)
XS_DEV uint64_t policy_stream()
{
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));   // pure: may be hoisted/CSE'd
    return p;
}
XS_DEV int ldg_index_stream(const int *p)         // unionized index row: touched once
{
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;"
                 : "=r"(v) : "l"(p), "l"(policy_stream()));
    return v;
}
XS_DEV int ldg_index_keep(const int *p)           // hash grid (13.5 MiB, 2 reads per nuclide): default policy
{
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
XS_DEV double ldg_search_f64(const double *p)       // UEG probe: touched once per lookup
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;"
                 : "=d"(v) : "l"(p), "l"(policy_stream()));
    return v;
}
XS_DEV uint32_t ldg_search_u32(const uint32_t *p)   // bucket table entry: touched once per lookup
{
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;"
                 : "=r"(v) : "l"(p), "l"(policy_stream()));
    return v;
}

// ---------------------------------------------------------------------------------------
// Unionized grid: row = grid_search(n_ueg, E, ueg) = clamp(upper_bound(E) - 1, 0, n-2).
// A bucket table over the value range replaces the top of the binary search: bucket b holds
// the rows whose energy maps to b under the SAME monotone function used for the query, so
// upper_bound(E) lies in [bucket[b], bucket[b+1]] and only that span is searched.
// Usable per lane (each lane may search its own energy).
// ---------------------------------------------------------------------------------------
XS_DEV int bucket_of(double e, double scale, int n_buckets)
{
    int b = (int)(e * scale);
    return b < 0 ? 0 : (b >= n_buckets ? n_buckets - 1 : b);
}

// STREAM = true : probes are loaded evict_first (in-order kernels: protect the grid's L2 share);
// STREAM = false: default policy (sampler kernels run alone, the 40 MB of bucket table + UEG
//                 then stay in L2 and only their first touch reaches DRAM).
template <bool STREAM>
XS_DEV long ueg_row_t(const Problem &P, double e)
{
    const int b = bucket_of(e, P.bucket_scale, P.n_buckets);
    // rows [lo, hi) are in bucket b.  A table entry is (first row << 4 | rows in the bucket, 15 = "15 or more"):
    // one 4-byte read gives both ends, and an empty bucket (37 % of them at one row per bucket) needs no
    // probe at all -- the sampler kernels are bound by the L1 tag stage, i.e. by sectors requested per lookup
    // (bucket_shift == 0: plain first-row entries, grids of 2^28 rows and more).
    // Row numbers are 32-bit (the table's entries are): 32-bit arithmetic, half the instructions of `long`.
    const uint32_t ent = STREAM ? ldg_search_u32(P.ueg_bucket + b) : __ldg(P.ueg_bucket + b);
    uint32_t lo = ent >> P.bucket_shift;
    const uint32_t cnt = P.bucket_shift ? (ent & 15u) : 15u;
    uint32_t hi = lo + cnt;
    if (cnt == 15u) hi = (STREAM ? ldg_search_u32(P.ueg_bucket + b + 1) : __ldg(P.ueg_bucket + b + 1)) >> P.bucket_shift;
    // upper_bound within [lo, hi): first row with ueg > e
    while (hi - lo > 4u) {
        const uint32_t mid = lo + (hi - lo) / 2u;
        const double u = STREAM ? ldg_search_f64(P.ueg + mid) : __ldg(P.ueg + mid);
        if (u > e) hi = mid; else lo = mid + 1u;
    }
    // the <= 4 remaining rows are probed at once (independent loads, one base address, no compare-then-load
    // chain); the energies ascend, so "not greater than e" holds for a prefix of them.  A row that is not
    // there counts as +infinity.
    const double *q = P.ueg + lo;
    const uint32_t left = hi - lo;
    uint32_t not_greater = 0;
#pragma unroll
    for (uint32_t k = 0; k < 4u; k++) {
        double u = __longlong_as_double(0x7ff0000000000000LL);
        if (k < left) u = STREAM ? ldg_search_f64(q + k) : __ldg(q + k);
        not_greater += (u > e) ? 0u : 1u;
    }
    lo += not_greater;
    uint32_t row = lo ? lo - 1u : 0u;
    const uint32_t last = (uint32_t)(P.n_ueg - 2);
    return (long)(row > last ? last : row);
}
XS_DEV long ueg_row(const Problem &P, double e) { return ueg_row_t<true>(P, e); }

// grid_search_nuclide on the energy field of one nuclide's points (48-byte stride).
XS_DEV int search_nuclide(const double2 *g, double e, int lo, int hi)
{
    while (hi - lo > 1) {
        const int mid = lo + (hi - lo) / 2;
        if (ldg_grid_energy(g + 3 * (long)mid) > e) hi = mid; else lo = mid;
    }
    return lo;
}

// Index of the lower bounding grid point of nuclide `nuc` for energy e.
//   unionized: index_grid[row][nuc]; nuclide: full search; hash: bracketed search.
// `where` is the UEG row (unionized) or the hash bin (hash); unused for nuclide.
// Energy of grid point k of a nuclide, read from the PAIR RECORDS (record r = base + k holds
// lo.E of point k in chunk 5 and hi.E = energy of point k+1 in chunk 6).  The searches of the
// sweep path probe the same 128-byte lines the gather reads afterwards.
XS_DEV double rec_energy(const Problem &P, long base, int k)
{
    const double2 *p = (k < P.n_gp - 1) ? P.pairs + 8 * (base + k) + 5 : P.pairs + 8 * (base + k - 1) + 6;
    return __ldg(&p->x);
}
XS_DEV int search_nuclide_rec(const Problem &P, long base, double e, int lo, int hi)
{
    while (hi - lo > 1) {
        const int mid = lo + (hi - lo) / 2;
        if (rec_energy(P, base, mid) > e) hi = mid; else lo = mid;
    }
    return lo;
}

// REC = false: energies from the reference-layout grid; REC = true: from the pair records.
// WIDE = true: the last <= 4 candidates of a search are probed at once (independent loads: fewer
//               round trips -- the latency-bound callers); false: one probe at a time (fewer
//               loads -- the windowed sweep, which is bound by its L1 traffic).
template <int GRID, bool REC = false, bool WIDE = true>
XS_DEV int nuclide_low(const Problem &P, double e, long where, int nuc)
{
    const int last = P.n_gp - 1;
    const long base = (long)nuc * P.n_gp;
    const double2 *g = P.grid + 3 * base;
    int low;
    if (GRID == kUnionized) {
        low = ldg_index_stream(P.index_grid + where * P.n_iso + nuc);
    } else if (GRID == kNuclide) {
        if (P.nuc_buckets > 0) {
            // upper_bound(e) lies in [bucket[b], bucket[b+1]] (same argument as for the UEG table);
            // grid_search_nuclide(0, last) = clamp(upper_bound - 1, 0, last - 1)
            const int b = bucket_of(e, (double)P.nuc_buckets, P.nuc_buckets);
            const uint32_t *t = P.nuc_bucket + (long)nuc * (P.nuc_buckets + 1) + b;
            int lo = (int)__ldg(t), hi = (int)__ldg(t + 1);
            while (hi - lo > 4) {
                const int mid = lo + (hi - lo) / 2;
                if (ldg_grid_energy(g + 3 * (long)mid) > e) hi = mid; else lo = mid + 1;
            }
            // the <= 4 remaining points are probed at once (independent loads); the energies
            // ascend, so "not greater than e" holds for a prefix of them
            if (WIDE) {
                int not_greater = 0;
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (lo + k < hi) not_greater += !(ldg_grid_energy(g + 3 * (long)(lo + k)) > e);
                lo += not_greater;
            } else {
                while (lo < hi && !(ldg_grid_energy(g + 3 * (long)lo) > e)) lo++;
            }
            low = lo - 1;
            if (low < 0) low = 0;
        } else {
            low = REC ? search_nuclide_rec(P, base, e, 0, last) : search_nuclide(g, e, 0, last);
        }
    } else {
        const int u_lo = ldg_index_keep(P.index_grid + where * P.n_iso + nuc);
        const int u_hi = (where == P.hash_bins - 1)
                             ? last
                             : ldg_index_keep(P.index_grid + (where + 1) * P.n_iso + nuc) + 1;
        const double e_lo = REC ? rec_energy(P, base, u_lo) : ldg_grid_energy(g + 3 * (long)u_lo);
        const double e_hi = REC ? rec_energy(P, base, u_hi) : ldg_grid_energy(g + 3 * (long)u_hi);
        if (e <= e_lo)      low = 0;
        else if (e >= e_hi) low = last;
        else if (WIDE && u_hi - u_lo <= 5) {
            // e_lo < e < e_hi: the bisection returns u_lo + #{interior points <= e} (the energies
            // ascend).  A hash bin brackets ~3 points: probe the interior ones at once
            // (independent loads) instead of bisecting (dependent ones).
            int not_greater = 0;
#pragma unroll
            for (int k = 1; k <= 4; k++) {
                if (u_lo + k < u_hi) {
                    const double ek = REC ? rec_energy(P, base, u_lo + k) : ldg_grid_energy(g + 3 * (long)(u_lo + k));
                    not_greater += !(ek > e);
                }
            }
            low = u_lo + not_greater;
        }
        else                low = REC ? search_nuclide_rec(P, base, e, u_lo, u_hi) : search_nuclide(g, e, u_lo, u_hi);
    }
    return low == last ? last - 1 : low;
}

// Hash bin of an energy: two roundings, exactly like the reference (du = 1/bins; E/du).
XS_DEV long hash_bin(const Problem &P, double e)
{
    const double du = 1.0 / (double)P.hash_bins;
    long b = (long)(e / du);
    b = b > P.hash_bins - 1 ? P.hash_bins - 1 : b;      // E == 1.0 would read out of bounds
    return b < 0 ? 0 : b;                               // (so would a negative energy; the sampler never draws one)
}

template <int GRID>
XS_DEV long locate(const Problem &P, double e)
{
    if (GRID == kUnionized) return ueg_row(P, e);
    if (GRID == kHash)      return hash_bin(P, e);
    return -1;
}

XS_DEV long locate_rt(const Problem &P, int grid_type, double e)      // sampler kernels
{
    if (grid_type == kUnionized) return ueg_row_t<false>(P, e);
    if (grid_type == kHash)      return hash_bin(P, e);
    return 0;
}

// hi - f*(hi - lo), f = (hi.E - E)/(hi.E - lo.E)
XS_DEV double lerp_xs(double lo, double hi, double f) { return hi - f * (hi - lo); }

// ---------------------------------------------------------------------------------------
// Serial, reference-order evaluation of one macroscopic lookup by a single thread.  Used
// (rarely) to settle near-ties so that integer results never depend on reduction order,
// and by the one-thread-per-lookup parity kernel.
// ---------------------------------------------------------------------------------------
template <int GRID>
__device__ __noinline__ void macro_xs_serial(const Problem &P, double e, int mat, double out[5])
{
    const long where = locate<GRID>(P, e);
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const int first = P.mat_first[mat], n = P.mat_first[mat + 1] - first;
    for (int j = 0; j < n; j++) {
        const int nuc = P.mat_nuc[first + j];
        const double conc = P.mat_conc[first + j];
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
    for (int k = 0; k < 5; k++) out[k] = acc[k];
}

// First index of the strict maximum with start value -1.0 (cuda/Simulation.cu:88-97), plus
// the relative gap between the two largest values (for the near-tie guard).
XS_DEV int argmax5(const double v[5], double &rel_gap)
{
    double best = -1.0, second = -1.0;
    int at = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        if (v[k] > best) { second = best; best = v[k]; at = k; }
        else if (v[k] > second) second = v[k];
    }
    rel_gap = (best - second) / best;
    return at;
}

}  // namespace xs
