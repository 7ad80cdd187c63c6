// xs_dense.cuh -- xs_dense_kernel: the lookup kernel of the sorted pipeline (-k 6, xs_gpu_lookup_samples)
// for DENSE segments, i.e. materials with many lookups per grid interval (large/fuel at 17 M: 209).
//
// What it computes (reference: calculate_macro_xs / calculate_micro_xs, cuda/Simulation.cu:102-236,
// as called by xs_lookup_kernel_optimization_4, :823-875): for every lookup of a material, the five
// macroscopic cross sections = sum over the material's nuclides, in the reference's order, of the
// interpolated microscopic values times the concentration; then the argmax into the checksum.
//
// How.  The 32 x PL energy-sorted lookups of a warp-group (PL consecutive ones per lane) fall, per
// nuclide, into ONE grid interval (63 % of the steps of large/fuel) or a few consecutive ones.  So only
// the group's lowest and highest energy are resolved per nuclide (lane l resolves nuclide c + l, a
// chunk of 32 nuclides ahead of the arithmetic: two index-row segments per group and chunk instead
// of 96 in unionized mode, two bucket searches per nuclide instead of 96 in hash / nuclide mode).
// The search is monotone in the energy, so that gives the first record k_min and the number of
// records n = k_max - k_min + 1 the group can touch.  min(n, 4) records per step travel through a
// per-warp shared-memory ring (cp.async, issued two batches of 4 steps ahead) and are read back as
// LDS broadcasts:
//   n == 1 (the fast path): every lookup of the group uses that record.  The step is 7 LDS.128, one
//           uniform load of the concentration and the FP64 work of PL lookups -- nothing else;
//   n == 2: one 64-bit integer compare of the energy's bit pattern with the bound between the two;
//   n >= 3: compares with the bounds of up to 4 ring records; a lookup beyond the ring, or exactly ON
//           a bound (where the reference's hash-grid procedure has its own ideas, cuda/Simulation.cu:
//           150-156), resolves its own interval with the reference's procedure (nuclide_low).
// Results therefore never depend on how well the batch is sorted: min / max are taken over the
// group, not assumed from positions.
//
// Round 2 rewrite (profiles/r02_notes.md): the round-1 kernel ran the FP64 pipe at 59 % with 2.3
// issued instructions per FP64 instruction.  Its fast and slow paths shared one body (per-lookup
// address / ok flags), which cost the n == 1 steps 56 bookkeeping instructions per 72 FP64 and put a
// branch between the lookups of a lane (no interleaving of their dependency chains).  Here the n == 1
// step is straight-line code, the ring is filled 4 steps per LDGSTS instruction (lane = (step, 16-byte
// piece)) and runs on across nuclide chunks instead of draining every 32 steps, and the steps are
// unrolled by the batch size so ring-slot addresses are immediates.
//
// Arithmetic (template parameter EXACT):
//   true  -- the reference's own roundings: f = (hi.E - E) / d from a stored correctly rounded
//            reciprocal by one Newton-Markstein step (4 operations), then per channel
//            hi - f * (hi - lo), * conc, += : 24 FP64 operations per (lookup, nuclide), macro_xs
//            bit-identical to the reference (tests/test_gpu_parity.py);
//   false -- fused: f = (hi.E - E) * (1/d), fma(-f, dlt, hi), fma(x, conc, acc): 12 operations,
//            macro_xs within a few ulp (contract: 1e-12 relative); a lookup whose two largest
//            channels are within 1e-10 (or, in history mode, a channel within 1e-10 of the "> 1"
//            test) is recomputed in the reference's exact order, so integers stay bit-exact.
#pragma once

#include "xs_kernels.cuh"

namespace xs {

#ifndef XS_DENSE_THREADS
#define XS_DENSE_THREADS 256
#endif
#ifndef XS_DENSE_BLOCKS
#define XS_DENSE_BLOCKS 2
#endif
constexpr int kDenseThreads = XS_DENSE_THREADS;
constexpr int kDenseWarps = kDenseThreads / 32;
#ifndef XS_DENSE_BATCH
#define XS_DENSE_BATCH 4             // steps per cp.async batch = unroll of the step loop (one LDGSTS covers 4 first records)
#endif
#ifndef XS_DENSE_DEPTH
#define XS_DENSE_DEPTH 3             // batches in flight ahead of the one being consumed (1 or 3: the ring is a power of two)
#endif
#ifndef XS_DENSE_STAGE
#define XS_DENSE_STAGE 1             // the warp's NEXT group: samples staged in shared memory by cp.async, index rows prefetched
#endif
#ifndef XS_DENSE_UNROLL
#define XS_DENSE_UNROLL 1            // copies of the step body inside a batch (1: see the note at the loop)
#endif
#ifndef XS_DENSE_SPAN
#define XS_DENSE_SPAN 4
#endif
#ifndef XS_DENSE_PER_LANE
#define XS_DENSE_PER_LANE 3          // round 1, lookup phase of -k 6, large: 2 -> 2.35 ms, 3 -> 2.21, 4 -> 2.26
#endif
constexpr int kDensePerLane = XS_DENSE_PER_LANE;       // consecutive lookups per lane
constexpr int kDenseGroup = 32 * kDensePerLane;        // lookups per warp-group
constexpr int kDenseBatch = XS_DENSE_BATCH;
constexpr int kDenseUnroll = XS_DENSE_UNROLL;
constexpr int kDenseDepth = XS_DENSE_DEPTH;
constexpr int kDenseRing = kDenseBatch * (kDenseDepth + 1);   // steps of records resident per warp
constexpr int kDenseSpan = XS_DENSE_SPAN;              // records per ring slot (2..4)
constexpr int kDenseSlotBytes = kDenseSpan * 128;
constexpr int kDenseRingBytes = kDenseRing * kDenseSlotBytes;
constexpr int kDenseFirstWords = 2 * 32 * 2;           // per warp: (first record, record count) per step, double-buffered by chunk
constexpr int kDenseStageBytes = XS_DENSE_STAGE ? kDenseGroup * 16 : 0;   // per warp: the next group's packed (energy, row) samples
static_assert(kDenseBatch == 4 && kDenseSpan >= 2 && kDenseSpan <= 4 && (kDenseRing & (kDenseRing - 1)) == 0, "dense kernel geometry");

XS_DEV uint2 lds_v2_u32(uint32_t smem_addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(smem_addr));
    return v;
}
XS_DEV int lds_s32(uint32_t smem_addr)
{
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_addr));
    return v;
}
XS_DEV void sts_v2_u32(uint32_t smem_addr, uint32_t x, uint32_t y)
{
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(smem_addr), "r"(x), "r"(y) : "memory");
}
XS_DEV long long lds_s64(uint32_t smem_addr)
{
    long long v;
    asm volatile("ld.shared.s64 %0, [%1];" : "=l"(v) : "r"(smem_addr));
    return v;
}

// Conditional record loads IN PLACE (predicated, operands tied): written as C++ ("if (p) r = load()")
// ptxas keeps the old and the new record in two register sets and merges them with 26 moves per
// lookup; and an unconditional per-lookup load costs what actually binds this kernel next to FP64
// issue -- the shared-memory data path: 128 bytes per clock per SM INTO registers, broadcast or not, so
// a warp-wide read of one record is 26 wavefronts, a read by the one lane that needs it is 7.
XS_DEV void lds_record_if(PairRecord &r, bool p, uint32_t smem_addr)
{
    asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %13, 0;\n"
                 "@q ld.shared.v2.f64 {%0,%1}, [%14+96];\n"
                 "@q ld.shared.f64 %2, [%14+112];\n"
                 "@q ld.shared.v2.f64 {%3,%4}, [%14];\n"
                 "@q ld.shared.v2.f64 {%5,%6}, [%14+16];\n"
                 "@q ld.shared.v2.f64 {%7,%8}, [%14+32];\n"
                 "@q ld.shared.v2.f64 {%9,%10}, [%14+48];\n"
                 "@q ld.shared.v2.f64 {%11,%12}, [%14+64];\n}"
                 : "+d"(r.hi_e), "+d"(r.d), "+d"(r.inv), "+d"(r.hi[0]), "+d"(r.dlt[0]), "+d"(r.hi[1]), "+d"(r.dlt[1]),
                   "+d"(r.hi[2]), "+d"(r.dlt[2]), "+d"(r.hi[3]), "+d"(r.dlt[3]), "+d"(r.hi[4]), "+d"(r.dlt[4])
                 : "r"((uint32_t)p), "r"(smem_addr));
}
XS_DEV void ldg_record_if(PairRecord &r, bool p, const double2 *rec)
{
    asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %13, 0;\n"
                 "@q ld.global.nc.v4.f64 {%3,%4,%5,%6}, [%14];\n"
                 "@q ld.global.nc.v4.f64 {%7,%8,%9,%10}, [%14+32];\n"
                 "@q ld.global.nc.v2.f64 {%11,%12}, [%14+64];\n"
                 "@q ld.global.nc.v2.f64 {%0,%1}, [%14+96];\n"
                 "@q ld.global.nc.f64 %2, [%14+112];\n}"
                 : "+d"(r.hi_e), "+d"(r.d), "+d"(r.inv), "+d"(r.hi[0]), "+d"(r.dlt[0]), "+d"(r.hi[1]), "+d"(r.dlt[1]),
                   "+d"(r.hi[2]), "+d"(r.dlt[2]), "+d"(r.hi[3]), "+d"(r.dlt[3]), "+d"(r.hi[4]), "+d"(r.dlt[4])
                 : "r"((uint32_t)p), "l"(rec));
}

// One (lookup, nuclide) term.  EXACT: the reference's roundings (see the header of this file).
template <bool EXACT>
XS_DEV void dense_term(const PairRecord &r, double e, double conc, double acc[5])
{
    if (EXACT) {
        record_step(r, e, conc, acc);
    } else {
        const double f = (r.hi_e - e) * r.inv;
#pragma unroll
        for (int k = 0; k < 5; k++) acc[k] = __fma_rn(__fma_rn(-f, r.dlt[k], r.hi[k]), conc, acc[k]);
    }
}

// Finished lookups of a lane (fused arithmetic): like finish_lane_lookups, but a lookup whose integer
// outputs could depend on the last bits is first recomputed in the reference's exact order.
template <int GRID, int PL>
XS_DEV void finish_lane_lookups_guarded(const Problem &P, const WindowArgs &A, const BatchSink &sink, int mat, long t0,
                                        const bool on[PL], const double e[PL], double acc[PL][5], unsigned int &my_sum)
{
#pragma unroll
    for (int w = 0; w < PL; w++) {
        if (!on[w]) continue;
        double gap;
        argmax5(acc[w], gap);
        bool redo = gap <= kTieGuard;
        if (sink.fwd_out) {
#pragma unroll
            for (int k = 0; k < 5; k++) redo |= fabs(acc[w][k] - 1.0) <= kTieGuard;
        }
        if (redo) {
            double exact[5];                                 // (a temporary: acc[][] must stay in registers)
            macro_xs_serial<GRID>(P, e[w], mat, exact);
#pragma unroll
            for (int k = 0; k < 5; k++) acc[w][k] = exact[k];
        }
    }
    finish_lane_lookups<PL>(A, sink, t0, on, acc, my_sum);
}

// Hide a value from the optimiser: what is computed from the result cannot be hoisted out of the rare
// path it is used in (ptxas otherwise moves the loop-invariant part of the per-lookup fallback -- bucket
// index, hash bin: 30 instructions -- in front of every batch of steps).
XS_DEV double opaque(double v) { asm volatile("" : "+d"(v)); return v; }
XS_DEV uint32_t opaque(uint32_t v) { asm volatile("" : "+r"(v)); return v; }
XS_DEV const double2 *opaque(const double2 *v) { asm volatile("" : "+l"(v)); return v; }

template <int WARPS>
XS_DEV unsigned long long block_sum_n(unsigned long long v, unsigned long long *s_part)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(kFullMask, v, off);
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) s_part[warp] = v;
    __syncthreads();
    unsigned long long t = 0;
    if (threadIdx.x == 0)
        for (int w = 0; w < WARPS; w++) t += s_part[w];
    __syncthreads();
    return t;      // valid on thread 0
}

template <int GRID, bool EXACT>
__global__ void __launch_bounds__(kDenseThreads, XS_DENSE_BLOCKS)
xs_dense_kernel(const __grid_constant__ Problem P, const __grid_constant__ WindowArgs A, const BatchSink sink, const __grid_constant__ ConcTable C)
{
    constexpr int PL = kDensePerLane;
    __shared__ unsigned long long s_part[kDenseWarps];
    extern __shared__ __align__(128) uint32_t s_dyn[];       // [record rings][(first record, count) per step][nuclide ids]
    uint32_t *s_first = s_dyn + kDenseWarps * kDenseRingBytes / 4;
    uint32_t *s_stage = s_first + kDenseWarps * kDenseFirstWords;
    int *s_nuc = (int *)(s_stage + kDenseWarps * kDenseStageBytes / 4);
    __shared__ SegTable T;
    load_seg_table(A, T);
    for (int i = threadIdx.x; i < P.mat_total; i += blockDim.x) s_nuc[i] = P.mat_nuc[i];
    __syncthreads();

    const int lane = (int)opaque((uint32_t)(threadIdx.x & 31)), warp = threadIdx.x >> 5;
    unsigned int my_sum = 0;
    uint2 *warp_first = (uint2 *)(s_first + warp * kDenseFirstWords);
    const uint32_t nuc_base = opaque((uint32_t)__cvta_generic_to_shared(s_nuc));
    // (opaque: ptxas otherwise re-derives these addresses from %tid and the shared window base in
    // front of every batch -- 30 instructions -- instead of keeping four registers)
    const uint32_t ring = opaque((uint32_t)__cvta_generic_to_shared(s_dyn) + warp * kDenseRingBytes);
    const uint32_t first_base = opaque((uint32_t)__cvta_generic_to_shared(warp_first));
    // this lane's part of a batch copy: 16-byte piece `piece` of the first record of step (batch start + q)
    const int q_l = lane >> 3, piece_l = lane & 7;
    const double2 *my_pairs = opaque(P.pairs + piece_l);
    const uint32_t dst_lane = opaque(ring + (uint32_t)(q_l * kDenseSlotBytes + piece_l * 16));
    const uint32_t desc_lane = opaque(first_base + (uint32_t)(q_l * 8));
    // this lane's PL entries of the warp's sample staging buffer
    const uint32_t stage_lane = opaque((uint32_t)__cvta_generic_to_shared(s_stage) + (uint32_t)(warp * kDenseStageBytes + lane * PL * 16));
    bool staged = false;                                     // the samples of the group about to start are in the staging buffer

    int g = warp_next_group(sink, lane);
    while (g < T.n_groups) {
        const int g_next = warp_next_group(sink, lane);      // (known early: its samples are requested below)
        const int sg = segment_of_group(T, g);
        const WindowSegment &S = T.seg[sg];
        const int group_first = (g - S.group_begin) * kDenseGroup;
        const int first_in_seg = group_first + lane * PL;
        const long t0 = S.offset + first_in_seg;
        double e[PL];
        uint32_t where32[PL];
        bool on[PL];
        if (XS_DENSE_STAGE && staged) {
            // fetched while the previous group was being worked on (slots past the end of the segment repeat
            // its last lookup: inside the group's energy range)
#pragma unroll
            for (int w = 0; w < PL; w++) {
                double2 v;
                asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(stage_lane + w * 16));
                e[w] = v.x;
                where32[w] = (uint32_t)__double_as_longlong(v.y);
                on[w] = first_in_seg + w < S.count;
            }
        } else {
            // idle slots repeat the group's first lookup: they do not widen the group's energy range
            load_lane_samples<PL>(A, S, first_in_seg, S.offset + group_first, e, where32, on);
        }
        // The group's energy range.  Energies are non-negative doubles: their bit patterns order
        // like the values (and 64-bit integer compares run on the ALU pipe instead of queueing
        // behind the FP64 work).  The UEG row / hash bin is monotone in the energy, so the
        // extreme rows belong to the extreme energies.
        long long eb_min = __double_as_longlong(e[0]), eb_max = eb_min;
        uint32_t where_min = where32[0], where_max = where32[0];
#pragma unroll
        for (int w = 1; w < PL; w++) {
            eb_min = min(eb_min, __double_as_longlong(e[w]));
            eb_max = max(eb_max, __double_as_longlong(e[w]));
            where_min = min(where_min, where32[w]);
            where_max = max(where_max, where32[w]);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            eb_min = min(eb_min, __shfl_xor_sync(kFullMask, eb_min, off));
            eb_max = max(eb_max, __shfl_xor_sync(kFullMask, eb_max, off));
        }
        where_min = __reduce_min_sync(kFullMask, where_min);
        where_max = __reduce_max_sync(kFullMask, where_max);
        const double e_min = __longlong_as_double(eb_min), e_max = __longlong_as_double(eb_max);

        const int n_nuc = S.j_end;                            // whole material (j_begin = 0)
        const int n_steps = (n_nuc + kDenseBatch - 1) & ~(kDenseBatch - 1);   // the tail is padded: concentration 0
        const int my_last = n_steps - q_l;                   // this lane copies for step s + q_l while s < my_last
        const int ci = C.first[S.mat];
        const uint32_t nucs = nuc_base + 4u * (uint32_t)S.first;     // shared address of the material's nuclide list
        double acc[PL][5];
#pragma unroll
        for (int w = 0; w < PL; w++)
#pragma unroll
            for (int k = 0; k < 5; k++) acc[w][k] = 0.0;

        // lane l: the records the group can touch in nuclide c + l (columns past the end repeat
        // the material's last nuclide: valid records, only ever used with concentration 0).
        // Returns the ballot of the steps of that chunk that need the per-lookup selection.
        auto resolve = [&](int c) -> uint32_t {
            const int nuc_l = lds_s32(nucs + 4u * (uint32_t)min(c + lane, n_nuc - 1));
            const int k_lo = nuclide_low<GRID, false>(P, e_min, (long)where_min, nuc_l);
            const int k_hi = nuclide_low<GRID, false>(P, e_max, (long)where_max, nuc_l);
            int n = k_hi - k_lo + 1;
            // The hash-grid procedure maps an energy that EQUALS the grid point at the edge of its
            // bin's bracket to the nuclide's first / last interval (cuda/Simulation.cu:150-156).
            // Should both ends of the group be such points, the lookups in between are not: let
            // every lookup of this step resolve itself (n = 0).
            if (GRID == kHash && k_lo == k_hi && (k_lo == 0 || k_lo == P.n_gp - 2)) n = 0;
            n = min(max(n, 0), 0xffff);
            const uint32_t no = (uint32_t)nuc_l * (uint32_t)P.n_gp + (uint32_t)k_lo;
            sts_v2_u32(first_base + (uint32_t)((((c >> 5) & 1) * 32 + lane) * 8), no, (uint32_t)n);
            // a record is used by ~one block only, so its first touch comes from DRAM: start now
            prefetch_l2(P.pairs + 8 * (size_t)no);
#pragma unroll
            for (int i = 1; i < kDenseSpan; i++)
                if (i < n) prefetch_l2(P.pairs + 8 * ((size_t)no + i));
            return __ballot_sync(kFullMask, n != 1);
        };
        // Copies of one batch: lane (q, piece) moves piece `piece` of the first record of step s + q, and
        // of the records behind it when that step's group spans several (predicated: rare).
        auto issue = [&](int s) {
            if (s < my_last) {
                const uint2 fc = lds_v2_u32(desc_lane + (uint32_t)((s & 63) * 8));
                const uint32_t dst = dst_lane + (uint32_t)((s & (kDenseRing - 1)) * kDenseSlotBytes);
                const double2 *src = my_pairs + 8 * (size_t)fc.x;
                cp_async_16(dst, src);
#pragma unroll
                for (int i = 1; i < kDenseSpan; i++)
                    if ((uint32_t)i < fc.y) cp_async_16(dst + i * 128, src + 8 * i);
            }
            cp_async_commit();
        };
        // The warp's next group: fetch its sample ids now, request the samples behind them after
        // the first chunk (two dependent random reads otherwise wait in front of every group).
        uint32_t next_id[PL];
        bool next_any = false;
        int next_first = 0, next_nuc = 0;                     // the next group's material: nuclide list and length
        if (A.indirect && A.pack) {
            if (g_next < T.n_groups) {
                const WindowSegment &S2 = T.seg[segment_of_group(T, g_next, sg)];
                const int first2 = (g_next - S2.group_begin) * kDenseGroup + lane * PL;
                next_any = true;
                next_first = S2.first; next_nuc = S2.j_end;
#pragma unroll
                for (int w = 0; w < PL; w++)
                    next_id[w] = A.sample_id[S2.offset + min(first2 + w, S2.count - 1)];
            }
        }

        __syncwarp();
        uint32_t multi = resolve(0), multi_next = 0;
        __syncwarp();
        if (XS_DENSE_STAGE && next_any) {
            // the next group's samples travel with this group's first batch of records (both come from DRAM:
            // one latency, not two in a row at the next group's start)
#pragma unroll
            for (int w = 0; w < PL; w++) cp_async_16(stage_lane + w * 16, A.pack + next_id[w]);
        }
#pragma unroll
        for (int b = 0; b < kDenseDepth; b++) issue(b * kDenseBatch);

        for (int s = 0; s < n_steps; s += kDenseBatch) {
            if ((s & 31) == 0) {
                // entering a chunk of 32 nuclides: its selection mask, then resolve the chunk after it
                // (its first copies are issued kDenseDepth batches before the chunk starts)
                if (s) multi = multi_next;
                if (GRID == kUnionized && s + 64 < n_nuc) {   // the index-row segments of the chunk after the next
                    const uint32_t row_w = (lane & 1) ? where_max : where_min;
                    const int *row = P.index_grid + (size_t)row_w * (uint32_t)P.n_iso;
                    if (lane < 2)        prefetch_l2(row + lds_s32(nucs + 4u * (uint32_t)(s + 64)));
                    else if (lane >= 30) prefetch_l2(row + lds_s32(nucs + 4u * (uint32_t)min(s + 95, n_nuc - 1)));
                }
                if (s + 32 < n_nuc) { multi_next = resolve(s + 32); __syncwarp(); }
                if (!XS_DENSE_STAGE && (s == 32 || (s == 0 && n_nuc <= 32)) && next_any) {
#pragma unroll
                    for (int w = 0; w < PL; w++) prefetch_l2(A.pack + next_id[w]);
                }
            }
            cp_async_wait_group<kDenseDepth - 1>();
            __syncwarp();
            uint32_t mb = multi >> (s & 31);
            uint32_t slot = ring + (uint32_t)((s & (kDenseRing - 1)) * kDenseSlotBytes);   // (the ring is a multiple of the batch: no wrap inside one)
            // NOT unrolled: one copy of the step body (fast path ~100 instructions, slow path with its three
            // inlined per-lookup fallbacks ~600) stays resident in the instruction cache; unrolled by the batch
            // it was 3,000 instructions, 20 % of the stall samples were instruction fetches, and the register
            // allocator spilled inside the loop (profiles/r02_notes.md)
#pragma unroll kDenseUnroll
            for (int u = 0; u < kDenseBatch && s + u < n_nuc; u++, mb >>= 1, slot += kDenseSlotBytes) {   // (padded steps of the last batch: skipped)
                const double conc = C.v[ci + s + u];
                if (!(mb & 1u)) {
                    // ---- one record for the whole group: straight-line code ----
                    const PairRecord r = lds_record(slot);
#pragma unroll
                    for (int w = 0; w < PL; w++) dense_term<EXACT>(r, e[w], conc, acc[w]);
                } else {
                    // ---- several records (or none usable): every lookup picks its own.  Each lookup reads
                    // its record with its own (per-lane) address -- unconditionally: "reload only when the
                    // record differs from the previous lookup's" cost a divergent branch and 26 register
                    // moves per lookup to merge the two ways (profiles/r02_notes.md)
                    const uint32_t n_here = lds_v2_u32(first_base + (uint32_t)(((s + u) & 63) * 8)).y;
                    uint32_t addr[PL];
                    bool ok[PL];
                    if (n_here == 2) {
                        // two records: one bound between them.  Nothing lies beyond the second
                        // (the search is monotone); ON the bound the reference decides.
                        const long long hi0 = lds_s64(slot + 96);
#pragma unroll
                        for (int w = 0; w < PL; w++) {
                            const long long eb = __double_as_longlong(e[w]);
                            addr[w] = slot + (eb > hi0 ? 128u : 0u);
                            ok[w] = eb != hi0;
                        }
                    } else {
                        // record i covers (hi[i-1], hi[i]); a bound past the group's n records is
                        // stale, but it is only looked at by a lookup already beyond them.  (The
                        // last record's own bound matters when the group spans more than the ring.)
                        long long hi[kDenseSpan];
#pragma unroll
                        for (int i = 0; i < kDenseSpan; i++) hi[i] = lds_s64(slot + i * 128 + 96);
                        const uint32_t n_ring = min(n_here, (uint32_t)kDenseSpan);
#pragma unroll
                        for (int w = 0; w < PL; w++) {
                            const long long eb = __double_as_longlong(e[w]);
                            uint32_t which = 0;
                            bool beyond = true, on_bound = false;
#pragma unroll
                            for (int i = 0; i < kDenseSpan; i++) {
                                beyond = beyond & (eb > hi[i]);
                                which += beyond ? 1u : 0u;
                                on_bound = on_bound | (eb == hi[i]);
                            }
                            ok[w] = (which < n_ring) & !on_bound;
                            addr[w] = slot + min(which, (uint32_t)(kDenseSpan - 1)) * 128;
                        }
                    }
                    // All lanes read the record of their first lookup (per-lane address); a lane whose next
                    // lookup needs another one -- in a sorted group that is the ONE lane straddling a bound --
                    // reloads in place.  A lookup that is not settled by the ring (rare) fetches the record the
                    // reference's own procedure names.
                    bool bad = false;
#pragma unroll
                    for (int w = 0; w < PL; w++) bad |= !ok[w];
                    const bool any_bad = __any_sync(kFullMask, bad);
                    PairRecord r = lds_record(addr[0]);
#pragma unroll
                    for (int w = 0; w < PL; w++) {
                        if (w > 0) lds_record_if(r, addr[w] != addr[w - 1] || !ok[w - 1], addr[w]);
                        if (any_bad) {                       // warp-uniform
                            const double2 *own = P.pairs;
                            if (!ok[w]) {
                                const int nuc = lds_s32(nucs + 4u * (uint32_t)min(s + u, n_nuc - 1));
                                const double e_w = opaque(e[w]);
                                const uint32_t no = (uint32_t)nuc * (uint32_t)P.n_gp
                                                    + (uint32_t)nuclide_low<GRID, false>(P, e_w, (long)opaque(where32[w]), nuc);
                                own = P.pairs + 8 * (size_t)no;
                            }
                            ldg_record_if(r, !ok[w], own);
                        }
                        dense_term<EXACT>(r, e[w], conc, acc[w]);
                    }
                }
            }
            __syncwarp();                                    // everyone is done with these slots
            issue(s + kDenseDepth * kDenseBatch);
            if (XS_DENSE_STAGE && GRID == kUnionized && next_any && s + kDenseBatch >= n_steps) {
                // last batch: the next group's samples have landed -- start its first chunk's two index-row
                // segments on their way from DRAM (first and last sector of each, like for the chunks after)
                uint32_t w_min = 0xffffffffu, w_max = 0u;
#pragma unroll
                for (int w = 0; w < PL; w++) {
                    uint32_t row;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(row) : "r"(stage_lane + w * 16 + 8));
                    w_min = min(w_min, row);
                    w_max = max(w_max, row);
                }
                w_min = __reduce_min_sync(kFullMask, w_min);
                w_max = __reduce_max_sync(kFullMask, w_max);
                const uint32_t nucs2 = nuc_base + 4u * (uint32_t)next_first;
                const int *row = P.index_grid + (size_t)((lane & 1) ? w_max : w_min) * (uint32_t)P.n_iso;
                if (lane < 2)        prefetch_l2(row + lds_s32(nucs2));
                else if (lane >= 30) prefetch_l2(row + lds_s32(nucs2 + 4u * (uint32_t)(min(32, next_nuc) - 1)));
            }
        }
        cp_async_wait_group<0>();
        staged = XS_DENSE_STAGE && next_any;

        if (EXACT) finish_lane_lookups<PL>(A, sink, t0, on, acc, my_sum);
        else       finish_lane_lookups_guarded<GRID, PL>(P, A, sink, S.mat, t0, on, e, acc, my_sum);
        g = g_next;
    }
    // (warp_groups_done: the last warp of the launch re-arms the hand-out counter)
    if (lane == 0) {
        __threadfence();
        if (atomicAdd(sink.batch_counter + 1, 1u) == gridDim.x * kDenseWarps - 1) {
            sink.batch_counter[0] = 0;
            sink.batch_counter[1] = 0;
            __threadfence();
        }
    }
    const unsigned long long bs = block_sum_n<kDenseWarps>(my_sum, s_part);
    if (threadIdx.x == 0) {
        if (bs) atomicAdd(sink.accum, bs);
        if (blockIdx.x == 0) {
            unsigned long long done = 0;
            for (int i = 0; i < T.n_seg; i++) done += (unsigned long long)T.seg[i].count;
            atomicAdd(sink.accum + 1, done);
        }
    }
}

}  // namespace xs
