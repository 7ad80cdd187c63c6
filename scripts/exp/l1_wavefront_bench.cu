// Microbenchmark: what does a gather of one record per lookup cost in the L1 data stage, as a
// function of load width, lanes per record and record stride?  (Informs the pair-record layout of
// xs_window_kernel.)   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1bench l1_wavefront_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

struct V4 { double a, b, c, d; };
__device__ __forceinline__ V4 ld256(const char *p) { V4 v; asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p)); return v; }
__device__ __forceinline__ double2 ld128(const char *p) { double2 v; asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p)); return v; }
__device__ __forceinline__ double ld64(const char *p) { double v; asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }

// PATTERN: lanes per record L, bytes per lane W, record stride S, sorted flag
template <int L, int W, int S, bool SORTED, int MASKLAST>
__global__ void __launch_bounds__(256, 4) bench(const char *base, uint32_t rec_mask, int iters, double *out, long long *cycles)
{
    const int lane = threadIdx.x & 31;
    const int slot = SORTED ? 0 : lane / L;
    const int sub = lane % L;
    const int n_slots = 32 / L;
    const bool active = (lane / L) < n_slots && !(MASKLAST && sub == L - 1);
    const unsigned warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint64_t x = (warp_global * 64ull + slot) * 0x9E3779B97F4A7C15ull + 12345;
    double acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x = x * 2806196910506780709ull + 1;
            const uint32_t r = (uint32_t)(x >> 35) & rec_mask;
            const char *p = base + (size_t)r * S + sub * W;
            if (active) {
                if (W == 32) { V4 v = ld256(p); acc += v.a + v.b + v.c + v.d; }
                else if (W == 16) { double2 v = ld128(p); acc += v.x + v.y; }
                else { acc += ld64(p); }
            }
        }
    }
    long long t1 = clock64();
    if (acc == 1.2345e-300) out[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// Broadcast pattern: lane-per-lookup on sorted lookups -- the 32 lanes read the SAME bytes of
// DIST distinct records (lanes split evenly); OFF selects the 16/32-byte chunk inside the record.
template <int W, int DIST, bool SMEM>
__global__ void __launch_bounds__(256, 4) bcast(const char *base, uint32_t rec_mask, int iters, double *out, long long *cycles)
{
    __shared__ __align__(128) char s_buf[8][1024];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = lane * DIST / 32;
    const unsigned warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint64_t x = (warp_global * 64ull) * 0x9E3779B97F4A7C15ull + 12345;
    for (int i = lane; i < 128; i += 32) ((double *)s_buf[warp])[i] = i;
    __syncwarp();
    double acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x = x * 2806196910506780709ull + 1;
            const uint32_t r = ((uint32_t)(x >> 35) + slot) & rec_mask;
            if (SMEM) {
                const char *p = s_buf[warp] + ((r & 7) * 128) + (u & 3) * W;
                if (W == 16) { double2 v = *(const double2 *)p; acc += v.x + v.y; }
                else { acc += *(const double *)p; }
            } else {
                const char *p = base + (size_t)r * 128 + (u & 3) * W;
                if (W == 32) { V4 v = ld256(p); acc += v.a + v.b + v.c + v.d; }
                else if (W == 16) { double2 v = ld128(p); acc += v.x + v.y; }
                else { acc += ld64(p); }
            }
        }
    }
    long long t1 = clock64();
    if (acc == 1.2345e-300) out[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int W, int DIST, bool SMEM>
void run_bcast(const char *name, const char *buf, size_t buf_bytes, int iters, double *out, long long *cyc, int blocks)
{
    uint32_t n_rec = 1; while ((size_t)(n_rec * 2) * 128 + 256 <= buf_bytes) n_rec *= 2;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    bcast<W, DIST, SMEM><<<blocks, 256>>>(buf, n_rec - 1, iters / 8, out, cyc);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    bcast<W, DIST, SMEM><<<blocks, 256>>>(buf, n_rec - 1, iters, out, cyc);
    cudaEventRecord(b); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long *h = (long long *)malloc(blocks * sizeof(long long));
    cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < blocks; i++) mean += h[i]; mean /= blocks;
    const double ldg_per_sm = blocks * 8.0 / 148.0 * iters * 8.0;
    printf("%-46s buf %6.1f MB  %8.3f ms  %6.2f cyc/warp-load (%d B/lane, %d distinct records, %s)\n", name, buf_bytes / 1e6, ms,
           mean / ldg_per_sm, W, DIST, SMEM ? "LDS" : "LDG");
    free(h);
}

template <int L, int W, int S, bool SORTED, int MASKLAST>
void run(const char *name, const char *buf, size_t buf_bytes, int iters, double *out, long long *cyc, int blocks)
{
    uint32_t n_rec = 1; while ((size_t)(n_rec * 2) * S + 256 <= buf_bytes) n_rec *= 2;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    bench<L, W, S, SORTED, MASKLAST><<<blocks, 256>>>(buf, n_rec - 1, iters / 8, out, cyc);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    bench<L, W, S, SORTED, MASKLAST><<<blocks, 256>>>(buf, n_rec - 1, iters, out, cyc);
    cudaEventRecord(b); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long *h = (long long *)malloc(blocks * sizeof(long long));
    cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < blocks; i++) mean += h[i]; mean /= blocks;
    const int slots = SORTED ? 32 / L : 32 / L;
    const double warps_per_sm = blocks * 8.0 / 148.0;
    const double ldg_per_sm = warps_per_sm * iters * 8.0;            // warp-level load instructions per SM
    const double cyc_per_ldg = mean / ldg_per_sm * 1.0;              // all resident at once (4 blocks/SM): wall cycles = mean
    const double recs = (double)blocks * 8 * iters * 8.0 * slots;
    printf("%-46s buf %6.1f MB  %8.3f ms  %6.2f cyc/warp-LDG  %6.3f cyc/record/SM  %7.1f Grec/s  %6.2f TB/s useful\n", name, buf_bytes / 1e6, ms,
           cyc_per_ldg, cyc_per_ldg / slots, recs / ms / 1e6, recs * (L - MASKLAST) * W / ms / 1e9);
    free(h);
}

int main(int argc, char **argv)
{
    const int blocks = 148 * 4;
    double *out; long long *cyc; CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&cyc, blocks * 8));
    for (size_t mb : {0, 48}) {
        size_t bytes = mb ? mb << 20 : 48 << 10;     // 48 KB: L1-resident;  48 MB: L2-resident (a 32-nuclide window)
        char *buf; CK(cudaMalloc(&buf, bytes + 4096)); CK(cudaMemset(buf, 0, bytes + 4096));
        int iters = mb ? 400 : 2000;
        run<4, 32, 128, false, 0>("A  4 lanes x 32 B, 128-B records (current)", buf, bytes, iters, out, cyc, blocks);
        run<4, 32, 128, true, 0>("B  same, all 8 lookups on one record (sorted)", buf, bytes, iters, out, cyc, blocks);
        run<4, 32, 128, false, 1>("C  3 of 4 lanes load (96 of 128 B)", buf, bytes, iters, out, cyc, blocks);
        run<3, 32, 96, false, 0>("D  3 lanes x 32 B, 96-B records, 10 per warp", buf, bytes, iters, out, cyc, blocks);
        run<3, 32, 128, false, 0>("E  3 lanes x 32 B, 128-B stride, 10 per warp", buf, bytes, iters, out, cyc, blocks);
        run<8, 16, 128, false, 0>("F  8 lanes x 16 B, 128-B records", buf, bytes, iters, out, cyc, blocks);
        run<6, 16, 48, false, 0>("G  6 lanes x 16 B, 48-B points (2 adjacent)", buf, bytes, iters, out, cyc, blocks);
        run<6, 16, 96, false, 0>("H  6 lanes x 16 B, 96-B records", buf, bytes, iters, out, cyc, blocks);
        run<16, 8, 128, false, 0>("I  16 lanes x 8 B, 128-B records", buf, bytes, iters, out, cyc, blocks);
        run<2, 32, 64, false, 0>("J  2 lanes x 32 B, 64-B records, 16 per warp", buf, bytes, iters, out, cyc, blocks);
        run<1, 32, 32, false, 0>("K  1 lane x 32 B, 32-B records, 32 per warp", buf, bytes, iters, out, cyc, blocks);
        run<2, 32, 128, false, 0>("L  2 lanes x 32 B of 128-B records", buf, bytes, iters, out, cyc, blocks);
        run_bcast<8, 1, false>("M  broadcast LDG.64, 1 record", buf, bytes, iters, out, cyc, blocks);
        run_bcast<16, 1, false>("N  broadcast LDG.128, 1 record", buf, bytes, iters, out, cyc, blocks);
        run_bcast<32, 1, false>("O  broadcast LDG.256, 1 record", buf, bytes, iters, out, cyc, blocks);
        run_bcast<16, 2, false>("P  LDG.128, 2 distinct records", buf, bytes, iters, out, cyc, blocks);
        run_bcast<32, 2, false>("Q  LDG.256, 2 distinct records", buf, bytes, iters, out, cyc, blocks);
        run_bcast<32, 4, false>("R  LDG.256, 4 distinct records", buf, bytes, iters, out, cyc, blocks);
        run_bcast<16, 4, false>("S  LDG.128, 4 distinct records", buf, bytes, iters, out, cyc, blocks);
        run_bcast<8, 1, true>("T  broadcast LDS.64", buf, bytes, iters, out, cyc, blocks);
        run_bcast<16, 1, true>("U  broadcast LDS.128", buf, bytes, iters, out, cyc, blocks);
        run_bcast<16, 4, true>("V  LDS.128, 4 distinct", buf, bytes, iters, out, cyc, blocks);
        CK(cudaFree(buf));
    }
    return 0;
}
