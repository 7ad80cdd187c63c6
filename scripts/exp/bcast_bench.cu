// Microbenchmark 2: cost of warp-wide loads whose lanes all read the same bytes (lane-per-lookup
// on energy-sorted lookups), with minimal ALU work around the loads.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int W, int DIST, int SPACE>   // SPACE 0 = global (nc), 1 = shared
__global__ void __launch_bounds__(256, 4) bcast(const char *base, uint32_t rec_mask, int iters, unsigned long long *out, long long *cycles)
{
    __shared__ __align__(128) char s_buf[8][2048];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = lane * DIST / 32;
    const unsigned warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int i = lane; i < 256; i += 32) ((double *)s_buf[warp])[i] = i;
    __syncwarp();
    uint32_t r = (warp_global * 2654435761u + slot * 3) & rec_mask;
    unsigned long long acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        r = (r + 977) & rec_mask;
        const char *p = SPACE ? (const char *)s_buf[warp] + (r & 7) * 128 + slot * 128 : base + (size_t)r * 128;
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const char *q = p + (u * W) % 128;
            if (SPACE == 0) {
                if (W == 32) { unsigned long long a, b, c, d; asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(q)); acc ^= a ^ b ^ c ^ d; }
                else if (W == 16) { unsigned long long a, b; asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(q)); acc ^= a ^ b; }
                else if (W == 8) { unsigned long long a; asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(a) : "l"(q)); acc ^= a; }
                else { unsigned a; asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(a) : "l"(q)); acc ^= a; }
            } else {
                const uint32_t sa = (uint32_t)__cvta_generic_to_shared(q);
                if (W == 16) { unsigned long long a, b; asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "r"(sa)); acc ^= a ^ b; }
                else if (W == 8) { unsigned long long a; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(a) : "r"(sa)); acc ^= a; }
                else { unsigned a; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(sa)); acc ^= a; }
            }
        }
    }
    long long t1 = clock64();
    if (acc == 0x123456789abcdefull) out[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int W, int DIST, int SPACE>
void run(const char *buf, size_t buf_bytes, int iters, unsigned long long *out, long long *cyc, int blocks)
{
    uint32_t n_rec = 1; while ((size_t)(n_rec * 2) * 128 + 256 <= buf_bytes) n_rec *= 2;
    bcast<W, DIST, SPACE><<<blocks, 256>>>(buf, n_rec - 1, iters / 8, out, cyc);
    CK(cudaDeviceSynchronize());
    bcast<W, DIST, SPACE><<<blocks, 256>>>(buf, n_rec - 1, iters, out, cyc);
    CK(cudaDeviceSynchronize());
    long long *h = (long long *)malloc(blocks * sizeof(long long));
    cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < blocks; i++) mean += h[i]; mean /= blocks;
    const double ld_per_sm = blocks * 8.0 / 148.0 * iters * 8.0;
    printf("%s.%-3d %2d distinct addresses per warp, buf %6.2f MB: %6.2f cyc/warp-load/SM\n", SPACE ? "LDS" : "LDG", W * 8, DIST, buf_bytes / 1e6, mean / ld_per_sm);
    free(h);
}

int main()
{
    const int blocks = 148 * 4;
    unsigned long long *out; long long *cyc; CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&cyc, blocks * 8));
    for (size_t bytes : {(size_t)32 << 10, (size_t)16 << 20}) {
        char *buf; CK(cudaMalloc(&buf, bytes + 4096)); CK(cudaMemset(buf, 1, bytes + 4096));
        int iters = 4000;
        run<4, 1, 0>(buf, bytes, iters, out, cyc, blocks);  run<8, 1, 0>(buf, bytes, iters, out, cyc, blocks);
        run<16, 1, 0>(buf, bytes, iters, out, cyc, blocks); run<32, 1, 0>(buf, bytes, iters, out, cyc, blocks);
        run<8, 2, 0>(buf, bytes, iters, out, cyc, blocks);  run<16, 2, 0>(buf, bytes, iters, out, cyc, blocks);
        run<8, 4, 0>(buf, bytes, iters, out, cyc, blocks);  run<16, 4, 0>(buf, bytes, iters, out, cyc, blocks);
        run<8, 8, 0>(buf, bytes, iters, out, cyc, blocks);  run<16, 8, 0>(buf, bytes, iters, out, cyc, blocks);
        run<8, 32, 0>(buf, bytes, iters, out, cyc, blocks); run<16, 32, 0>(buf, bytes, iters, out, cyc, blocks);
        run<4, 1, 1>(buf, bytes, iters, out, cyc, blocks);  run<8, 1, 1>(buf, bytes, iters, out, cyc, blocks);
        run<16, 1, 1>(buf, bytes, iters, out, cyc, blocks); run<16, 4, 1>(buf, bytes, iters, out, cyc, blocks);
        run<16, 8, 1>(buf, bytes, iters, out, cyc, blocks);
        CK(cudaFree(buf));
    }
    return 0;
}
