// Throughput of DADD / DMUL / DFMA per SM (one block per SM, 16 warps, 8 independent chains per thread).
// Question behind it: the lookup arithmetic is 11 DADD + 11 DMUL + 2 DFMA per (lookup, nuclide) -- do
// DADD and DMUL issue at the DFMA rate?  (a + b == fma(a, 1, b) and a * b == fma(a, b, -0.0) bit for bit.)
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double *out, int iters, long long *cyc)
{
    double a[8];
    for (int k = 0; k < 8; k++) a[k] = threadIdx.x * 1e-3 + 1.0 + k;
    const double m = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (OP == 0) a[k] = __fma_rn(a[k], m, c);
            if (OP == 1) a[k] = __dadd_rn(a[k], c);
            if (OP == 2) a[k] = __dmul_rn(a[k], m);
            if (OP == 3) a[k] = __fma_rn(a[k], 1.0, c);          // add as fma
            if (OP == 4) a[k] = __fma_rn(a[k], m, -0.0);         // mul as fma
            if (OP == 5) a[k] = __dmul_rn(__dadd_rn(a[k], c), m);   // add then mul (2 ops)
        }
    }
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < 8; k++) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char *name, int ops, double *out, long long *cyc)
{
    long long h; const int iters = 4096;
    for (int warps : {4, 16, 32}) {
        chain<OP><<<148, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-14s warps/SM %2d: %6.3f cycles per warp-instruction per SM\n", name, warps, (double)h / iters / 8 / ops / warps);
    }
}
int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    run<0>("DFMA", 1, out, cyc); run<1>("DADD", 1, out, cyc); run<2>("DMUL", 1, out, cyc);
    run<3>("fma(a,1,c)", 1, out, cyc); run<4>("fma(a,m,-0)", 1, out, cyc); run<5>("DADD+DMUL", 2, out, cyc);
    return 0;
}
