// FP64 dependent-issue latency and per-SM throughput on this GPU (informs the ILP the sorted kernel needs).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain(double *out, int iters, int ilp, long long *cyc)
{
    double a0 = threadIdx.x * 1e-3 + 1.0, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    if (ilp == 1) for (int i = 0; i < iters; i++) { a0 = __fma_rn(a0, m, c); }
    else if (ilp == 2) for (int i = 0; i < iters; i++) { a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); }
    else if (ilp == 4) for (int i = 0; i < iters; i++) { a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c); }
    else for (int i = 0; i < iters; i++) { a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
                                          a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c); }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int warps : {1, 4, 8, 16, 32})
        for (int ilp : {1, 2, 4, 8}) {
            chain<<<148, warps * 32>>>(out, iters, ilp, cyc);   // 1 block per SM
            cudaDeviceSynchronize();
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("warps/SM %2d ilp %d: %6.2f cycles per dependent DFMA step, %6.2f cycles per warp-DFMA per SM\n", warps, ilp,
                   (double)h / iters, (double)h / iters / ilp / warps);
        }
    return 0;
}
