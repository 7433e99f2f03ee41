// tools/seed_probe.cu -- accuracy of MUFU seeds (rsqrt.approx.ftz.f64 / rcp.approx.ftz.f64) + k Newton steps
// against correctly rounded 1/sqrt(x) and 1/x, over the ranges the MPC kernels use them on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/seed_probe tools/seed_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>

__device__ double rsq(double a, int steps)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double h = 0.5 * a;
    for (int k = 0; k < steps; ++k) {
        const double e = fma(-(h * y), y, 0.5);
        y = fma(y, e, y);
    }
    return y;
}
__device__ double rcp(double a, int steps)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    for (int k = 0; k < steps; ++k) {
        const double e = fma(-a, y, 1.0);
        y = fma(y, e, y);
    }
    return y;
}

__global__ void probe(double* out)
{
    // x sweeps [1e-6, 1e6] logarithmically; per thread 4096 samples
    double worst[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int i = 0; i < 4096; ++i) {
        const double t = ((double)tid + (double)i * nt) / (4096.0 * nt);
        const double x = exp(log(1e-6) + t * (log(1e6) - log(1e-6))) * (1.0 + 1e-9 * i);
        const double r0 = 1.0 / sqrt(x), c0 = 1.0 / x;
        for (int s = 0; s < 4; ++s) {
            worst[s] = fmax(worst[s], fabs(rsq(x, s) - r0) / r0);
            worst[4 + s] = fmax(worst[4 + s], fabs(rcp(x, s) - c0) / c0);
        }
    }
    for (int s = 0; s < 8; ++s) {
        for (int o = 16; o > 0; o >>= 1) worst[s] = fmax(worst[s], __shfl_xor_sync(0xffffffffu, worst[s], o));
        if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&out[s], __double_as_longlong(worst[s]));
    }
}

int main()
{
    double* d;
    cudaMalloc(&d, 64);
    cudaMemset(d, 0, 64);
    probe<<<148, 256>>>(d);
    double h[8];
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("max relative error vs correctly rounded, x in [1e-6, 1e6]\n");
    for (int s = 0; s < 4; ++s) printf("  %d Newton steps: rsqrt %.3e   rcp %.3e\n", s, h[s], h[4 + s]);
    return cudaGetLastError() != cudaSuccess;
}
