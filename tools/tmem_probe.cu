// tools/tmem_probe.cu -- probe: tensor memory (TMEM) as a lane-private FP64 scratchpad on sm_100a.
// Checks tcgen05.st/ld round trips at every chunk size / column offset the MPC kernel uses, with 4 warps per
// CTA (one lane quarter each) and several CTAs per SM, and times a dependent LDTM chain against LDS.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tmem_probe tools/tmem_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../ac_mpc_b200/csrc/simt.cuh"

using namespace acmpc;

template <int COLS>
__global__ void __launch_bounds__(128) probe(double* out, int* errs, long long* clk)
{
    __shared__ uint32_t tbase;
    __shared__ double sm[4][16 * 32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&tbase)),
                     "n"(COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    Tm t;
    t.a = tbase + ((uint32_t)(32 * w) << 16);
    const int nd = COLS / 2;
    // fill every double column with a lane/warp/cta specific pattern through x32 stores
    for (int d = 0; d < nd; d += 16) {
        double v[16];
        for (int k = 0; k < 16; ++k) v[k] = 1e6 * blockIdx.x + 1e4 * w + 100.0 * lane + (d + k) + 0.25;
        tm_st<16>(t, d, v);
    }
    int bad = 0;
    // read back with every chunk size at odd offsets
    for (int d = 0; d + 16 <= nd; d += 3) {
        double a16[16], a8[8], a4[4], a2[2], a1[1];
        tm_ld<16>(t, d, a16), tm_ld<8>(t, d, a8), tm_ld<4>(t, d, a4), tm_ld<2>(t, d, a2), tm_ld<1>(t, d, a1);
        for (int k = 0; k < 16; ++k) {
            double want = 1e6 * blockIdx.x + 1e4 * w + 100.0 * lane + (d + k) + 0.25;
            if (a16[k] != want) ++bad;
            if (k < 8 && a8[k] != want) ++bad;
            if (k < 4 && a4[k] != want) ++bad;
            if (k < 2 && a2[k] != want) ++bad;
            if (k < 1 && a1[k] != want) ++bad;
        }
    }
    // small stores at odd offsets
    {
        double v1[1] = {-7.5 - lane}, v2[2] = {-1.0 - lane, -2.0 - lane};
        tm_st<1>(t, 5, v1);
        tm_st<2>(t, 9, v2);
        double a[8];
        tm_ld<8>(t, 4, a);
        if (a[1] != -7.5 - lane || a[5] != -1.0 - lane || a[6] != -2.0 - lane) ++bad;
        if (a[0] != 1e6 * blockIdx.x + 1e4 * w + 100.0 * lane + 4 + 0.25) ++bad;
    }
    if (bad) atomicAdd(errs, bad);
    // timing: dependent chain of x32 loads (address depends on the previous value) vs LDS.64 x16
    for (int k = 0; k < 16; ++k) sm[w][k * 32 + lane] = 0.0;
    {
        double z[16];
        for (int k = 0; k < 16; ++k) z[k] = 0.0;
        tm_st<16>(t, 16, z);
    }
    __syncwarp();
    long long t0 = clock64();
    double acc = 0.0;
    int off = 16;
    for (int it = 0; it < 256; ++it) {
        double a[16];
        tm_ld<16>(t, off, a);
        double s = 0;
        for (int k = 0; k < 16; ++k) s += a[k];
        acc += s;
        off = 16 + (int)s;   // s == 0: keeps the chain dependent
    }
    long long t1 = clock64();
    int so = 0;
    for (int it = 0; it < 256; ++it) {
        double s = 0;
        for (int k = 0; k < 16; ++k) s += sm[w][(k + so) * 32 + lane];
        acc += s;
        so = (int)s;
    }
    long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = (t1 - t0) / 256, clk[1] = (t2 - t1) / 256;
    out[blockIdx.x * 128 + threadIdx.x] = acc;
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(COLS));
}

int main()
{
    double* out;
    int* errs;
    long long* clk;
    cudaMalloc(&out, 8 * 128 * 4096);
    cudaMalloc(&errs, 4);
    cudaMalloc(&clk, 16);
    cudaMemset(errs, 0, 4);
    probe<256><<<1184, 128>>>(out, errs, clk);   // 8 waves of 148 SMs, 2 CTAs per SM by TMEM columns
    cudaError_t e = cudaDeviceSynchronize();
    int h = -1;
    long long hc[2];
    cudaMemcpy(&h, errs, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc, clk, 16, cudaMemcpyDeviceToHost);
    printf("probe<256>: %s, mismatches=%d, dependent x32 LDTM+sum iteration %lld clk, 16 LDS.64+sum iteration %lld clk\n",
           cudaGetErrorString(e), h, hc[0], hc[1]);
    cudaMemset(errs, 0, 4);
    probe<128><<<1184, 128>>>(out, errs, clk);
    e = cudaDeviceSynchronize();
    cudaMemcpy(&h, errs, 4, cudaMemcpyDeviceToHost);
    printf("probe<128>: %s, mismatches=%d\n", cudaGetErrorString(e), h);
    return (e != cudaSuccess || h != 0);
}
