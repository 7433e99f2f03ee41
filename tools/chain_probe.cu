// chain_probe.cu -- the communication pattern of map_profile.cuh's scan chain without the arithmetic:
// every CTA publishes one flag-in-data aggregate per step and waits for the aggregates of all CTAs before it
// (odd steps: lower CTA ids, even steps: higher ones).  Prints cycles per step pair.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/chain_probe tools/chain_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void chain(unsigned long long* slots, int steps, int stride_words, int work, long long* cycles, double* sink)
{
    // blockDim.x > 32: the other warps wait at __syncthreads around the chain, like the solver's 15 other warps
    if (blockDim.x > 32) {
        const int ncta = gridDim.x;
        double acc = 1.0 + threadIdx.x;
        const long long t0 = clock64();
        for (unsigned step = 1; step <= (unsigned)steps; ++step) {
            for (int k = 0; k < work; ++k) acc = fma(acc, 1.0000001, 1e-9);
            __syncthreads();
            if (threadIdx.x < 32) {
                const int l = threadIdx.x, per = (ncta + 31) >> 5;
                const int pos = (step & 1u) ? (int)blockIdx.x : ncta - 1 - (int)blockIdx.x;
                unsigned long long* buf = slots + (size_t)(step & 1u) * (ncta + 1) * stride_words;
                const unsigned long long tag = (unsigned long long)step << 32;
                if (l == 16) {
                    unsigned long long* w = buf + (size_t)(pos + 1) * stride_words;
                    const unsigned long long a = (unsigned long long)__double_as_longlong(acc);
                    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(w), "l"((a & 0xffffffffull) | tag),
                                 "l"((a >> 32) | tag) : "memory");
                }
                __syncwarp();
                for (int k = 0; k < per; ++k) {
                    const int idx = l * per + k + 1;
                    if (idx <= pos) {
                        const unsigned long long* w = buf + (size_t)idx * stride_words;
                        unsigned long long w0, w1;
                        for (;;) {
                            asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(w) : "memory");
                            if ((unsigned)(w0 >> 32) == step && (unsigned)(w1 >> 32) == step) break;
                        }
                        acc += 1e-30 * __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
                    }
                }
                acc += __shfl_sync(0xffffffffu, acc, 31) * 1e-30;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
        if (acc == 123.0) *sink = acc;
        return;
    }
    const int l = threadIdx.x, ncta = gridDim.x;
    const int per = (ncta + 31) >> 5;
    double acc = 1.0;
    const long long t0 = clock64();
    for (unsigned step = 1; step <= (unsigned)steps; ++step) {
        const int pos = (step & 1u) ? (int)blockIdx.x : ncta - 1 - (int)blockIdx.x;
        unsigned long long* buf = slots + (size_t)(step & 1u) * (ncta + 1) * stride_words;
        for (int k = 0; k < work; ++k) acc = fma(acc, 1.0000001, 1e-9);   // dependent FP64 chain = local compute
        const unsigned long long tag = (unsigned long long)step << 32;
        if (l == 16) {
            unsigned long long* w = buf + (size_t)(pos + 1) * stride_words;
            const unsigned long long a = (unsigned long long)__double_as_longlong(acc);
            asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(w), "l"((a & 0xffffffffull) | tag),
                         "l"((a >> 32) | tag) : "memory");
        }
        __syncwarp();
        for (int k = 0; k < per; ++k) {
            const int idx = l * per + k + 1;
            if (idx <= pos) {
                const unsigned long long* w = buf + (size_t)idx * stride_words;
                unsigned long long w0, w1;
                for (;;) {
                    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(w) : "memory");
                    if ((unsigned)(w0 >> 32) == step && (unsigned)(w1 >> 32) == step) break;
                }
                acc += 1e-30 * __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
            }
        }
        acc += __shfl_sync(0xffffffffu, acc, 31) * 1e-30;
    }
    if (l == 0) cycles[blockIdx.x] = clock64() - t0;
    if (acc == 123.0) *sink = acc;
}

int main()
{
    unsigned long long* slots;
    long long* d_cycles;
    double* sink;
    cudaMalloc(&slots, 2 * 149 * 16 * 8);
    cudaMalloc(&d_cycles, 148 * 8);
    cudaMalloc(&sink, 8);
    const int steps = 4000;
    for (int threads : {32, 512})
        for (int ncta : {2, 23, 82})
            for (int stride_words : {16})
                for (int work : {0, 100}) {
                    cudaMemset(slots, 0, 2 * 149 * 16 * 8);
                    void* args[] = {&slots, (void*)&steps, &stride_words, &work, &d_cycles, &sink};
                    cudaLaunchCooperativeKernel((const void*)chain, dim3(ncta), dim3(threads), args, 0, 0);
                    long long c[148];
                    cudaMemcpy(c, d_cycles, ncta * 8, cudaMemcpyDeviceToHost);
                    printf("threads %3d ctas %3d  slot stride %3d B  local work %3d DFMA: %7.0f cycles per fwd+bwd pair  %s\n",
                           threads, ncta, stride_words * 8, work, 2.0 * c[0] / steps, cudaGetErrorString(cudaGetLastError()));
                }
    return 0;
}
