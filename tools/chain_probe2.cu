// chain_probe2.cu -- map_profile.cuh's Grid::chain() lifted into a stand-alone kernel (512 threads per CTA, the other
// 15 warps at __syncthreads, synthetic local work between the chains) to bisect what makes a chain step slow.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/chain_probe2 tools/chain_probe2.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr unsigned kFull = 0xffffffffu;
constexpr int kWarps = 16;

template <int VARIANT>
__device__ __forceinline__ double chain(ulonglong4* slots, unsigned chains, int pos, int ncta, double A, double B, int l)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double Ap = __shfl_up_sync(kFull, A, d), Bp = __shfl_up_sync(kFull, B, d);
        if (l >= d) B = fma(A, Bp, B), A = A * Ap;
    }
    const unsigned long long tag = (unsigned long long)chains << 32;
    ulonglong4* slot = slots + (size_t)(chains & 1u) * 149 * 4;
    if (l == kWarps) {
        const unsigned long long a = (unsigned long long)__double_as_longlong(A), b = (unsigned long long)__double_as_longlong(B);
        unsigned long long* w = (unsigned long long*)(slot + (pos + 1) * 4);
        asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(w), "l"((a & 0xffffffffull) | tag), "l"((a >> 32) | tag) : "memory");
        if (VARIANT != 1)
            asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(w + 2), "l"((b & 0xffffffffull) | tag), "l"((b >> 32) | tag) : "memory");
    }
    __syncwarp();
    const int per = (ncta + 31) >> 5;
    double FA = 1.0, FB = 0.0;
    if (VARIANT == 4) {
        // warp-uniform polling: every lane stays in the loop until ALL lanes have their aggregate
        for (int k = 0; k < per; ++k) {
            const int idx = l * per + k + 1;
            bool ready = idx > pos;
            unsigned long long w0 = 0, w1 = 0, w2 = 0, w3 = 0;
            const unsigned long long* w = (const unsigned long long*)(slot + (ready ? 0 : idx) * 4);
            do {
                if (!ready) {
                    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(w) : "memory");
                    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "l"(w + 2) : "memory");
                    ready = (unsigned)(w0 >> 32) == chains && (unsigned)(w1 >> 32) == chains &&
                            (unsigned)(w2 >> 32) == chains && (unsigned)(w3 >> 32) == chains;
                }
            } while (!__all_sync(kFull, ready));
            if (idx <= pos) {
                const double va = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
                const double vb = __longlong_as_double((long long)((w2 & 0xffffffffull) | (w3 << 32)));
                FB = fma(va, FB, vb), FA = va * FA;
            }
        }
    } else
    for (int k = 0; k < per; ++k) {
        const int idx = l * per + k + 1;
        if (idx <= pos) {
            const unsigned long long* w = (const unsigned long long*)(slot + idx * 4);
            unsigned long long w0, w1, w2 = tag, w3 = tag;
            for (;;) {
                asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(w) : "memory");
                if (VARIANT != 1)
                    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "l"(w + 2) : "memory");
                if ((unsigned)(w0 >> 32) == chains && (unsigned)(w1 >> 32) == chains && (unsigned)(w2 >> 32) == chains &&
                    (unsigned)(w3 >> 32) == chains)
                    break;
            }
            const double va = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
            const double vb = __longlong_as_double((long long)((w2 & 0xffffffffull) | (w3 << 32)));
            FB = fma(va, FB, vb), FA = va * FA;
        }
    }
    if (VARIANT == 3) __syncwarp();
    if (VARIANT != 2) {
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double Ap = __shfl_up_sync(kFull, FA, d), Bp = __shfl_up_sync(kFull, FB, d);
            if (l >= d) FB = fma(FA, Bp, FB), FA = FA * Ap;
        }
    }
    const double Gin = __shfl_sync(kFull, FB, 31);
    return fma(A, Gin, B);
}

template <int VARIANT>
__global__ void __launch_bounds__(512, 1) probe(ulonglong4* slots, int iters, int work, long long* cycles, double* sink)
{
    __shared__ double carry[kWarps];
    const int l = threadIdx.x & 31, warp = threadIdx.x >> 5, ncta = gridDim.x;
    double x = 1.0 + 1e-3 * threadIdx.x;
    unsigned chains = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int dir = 0; dir < 2; ++dir) {
            for (int k = 0; k < work; ++k) x = fma(x, 0.999999, 1e-7);
            __syncthreads();
            ++chains;
            if (warp == 0) {
                const double c = chain<VARIANT>(slots, chains, dir ? ncta - 1 - (int)blockIdx.x : (int)blockIdx.x, ncta, 0.5, x, l);
                if (l < kWarps) carry[l] = c;
            }
            __syncthreads();
            x += 1e-12 * carry[warp];
        }
    }
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
    if (x == 123.0) *sink = x;
}

template <int VARIANT>
void run(const char* name, ulonglong4* slots, long long* d_cycles, double* sink, int ncta, int work)
{
    const int iters = 2000;
    cudaMemset(slots, 0, 2 * 149 * 4 * 32);
    void* args[] = {&slots, (void*)&iters, &work, &d_cycles, &sink};
    cudaLaunchCooperativeKernel((const void*)probe<VARIANT>, dim3(ncta), dim3(512), args, 0, 0);
    long long c[148];
    cudaMemcpy(c, d_cycles, ncta * 8, cudaMemcpyDeviceToHost);
    printf("%-34s ctas %3d work %3d: %7.0f cycles per iteration (2 chains)  %s\n", name, ncta, work, (double)c[0] / iters,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    ulonglong4* slots;
    long long* d_cycles;
    double* sink;
    cudaMalloc(&slots, 2 * 149 * 4 * 32);
    cudaMalloc(&d_cycles, 148 * 8);
    cudaMalloc(&sink, 8);
    for (int ncta : {2, 23, 82})
        for (int work : {0, 100}) {
            run<0>("as in map_profile.cuh", slots, d_cycles, sink, ncta, work);
            run<1>("one 16-byte word pair", slots, d_cycles, sink, ncta, work);
            run<2>("no second scan", slots, d_cycles, sink, ncta, work);
            run<4>("warp-uniform poll loop (vote)", slots, d_cycles, sink, ncta, work);
        }
    return 0;
}
