// flag_latency_probe.cu -- how long does a flag written by one SM take to be seen by a polling SM?
// Ping-pong between CTA 0 and CTA k of a 148-CTA grid (one CTA per SM), 2000 round trips, for several
// store / load flavours.  Prints cycles and nanoseconds per ONE-WAY hop.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/flag_latency_probe tools/flag_latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ void put(unsigned long long* p, unsigned long long v)
{
    if (MODE == 0) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (MODE == 1) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (MODE == 2) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (MODE == 3) asm volatile("red.relaxed.gpu.global.add.u64 [%0], 1;" ::"l"(p) : "memory");
    if (MODE == 4) asm volatile("st.global.cg.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
template <int MODE>
__device__ __forceinline__ unsigned long long get(unsigned long long* p)
{
    unsigned long long v;
    if (MODE == 0) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (MODE == 1 || MODE == 3) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (MODE == 2) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (MODE == 4) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <int MODE>
__global__ void pingpong(unsigned long long* flags, int peer, int rounds, long long* cycles)
{
    if (threadIdx.x != 0) return;
    unsigned long long* a = flags;        // written by CTA 0
    unsigned long long* b = flags + 32;   // written by the peer (another 128-byte line... 256 B away)
    if (blockIdx.x == 0) {
        const long long t0 = clock64();
        for (int r = 1; r <= rounds; ++r) {
            put<MODE>(a, (unsigned long long)r);
            while (get<MODE>(b) < (unsigned long long)r) {}
        }
        *cycles = clock64() - t0;
    } else if ((int)blockIdx.x == peer) {
        for (int r = 1; r <= rounds; ++r) {
            while (get<MODE>(a) < (unsigned long long)r) {}
            put<MODE>(b, (unsigned long long)r);
        }
    }
}

template <int MODE>
void run(const char* name, unsigned long long* flags, long long* d_cycles, int peer)
{
    const int rounds = 2000;
    cudaMemset(flags, 0, 1024);
    void* args[] = {&flags, &peer, (void*)&rounds, &d_cycles};
    cudaLaunchCooperativeKernel((const void*)pingpong<MODE>, dim3(148), dim3(32), args, 0, 0);
    long long c = 0;
    cudaMemcpy(&c, d_cycles, 8, cudaMemcpyDeviceToHost);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("%-28s peer %3d: %7.0f cycles one way (%.2f us at %d MHz)  %s\n", name, peer, c / (2.0 * rounds),
           c / (2.0 * rounds) / (khz * 1e-3), khz / 1000, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    unsigned long long* flags;
    long long* d_cycles;
    cudaMalloc(&flags, 1024);
    cudaMalloc(&d_cycles, 8);
    for (int peer : {1, 2, 74, 147}) {
        run<0>("st/ld.volatile", flags, d_cycles, peer);
        run<1>("st/ld.relaxed.gpu", flags, d_cycles, peer);
        run<2>("st.release/ld.acquire.gpu", flags, d_cycles, peer);
        run<3>("red.add / ld.relaxed.gpu", flags, d_cycles, peer);
        run<4>("st.cg / ld.cg", flags, d_cycles, peer);
    }
    return 0;
}
