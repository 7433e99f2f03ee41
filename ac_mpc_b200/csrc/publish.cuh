// publish.cuh -- the caller side of the MPC step, batched over B instances:
//   control/controller.py:257-267   ControlProcess._reference_path: perceived centre line (P,2) float32 -> (H,3)
//   control/controller.py:274-280   _update_shared_memory: the float32 publish of projected_control.T, cum_time and
//                                   current_prediction (perception/shared_memory.py:90-104)
//   control/commands.py:8-38        TemporalCommandSelector (the production command lookup, controller.py:100,112)
//   control/commands.py:41-99       TemporalCommandInterpolator (vectors: tests/test_commands.py:26-58)
// Nothing is "fixed": the selector's index -1 (elapsed time before the first command) wraps to the LAST command exactly as
// Python's negative index does, and all arithmetic is done in the element type of the arrays (float32 for the shared
// memory views: numpy keeps `float32_array - python_float` in float32), without FMA contraction.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace acmpc {
namespace pub {

// (P,2) float32 -> (H,3) float64: rows 0::ds with ds = int(P / H), widths np.linspace(10.0, 6.0, H)
__global__ void reference_paths_kernel(const float* __restrict__ c, int B, int P, int H, int ds, double* __restrict__ paths)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H) return;
    const int b = i / H, k = i - b * H;
    const float* src = c + ((size_t)b * P + (size_t)k * ds) * 2;
    double* dst = paths + (size_t)i * 3;
    dst[0] = (double)src[0];
    dst[1] = (double)src[1];
    // numpy.linspace: arange(num) * step + start, last sample = stop
    const double step = (6.0 - 10.0) / (double)(H - 1);
    dst[2] = k == H - 1 ? 6.0 : __dadd_rn(__dmul_rn((double)k, step), 10.0);
}

// controls [B,2,n] f64 -> control_inputs [B,n,2] f32 ; cum_time [B,n] -> f32 ; prediction [B,n,2] -> f32
__global__ void publish_kernel(const double* __restrict__ controls, const double* __restrict__ cum_time,
                               const double* __restrict__ prediction, int B, int n, float* __restrict__ control_inputs,
                               float* __restrict__ control_cumtime, float* __restrict__ predicted_locations)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * n) return;
    const int b = i / n, k = i - b * n;
    if (controls && control_inputs) {
        control_inputs[(size_t)i * 2] = (float)controls[((size_t)b * 2) * n + k];
        control_inputs[(size_t)i * 2 + 1] = (float)controls[((size_t)b * 2 + 1) * n + k];
    }
    if (cum_time && control_cumtime) control_cumtime[i] = (float)cum_time[i];
    if (prediction && predicted_locations) {
        predicted_locations[(size_t)i * 2] = (float)prediction[(size_t)i * 2];
        predicted_locations[(size_t)i * 2 + 1] = (float)prediction[(size_t)i * 2 + 1];
    }
}

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// one thread per instance.  mode 0: TemporalCommandSelector.get_command, mode 1: TemporalCommandInterpolator.get_command
template <typename T>
__global__ void select_commands_kernel(const T* __restrict__ cum_time, const T* __restrict__ commands,
                                       const double* __restrict__ elapsed, int B, int n, int mode, T* __restrict__ out,
                                       int32_t* __restrict__ index_out)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const T* ct = cum_time + (size_t)b * n;
    const T* cm = commands + (size_t)b * n * 2;
    const T t = (T)elapsed[b];
    // index = np.argmin(abs(cum_time - elapsed_time)): first minimum
    int idx = 0;
    T best = fabs(ct[0] - t);
    for (int k = 1; k < n; ++k) {
        const T a = fabs(ct[k] - t);
        if (a < best) best = a, idx = k;
    }
    const T dist = ct[idx] - t;
    T o0, o1;
    int ia = idx, ib = idx;
    if (mode == 0) {
        if (dist > (T)0) ia -= 1;     // commands.py:33-34
        if (ia >= n) ia = n - 1;      // commands.py:25-26 (never taken)
        if (ia < 0) ia += n;          // Python's negative index: the LAST command
        ib = ia;
        o0 = cm[2 * ia], o1 = cm[2 * ia + 1];
    } else {
        if (idx == 0 || idx == n - 1) ib = idx;
        else if (dist < (T)0) ib = idx + 1;
        else ib = idx - 1;
        if (ia == ib) {
            o0 = cm[2 * ia], o1 = cm[2 * ia + 1];
        } else {
            const T xa = ct[ia], xb = ct[ib];
            const T pa = (xb - t) / (xb - xa), pb = (t - xa) / (xb - xa);
            o0 = add_rn(mul_rn(cm[2 * ia], pa), mul_rn(cm[2 * ib], pb));
            o1 = add_rn(mul_rn(cm[2 * ia + 1], pa), mul_rn(cm[2 * ib + 1], pb));
        }
    }
    out[2 * b] = o0, out[2 * b + 1] = o1;
    if (index_out) index_out[2 * b] = ia, index_out[2 * b + 1] = ib;
}

}   // namespace pub
}   // namespace acmpc
