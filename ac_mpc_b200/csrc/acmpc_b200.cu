// acmpc_b200.cu -- sm_100a kernels + C ABI (include/acmpc_b200.h) of the batched MPC step.
//
// One warp owns one problem instance.  A step is two kernels (three for batches of 1024+ instances):
//   acmpc_order_kernel     bins the instances by their live v_max (longest-first scheduling)
//   acmpc_speed_kernel<C>  waypoints + speed-profile QP, registers only, one warp per CTA
//   acmpc_control_kernel<C> control QP + unpack + rollout + cost; four warps = four instances per CTA; the warp's
//                          (H,3) reference-path slice is staged into shared memory with a TMA bulk copy
//                          (cp.async.bulk + mbarrier), the ADMM iterates stay in registers, the scaled problem and
//                          its factor in the warp's quarter of the CTA's TENSOR MEMORY allocation (a lane-private
//                          FP64 scratchpad through tcgen05.ld/st), the cross-lane data in its slice of shared memory
// The per-instance algorithm is in mpc_warp.cuh; the kernels are instantiated for C = ceil(H/32) = 1..4 horizon
// stages per lane (C = 3: split layout + CTA-phased rounds, C = 4: one CTA per SM with an L1-friendly carve-out).
// Host entry points: zero-copy for B <= 64 and for batches in pinned memory (the kernels read / write the host buffers
// over PCIe), a staged 4-chunk pipeline otherwise.  Multi-GPU: the outputs may be peer-mapped memory of another GPU; the
// control kernel's last CTA then raises a completion flag there (acmpc_attach_completion).
//
// There is NO CPU path in this library: acmpc_create fails with ACMPC_ERR_NO_DEVICE without a GPU.
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>

#include "mpc_warp.cuh"
#include "map_profile.cuh"
#include "publish.cuh"
#include "track_prep.cuh"
#include "model.cuh"

namespace {

struct KernelParams {
    acmpc_config cfg;
    const double* paths;
    const double* offsets;
    const double* vmax;
    double* vel;   // [B,n] speed profile: written by the speed kernel, read by the control kernel
    double* way;   // NULL, or [B,7,n] ReferencePath rows of the stand-alone speed profile (see speed_instance)
    double* cold;  // split layout (C = 3): [launched warps, Layout<C>::kColdDoubles] cold per-stage fields in global memory
    double* warm;  // NULL or [B, Layout<C>::kWarmDoubles] warm-start records (read when use_warm, always rewritten)
    int32_t use_warm;
    acmpc_outputs out;
    int32_t B;
    int32_t is_localised;
    int32_t use_tma;
    // work queue of the persistent warps: a ticket counter that is never reset; the launch that starts at
    // ticket `queue_base` hands out instance (ticket - queue_base) + (warps launched).  Launches of one
    // handle are stream-ordered, and every solved instance draws exactly one ticket, so the host knows
    // the base of the next launch without touching the device.
    uint32_t* queue;
    uint32_t queue_base;
    uint32_t warps_launched;
    int32_t persistent;   // 0: one instance per warp, grid = ceil(B/4) CTAs (warps of a CTA stay in phase)
    // Longest-first scheduling (NULL = instance order).  order = [2 x 8 counters | 4 bins x B ids | 4 bins x B ids]:
    // counters 0..3 / bins "a" order the speed kernel's CTAs by the live v_max (higher limit = more ADMM
    // iterations), counters 4..7 / bins "b" are filled BY the speed kernel with its iteration count class and
    // order the control kernel's work queue (the control QP of a long speed solve is usually long too).
    // Two counter sets alternate between launches: the control kernel zeroes the set of the NEXT launch.
    int32_t* order;
    int32_t order_set;
    // Multi-GPU completion protocol (ShardedMPC, transport "peer"): the outputs of this launch may live in ANOTHER GPU's
    // memory (peer-mapped over NVLink).  `flag` (in the consumer's memory) receives `flag_value` once every store of
    // both kernels is visible system-wide: the last CTA of the control kernel to finish (counted in `done`) writes it.
    // `credit_table[0..credit_n)` = addresses of the producers' credit words, written with `credit_value` when the speed
    // kernel starts (the consumer's own launch hands back the buffers it has finished reading).
    uint32_t* done;
    uint32_t* flag;
    uint32_t flag_value;
    uint32_t credit_value;
    const unsigned long long* credit_table;
    int32_t credit_n;
    // producer side of the credits: no CTA of the speed kernel starts before *credit_wait >= credit_need (a LOCAL word
    // the consumer's launch writes remotely) -- the buffers this launch stores into are free from then on.  Polled in
    // the kernel rather than with a stream wait op, so back-to-back launches stay pipelined.
    const uint32_t* credit_wait;
    uint32_t credit_need;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Stage `bytes` (multiple of 16, both ends 16-byte aligned) global -> shared through the TMA engine.
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* mbar, int lane)
{
    const uint32_t bar = smem_u32(mbar);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_u32(dst)),
            "l"(src), "r"(bytes), "r"(bar)
            : "memory");
    }
    __syncwarp();
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar)
            : "memory");
    }
    // the barrier word is re-initialised for the warp's next instance
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
    __syncwarp();
}

constexpr int kOrderBins = 4;

// item t of a binned order -> instance id
__device__ __forceinline__ int ordered_instance(const int32_t* cnt, const int32_t* bins, int B, int t)
{
    int k = 0;
#pragma unroll
    for (int i = 0; i < kOrderBins - 1; ++i) {
        const int c = cnt[i];
        if (k == i && t >= c) t -= c, k = i + 1;
    }
    return bins[(size_t)k * B + t];
}

// `cnt` = the launch's counter set, `bins` = order + 16
__global__ void acmpc_order_kernel(const double* __restrict__ vmax, int B, double v_lo, double v_hi, int32_t* cnt,
                                   int32_t* bins)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    int k = -1;
    if (b < B) {
        const double f = (vmax[b] - v_lo) / (v_hi - v_lo);
        k = kOrderBins - 1 - (int)(f * kOrderBins);          // highest limit first
        k = k < 0 ? 0 : (k > kOrderBins - 1 ? kOrderBins - 1 : k);
    }
    // one atomic per (warp, bin) instead of one per instance
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < kOrderBins; ++i) {
        const unsigned m = __ballot_sync(0xffffffffu, k == i);
        if (m == 0) continue;
        int base = 0;
        if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(cnt + i, __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (k == i) bins[(size_t)i * B + base + __popc(m & ((1u << lane) - 1))] = b;
    }
}

constexpr int kEventRing = 256;
constexpr int kSlots = 5, kDeviceSlot = 4;
constexpr int kWarpsPerCta = 4;   // one instance per warp; four warps share one tensor-memory allocation

template <int C>
__host__ __device__ constexpr size_t warp_smem_bytes()   // control kernel
{
    return sizeof(double) * (size_t)acmpc::Layout<C>::kDoubles + 16;   // + the warp's TMA mbarrier
}
template <int C>
__host__ __device__ constexpr size_t warp_smem_bytes_speed()   // speed kernel
{
    return sizeof(double) * (size_t)acmpc::Layout<C>::kSpeedDoubles + 16;
}

__device__ __forceinline__ void stage_path(double* raw, const double* src, int H, bool use_tma, uint64_t* mbar, int lane)
{
    if (use_tma) {
        tma_load_1d(raw, src, (uint32_t)(3 * H * sizeof(double)), mbar, lane);
    } else {
        for (int i = lane; i < 3 * H; i += 32) raw[i] = src[i];
        __syncwarp();
    }
}

__device__ __forceinline__ acmpc::InstanceOut slice_outputs(const acmpc_outputs& g, int b, int H)
{
    const int n = H - 1;
    acmpc::InstanceOut o;
    o.controls = g.controls ? g.controls + (size_t)b * 2 * n : nullptr;
    o.prediction = g.prediction ? g.prediction + (size_t)b * 2 * n : nullptr;
    o.cum_time = g.cum_time ? g.cum_time + (size_t)b * n : nullptr;
    o.states = g.states ? g.states + (size_t)b * 3 * H : nullptr;
    o.v_ref = g.v_ref ? g.v_ref + (size_t)b * n : nullptr;
    o.cost = g.cost ? g.cost + b : nullptr;
    o.pri_res = g.pri_res ? g.pri_res + b : nullptr;
    o.dua_res = g.dua_res ? g.dua_res + b : nullptr;
    o.status = g.status ? g.status + b : nullptr;
    o.status_speed = g.status_speed ? g.status_speed + b : nullptr;
    o.iters = g.iters ? g.iters + (size_t)b * 2 : nullptr;
    o.rho_updates = g.rho_updates ? g.rho_updates + (size_t)b * 2 : nullptr;
    o.waypoints = g.waypoints ? g.waypoints + (size_t)b * 7 * n : nullptr;
    o.derived = g.derived ? g.derived + (size_t)b * 3 * (n - 1) : nullptr;
    return o;
}

// Kernel 1: waypoints + speed-profile QP, one warp per instance, registers only (plus 2 KB of scratch):
// small code, high occupancy.  Hands the speed profile to kernel 2 through p.vel ([B,n], = out.v_ref when
// the caller asked for that field).
// OCC = CTAs (= warps) per SM the register allocation aims at: 12 (168 registers) is the faster warp and serves every
// batch that fits in one wave (and batch 1: 0.217 ms p50 against 0.230 ms); 16 (128 registers, more spills) wins once the
// batch needs several waves -- 4096 instances at H = 50: 0.1355 ms against 0.1443 ms (launch(), speed_kernel_for()).
template <int C, int OCC = 12>
__global__ void __launch_bounds__(32, OCC) acmpc_speed_kernel(const __grid_constant__ KernelParams p)
{
    // one warp = one CTA: instances need 25..100+ iterations, and a multi-warp CTA would hold its slots until its
    // slowest warp is done
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    if (p.credit_n > 0 && blockIdx.x == 0 && lane < p.credit_n) {
        uint32_t* w = reinterpret_cast<uint32_t*>(p.credit_table[lane]);
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(w), "r"(p.credit_value) : "memory");
    }
    if (p.credit_wait) {
        if (lane == 0) {
            uint32_t v;
            for (;;) {
                // relaxed is enough: the stores this guards come after a branch on the loaded value, and the word only grows
                asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.credit_wait) : "memory");
                if (v >= p.credit_need) break;
                __nanosleep(200);
            }
        }
        __syncwarp();
    }
    int32_t* cnt = p.order ? p.order + 8 * p.order_set : nullptr;
    const int b = (p.order && p.vmax) ? ordered_instance(cnt, p.order + 16, p.B, blockIdx.x) : (int)blockIdx.x;
    const int H = p.cfg.horizon, n = H - 1;
    acmpc::Ctx<C> c;
    c.S = nullptr, c.HS = nullptr;
    c.W = reinterpret_cast<double*>(smem_raw);
    c.tm.a = 0;
    c.H = H, c.n = n, c.cfg = &p.cfg, c.lane = lane;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(c.W + acmpc::Layout<C>::kSpeedDoubles);
#ifdef ACMPC_PHASE_TIMING
    c.tl = clock64();
#endif
    if (!p.way) stage_path(c.W, p.paths + (size_t)b * 3 * H, H, p.use_tma, mbar, lane);
    const double vmax = p.vmax ? p.vmax[b] : p.cfg.v_max;
    double* wrec = p.warm ? p.warm + (size_t)b * acmpc::Layout<C>::kWarmDoubles : nullptr;
    const int iters = acmpc::speed_instance<C>(c, c.W, vmax, p.is_localised, p.vel ? p.vel + (size_t)b * n : nullptr,
                                               slice_outputs(p.out, b, H), wrec, p.use_warm != 0,
                                               p.way ? p.way + (size_t)b * 7 * n : nullptr);
    if (p.order && lane == 0) {   // class of this solve for the control kernel's queue: most iterations first
        const int per = p.cfg.check_termination > 0 ? p.cfg.check_termination : 25;
        int k = kOrderBins - iters / per;
        k = k < 0 ? 0 : (k > kOrderBins - 1 ? kOrderBins - 1 : k);
        const int pos = atomicAdd(cnt + 4 + k, 1);
        p.order[16 + (size_t)(kOrderBins + k) * p.B + pos] = b;
    }
}

// Kernel 2: control QP + unpack + rollout + cost.
template <int C>
__global__ void __launch_bounds__(32 * kWarpsPerCta, (C == 1 ? 3 : (C <= 3 ? 2 : 1)))
    acmpc_control_kernel(const __grid_constant__ KernelParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tensor memory: warp 0 allocates the CTA's columns, every warp then owns its 32-lane quarter of them
    constexpr uint32_t kCols = acmpc::Layout<C>::kTmemCols;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)),
                     "n"(kCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (p.order && blockIdx.x == 0 && threadIdx.x < 8) p.order[8 * (1 - p.order_set) + threadIdx.x] = 0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // optionally persistent warps: the first instance is the warp's global index, further ones come from the
    // queue, so a warp whose instance converges early does not idle behind its CTA's slowest one
    // C = 3: the hot loop is 1.5x the C = 2 one and no longer fits the instruction cache once the 8 warps of an SM sit in
    // 8 different phases of it (ncu, H = 80: 34 % of the stall samples were `no_instructions`).  There the four warps of a
    // CTA stay IN PHASE instead: the CTA draws four instances at a time and meets at a barrier between rounds, so an SM
    // executes two code positions, not eight -- at the price of waiting for the slowest of four instances.
    constexpr bool kPhased = (C == 3);
    __shared__ uint32_t cta_ticket;
    uint32_t round_base = blockIdx.x * kWarpsPerCta;
    for (int item = blockIdx.x * kWarpsPerCta + warp; kPhased ? (round_base < (uint32_t)p.B) : (item < p.B);) {
        if (kPhased && item >= p.B) {   // a partial last round: this warp has no instance but keeps the barriers
            __syncthreads();
            if (threadIdx.x == 0) cta_ticket = atomicAdd(p.queue, (uint32_t)kWarpsPerCta);
            __syncthreads();
            const uint32_t nx = (cta_ticket - p.queue_base) + p.warps_launched;
            round_base = nx < (uint32_t)p.B ? nx : (uint32_t)p.B;
            item = (int)round_base + warp;
            continue;
        }
        const int b = p.order ? ordered_instance(p.order + 8 * p.order_set + 4, p.order + 16 + (size_t)kOrderBins * p.B,
                                                 p.B, item)
                              : item;
        const int H = p.cfg.horizon, n = H - 1;
        acmpc::Ctx<C> c;
        using L = acmpc::Layout<C>;
        double* base = reinterpret_cast<double*>(smem_raw + (size_t)warp * warp_smem_bytes<C>());
        c.S = L::kColdGlobal ? p.cold + (size_t)(blockIdx.x * kWarpsPerCta + warp) * L::kColdDoubles : base;
        c.W = L::kColdGlobal ? base : base + L::kColdDoubles;
        c.HS = c.W + L::kScratch;
        c.tm.a = tmem_base + ((uint32_t)(32 * warp) << 16);
        c.H = H, c.n = n, c.cfg = &p.cfg, c.lane = lane;
        uint64_t* mbar = reinterpret_cast<uint64_t*>(base + L::kDoubles);
        // the raw path slice lands in the scratch region: it is dead before the first factorisation
#ifdef ACMPC_PHASE_TIMING
        c.tl = clock64();
#endif
        stage_path(c.W, p.paths + (size_t)b * 3 * H, H, p.use_tma, mbar, lane);
        const double offset = p.offsets ? p.offsets[b] : 0.0;
        double* wrec = p.warm ? p.warm + (size_t)b * acmpc::Layout<C>::kWarmDoubles : nullptr;
        acmpc::control_instance<C>(c, c.W, p.vel + (size_t)b * n, offset, slice_outputs(p.out, b, H), wrec,
                                   p.use_warm != 0);
        if (!p.persistent) break;
        if (kPhased) {
            __syncthreads();
            if (threadIdx.x == 0) cta_ticket = atomicAdd(p.queue, (uint32_t)kWarpsPerCta);
            __syncthreads();
            const uint32_t nx = (cta_ticket - p.queue_base) + p.warps_launched;
            round_base = nx < (uint32_t)p.B ? nx : (uint32_t)p.B;
            item = (int)round_base + warp;
            continue;
        }
        uint32_t ticket = 0;
        if (lane == 0) ticket = atomicAdd(p.queue, 1u);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        const uint32_t next = (ticket - p.queue_base) + p.warps_launched;
        item = next < (uint32_t)p.B ? (int)next : p.B;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kCols) : "memory");
    if (p.flag && threadIdx.x == 0) {
        // ONE system-scope fence per CTA: the barrier above orders every thread's output stores (possibly to a peer GPU)
        // before this point and fences are cumulative, so the release below covers them all (the grid-sync pattern).
        // A fence in every thread cost 15 us per launch.
        __threadfence_system();
        const unsigned prev = atomicAdd(p.done, 1u);
        if (prev == gridDim.x - 1) {      // last CTA of the launch: everybody's stores are ordered before this point
            *p.done = 0;                  // the next launch on this slot starts from zero (stream-ordered)
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.flag), "r"(p.flag_value) : "memory");
        }
    }
}

// FP64 FMA throughput probe: 8 independent chains per thread, no memory traffic.
__global__ void fp64_peak_kernel(double* sink, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c), a1 = fma(a1, m, c), a2 = fma(a2, m, c), a3 = fma(a3, m, c);
        a4 = fma(a4, m, c), a5 = fma(a5, m, c), a6 = fma(a6, m, c), a7 = fma(a7, m, c);
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;
}

}  // namespace

struct acmpc_handle {
    acmpc_config cfg;
    int device;
    int sm_count;
    int ctas_per_sm;
    std::string err;
    cudaStream_t stream;     // owned, used by the host entry point (= streams[0])
    cudaStream_t streams[4]; // the host entry point pipelines large batches as up to 4 chunks
    // device arena for the host entry point
    void* d_arena;
    size_t arena_bytes;
    void* h_stage;           // pinned mirror of the arena for small batches (one H2D + one D2H)
    size_t stage_bytes;
    // slots 0..3 = the host entry point's chunk streams, slot 4 (kDeviceSlot) = the device entry point, whose launches
    // run on the caller's stream and must not share a ticket counter / order buffer with the handle's own streams
    int32_t* d_order[5];     // longest-first order buffers, see KernelParams::order
    size_t order_cap[5];     // instances each can hold
    int order_parity[5];     // counter set of the last launch
    int order_on;            // ACMPC_ORDER=0 switches the ordering off
    int order_min;           // smallest batch that is ordered (ACMPC_ORDER_MIN, default 1024)
    void* d_warm;            // warm-start records of the host entry point (keep_warm)
    int warm_B;
    void* d_vel;             // speed-profile hand-over buffer of the device entry point (when v_ref is not requested)
    size_t vel_bytes;
    uint32_t* d_queue;       // 5 ticket counters of the persistent warps (see KernelParams), one per slot
    uint32_t queue_pos[5];   // their values once every launch issued so far has completed
    int profiling;           // record events around the two kernels (acmpc_set_profiling)
    cudaEvent_t* ev;         // 3 * kEventRing events
    int ev_head, ev_count;
    double* d_cold;          // split layout: cold per-stage fields of every resident warp, one region per slot
    size_t cold_stride;      // bytes per slot (0: this horizon keeps them in shared memory)
    uint32_t* d_done;        // "CTAs finished" counters of the completion protocol, one per slot
    // one-shot attachment for the next device-entry launch (acmpc_attach_completion)
    uint32_t* att_flag;
    uint32_t att_flag_value, att_credit_value;
    const unsigned long long* att_credit_table;
    int att_credit_n;
    const uint32_t* att_credit_wait;
    uint32_t att_credit_need;
    int persistent;          // persistent control-kernel warps + work queue (ACMPC_PERSISTENT=0 switches it off)
    int last_launches, last_smem, last_threads, last_ipc;
};

namespace {

bool fail(acmpc_handle* h, cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return false;
    char buf[512];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    if (h) h->err = buf;
    return true;
}

bool valid_config(const acmpc_config* c, std::string* why)
{
    if (c->horizon < ACMPC_MIN_HORIZON || c->horizon > ACMPC_MAX_HORIZON) {
        *why = "horizon out of range";
        return false;
    }
    if (c->max_iter < 1 || c->scaling < 0 || c->check_termination < 0) {
        *why = "bad iteration settings";
        return false;
    }
    if (!(c->rho > 0) || !(c->sigma > 0) || !(c->alpha > 0 && c->alpha < 2)) {
        *why = "bad rho/sigma/alpha";
        return false;
    }
    if (!(c->wheelbase > 0) || !(c->width >= 0)) {
        *why = "bad vehicle constants";
        return false;
    }
    return true;
}

int stages_per_lane(int H) { return (H + 31) / 32; }

size_t smem_bytes_for(int H)   // dynamic shared memory per CTA
{
    switch (stages_per_lane(H)) {
        case 1: return kWarpsPerCta * warp_smem_bytes<1>();
        case 2: return kWarpsPerCta * warp_smem_bytes<2>();
        case 3: return kWarpsPerCta * warp_smem_bytes<3>();
        default: return kWarpsPerCta * warp_smem_bytes<4>();
    }
}

size_t warm_bytes_for(int H)
{
    switch (stages_per_lane(H)) {
        case 1: return sizeof(double) * acmpc::Layout<1>::kWarmDoubles;
        case 2: return sizeof(double) * acmpc::Layout<2>::kWarmDoubles;
        case 3: return sizeof(double) * acmpc::Layout<3>::kWarmDoubles;
        default: return sizeof(double) * acmpc::Layout<4>::kWarmDoubles;
    }
}

// Shared-memory carve-out of the control kernel.  Horizons up to 96 fill an SM's shared memory with their 2-3 CTAs: maximum
// carve-out.  C = 4 (H >= 97) fits ONE CTA (155 KB) and spills registers heavily (84 doubles of iterates per lane): it asks
// only for what that CTA needs, so the rest of the 256 KB stays L1 and catches the spills instead of sending them to L2.
int carveout_percent_for(int H, int smem_per_sm)
{
    if (stages_per_lane(H) < 4) return cudaSharedmemCarveoutMaxShared;
    const size_t need = smem_bytes_for(H) + 2048;
    int pct = (int)((need * 100 + (size_t)smem_per_sm - 1) / (size_t)smem_per_sm);
    return pct > 100 ? 100 : pct;
}

size_t cold_doubles_for(int H)   // per warp, 0 unless the horizon's layout keeps the cold fields in global memory
{
    switch (stages_per_lane(H)) {
        case 3: return acmpc::Layout<3>::kColdGlobal ? acmpc::Layout<3>::kColdDoubles : 0;
        default: return 0;
    }
}

int tmem_cols_for(int H)
{
    switch (stages_per_lane(H)) {
        case 1: return acmpc::Layout<1>::kTmemCols;
        case 2: return acmpc::Layout<2>::kTmemCols;
        case 3: return acmpc::Layout<3>::kTmemCols;
        default: return acmpc::Layout<4>::kTmemCols;
    }
}

const void* kernel_for(int H)   // control kernel
{
    switch (stages_per_lane(H)) {
        case 1: return reinterpret_cast<const void*>(&acmpc_control_kernel<1>);
        case 2: return reinterpret_cast<const void*>(&acmpc_control_kernel<2>);
        case 3: return reinterpret_cast<const void*>(&acmpc_control_kernel<3>);
        default: return reinterpret_cast<const void*>(&acmpc_control_kernel<4>);
    }
}

const void* speed_kernel_for(int H, bool dense = false)
{
    switch (stages_per_lane(H)) {
        case 1: return reinterpret_cast<const void*>(&acmpc_speed_kernel<1>);
        case 2: return dense ? reinterpret_cast<const void*>(&acmpc_speed_kernel<2, 16>)
                             : reinterpret_cast<const void*>(&acmpc_speed_kernel<2>);
        case 3: return reinterpret_cast<const void*>(&acmpc_speed_kernel<3>);
        default: return reinterpret_cast<const void*>(&acmpc_speed_kernel<4>);
    }
}

size_t speed_smem_bytes_for(int H)
{
    switch (stages_per_lane(H)) {
        case 1: return warp_smem_bytes_speed<1>();
        case 2: return warp_smem_bytes_speed<2>();
        case 3: return warp_smem_bytes_speed<3>();
        default: return warp_smem_bytes_speed<4>();
    }
}

// (re)allocate the order buffer of chunk stream `qi` for B instances; synchronous, only when it has to grow
bool ensure_order(acmpc_handle* h, int qi, int B)
{
    if (!h->order_on || B < h->order_min || (size_t)B <= h->order_cap[qi]) return true;
    if (h->d_order[qi]) {
        if (fail(h, cudaDeviceSynchronize(), "cudaDeviceSynchronize")) return false;
        cudaFree(h->d_order[qi]);
    }
    h->d_order[qi] = nullptr, h->order_cap[qi] = 0;
    const size_t ints = 16 + 2 * (size_t)kOrderBins * B;
    if (fail(h, cudaMalloc(&h->d_order[qi], ints * sizeof(int32_t)), "cudaMalloc(order)") ||
        fail(h, cudaMemset(h->d_order[qi], 0, 16 * sizeof(int32_t)), "cudaMemset(order)"))
        return false;
    h->order_cap[qi] = (size_t)B;
    return true;
}

// Two launches on `stream`: the speed-profile kernel, then the control kernel.  `d_vel` [B,n] is the
// hand-over buffer between them.
int launch(acmpc_handle* h, int B, const double* d_paths, const double* d_offsets, const double* d_vmax,
           int is_localised, const acmpc_outputs* d_out, double* d_vel, double* d_warm, int use_warm,
           cudaStream_t stream, int qi = 0)
{
    KernelParams p;
    memset(&p, 0, sizeof(p));
    p.cfg = h->cfg;
    p.paths = d_paths, p.offsets = d_offsets, p.vmax = d_vmax;
    p.out = *d_out;
    p.vel = d_vel;
    p.warm = d_warm, p.use_warm = (d_warm && use_warm) ? 1 : 0;
    p.B = B, p.is_localised = is_localised ? 1 : 0;
    const int H = h->cfg.horizon;
    // TMA bulk copies need 16-byte aligned ends and a size that is a multiple of 16
    p.use_tma = ((3 * H * sizeof(double)) % 16 == 0) && ((reinterpret_cast<uintptr_t>(d_paths) & 15) == 0);
    const size_t smem = smem_bytes_for(H);
    void* args[] = {&p};
    // persistent CTAs: at most what the device holds at once (tensor memory / shared memory allow
    // ctas_per_sm of them per SM), fewer for small batches
    int ctas = (B + kWarpsPerCta - 1) / kWarpsPerCta;
    const int resident = h->sm_count * h->ctas_per_sm;
    p.persistent = h->persistent;
    if (p.persistent && ctas > resident) ctas = resident;
    p.queue = h->d_queue + qi, p.queue_base = h->queue_pos[qi], p.warps_launched = (uint32_t)(ctas * kWarpsPerCta);
    p.cold = h->d_cold ? reinterpret_cast<double*>(reinterpret_cast<char*>(h->d_cold) + (size_t)qi * h->cold_stride) : nullptr;
    if (qi == kDeviceSlot && (h->att_flag || h->att_credit_n > 0 || h->att_credit_wait)) {   // consumed by this launch
        p.done = h->d_done + qi, p.flag = h->att_flag, p.flag_value = h->att_flag_value;
        p.credit_table = h->att_credit_table, p.credit_n = h->att_credit_n, p.credit_value = h->att_credit_value;
        p.credit_wait = h->att_credit_wait, p.credit_need = h->att_credit_need;
        h->att_flag = nullptr, h->att_credit_table = nullptr, h->att_credit_n = 0, h->att_credit_wait = nullptr;
    }
    // longest-first order for batches that run several rounds of the device (see KernelParams::order)
    p.order = nullptr;
    if (h->order_on && B >= h->order_min && h->d_order[qi] && (size_t)B <= h->order_cap[qi]) {
        p.order = h->d_order[qi];
        p.order_set = h->order_parity[qi] ^ 1;
        if (d_vmax) {
            acmpc_order_kernel<<<(B + 255) / 256, 256, 0, stream>>>(d_vmax, B, h->cfg.v_min, h->cfg.v_max,
                                                                    p.order + 8 * p.order_set, p.order + 16);
            h->last_launches += 1;
        }
    }
    cudaEvent_t* ev = nullptr;
    if (h->profiling && h->ev) {
        ev = h->ev + 3 * h->ev_head;
        h->ev_head = (h->ev_head + 1) % kEventRing;
        if (h->ev_count < kEventRing) ++h->ev_count;
        cudaEventRecord(ev[0], stream);
    }
    const bool dense = B > 12 * h->sm_count;   // more instances than one wave of the 12-per-SM variant holds
    if (fail(h, cudaLaunchKernel(speed_kernel_for(H, dense), dim3(B), dim3(32), args, speed_smem_bytes_for(H), stream),
             "speed kernel launch"))
        return ACMPC_ERR_CUDA;
    if (ev) cudaEventRecord(ev[1], stream);
    if (fail(h, cudaLaunchKernel(kernel_for(H), dim3(ctas), dim3(32 * kWarpsPerCta), args, smem, stream), "kernel launch")) {
        h->order_cap[qi] = 0;   // the speed kernel has filled a counter set nobody will reset: rebuild the buffer
        return ACMPC_ERR_CUDA;
    }
    if (ev) cudaEventRecord(ev[2], stream);
    // host-side mirrors of the device counters move only once both kernels are in the stream: a failed launch leaves
    // the ticket base and the counter-set parity where the device still has them
    // every solved instance draws one ticket; the phased C = 3 kernel draws four per CTA round, partial rounds included
    if (p.persistent)
        h->queue_pos[qi] += stages_per_lane(H) == 3 ? (uint32_t)((B + kWarpsPerCta - 1) / kWarpsPerCta * kWarpsPerCta) : (uint32_t)B;
    if (p.order) h->order_parity[qi] = p.order_set;
    h->last_launches += 2, h->last_smem = (int)smem, h->last_threads = 32 * kWarpsPerCta, h->last_ipc = kWarpsPerCta;
    if (fail(h, cudaGetLastError(), "kernel launch")) return ACMPC_ERR_CUDA;
    return ACMPC_OK;
}

// small staged call: copy the inputs in, run `launch`, copy the outputs back, synchronise
struct Staged {
    acmpc_handle* h;
    char* base = nullptr;
    size_t used = 0, cap = 0;
    int32_t rc = ACMPC_OK;
    Staged(acmpc_handle* hh, size_t bytes) : h(hh), cap(bytes + 4096)
    {
        if (fail(h, cudaSetDevice(h->device), "cudaSetDevice") || fail(h, cudaMalloc((void**)&base, cap), "cudaMalloc(staged)"))
            rc = ACMPC_ERR_CUDA;
    }
    ~Staged()
    {
        if (base) cudaFree(base);
    }
    void* in(const void* src, size_t bytes)
    {
        if (rc != ACMPC_OK || !src) return nullptr;
        void* d = base + used;
        used += (bytes + 255) & ~(size_t)255;
        if (fail(h, cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, h->stream), "H2D(staged)")) rc = ACMPC_ERR_CUDA;
        return d;
    }
    void* out(const void* dst, size_t bytes)
    {
        if (rc != ACMPC_OK || !dst) return nullptr;
        void* d = base + used;
        used += (bytes + 255) & ~(size_t)255;
        return d;
    }
    void* scratch(size_t bytes)
    {
        if (rc != ACMPC_OK) return nullptr;
        void* d = base + used;
        used += (bytes + 255) & ~(size_t)255;
        return d;
    }
    void back(void* dst, const void* d, size_t bytes)
    {
        if (rc != ACMPC_OK || !dst) return;
        if (fail(h, cudaMemcpyAsync(dst, d, bytes, cudaMemcpyDeviceToHost, h->stream), "D2H(staged)")) rc = ACMPC_ERR_CUDA;
    }
    int32_t finish(int launches, int threads)
    {
        if (rc == ACMPC_OK && (fail(h, cudaGetLastError(), "publish kernels") || fail(h, cudaStreamSynchronize(h->stream), "publish kernels")))
            rc = ACMPC_ERR_CUDA;
        if (rc == ACMPC_OK) h->last_launches = launches, h->last_smem = 0, h->last_threads = threads, h->last_ipc = 1;
        return rc;
    }
};

// The speed-profile kernel alone, on ReferencePath rows the caller built (SpatialMPC.compute_speed_profile).
int launch_speed_only(acmpc_handle* h, int B, double* d_way, const double* d_vmax, int is_localised, int has_end_vel,
                      double end_vel, double* d_solution, int32_t* d_status, int32_t* d_iters2, int32_t* d_rho2,
                      double* d_warm, int use_warm, cudaStream_t stream)
{
    KernelParams p;
    memset(&p, 0, sizeof(p));
    p.cfg = h->cfg;
    p.cfg.has_end_velocity = has_end_vel ? 1 : 0, p.cfg.end_velocity = has_end_vel ? end_vel : 0.0;
    p.way = d_way, p.vmax = d_vmax, p.vel = d_solution;
    p.out.status_speed = d_status, p.out.iters = d_iters2, p.out.rho_updates = d_rho2;
    p.warm = d_warm, p.use_warm = (d_warm && use_warm) ? 1 : 0;
    p.B = B, p.is_localised = is_localised ? 1 : 0;
    const int H = h->cfg.horizon;
    void* args[] = {&p};
    const bool dense = B > 12 * h->sm_count;
    if (fail(h, cudaLaunchKernel(speed_kernel_for(H, dense), dim3(B), dim3(32), args, speed_smem_bytes_for(H), stream),
             "speed kernel launch") ||
        fail(h, cudaGetLastError(), "speed kernel launch"))
        return ACMPC_ERR_CUDA;
    h->last_launches = 1, h->last_smem = (int)speed_smem_bytes_for(H), h->last_threads = 32, h->last_ipc = 1;
    return ACMPC_OK;
}

// acmpc_outputs is 14 pointers in ABI order; per-instance element count and element size of field f
constexpr int kNumFields = 14;
static_assert(sizeof(acmpc_outputs) == kNumFields * sizeof(void*), "acmpc_outputs: one pointer per field");
size_t field_elems(int f, int H)
{
    const size_t n = (size_t)H - 1;
    switch (f) {
        case 0: case 1: return 2 * n;           // controls, prediction
        case 2: case 4: return n;               // cum_time, v_ref
        case 3: return 3 * (size_t)H;           // states
        case 5: case 6: case 7: return 1;       // cost, pri_res, dua_res
        case 8: case 9: return 1;               // status, status_speed
        case 10: case 11: return 2;             // iters, rho_updates
        case 12: return 7 * n;                  // waypoints
        default: return 3 * (n - 1);            // derived
    }
}
size_t field_esize(int f) { return (f >= 8 && f <= 11) ? 4 : 8; }
void* const& field_ptr(const acmpc_outputs& o, int f) { return reinterpret_cast<void* const*>(&o)[f]; }
void*& field_ptr(acmpc_outputs& o, int f) { return reinterpret_cast<void**>(&o)[f]; }
constexpr int kFieldVref = 4;

// (re)allocate the host entry points' warm-start records: dropped when keep_warm is 0 or the batch size changes
bool ensure_warm(acmpc_handle* h, int B, int keep_warm)
{
    if (!keep_warm || B != h->warm_B) {
        if (h->d_warm) cudaFree(h->d_warm);
        h->d_warm = nullptr, h->warm_B = 0;
    }
    if (keep_warm && !h->d_warm) {
        const size_t wb = (size_t)B * warm_bytes_for(h->cfg.horizon);
        if (fail(h, cudaMalloc(&h->d_warm, wb), "cudaMalloc(warm)")) return false;
        if (fail(h, cudaMemsetAsync(h->d_warm, 0, wb, h->stream), "cudaMemset(warm)")) return false;
        h->warm_B = B;
    }
    return true;
}

}  // namespace

extern "C" {

int32_t acmpc_abi_version(void) { return ACMPC_ABI_VERSION; }

void acmpc_default_config(acmpc_config* c)
{
    memset(c, 0, sizeof(*c));
    c->horizon = 50, c->max_iter = 4000;
    c->v_min = 8.0, c->v_max = 84.0, c->a_min = -1.3, c->a_max = 1.0;
    c->ay_max = 5.5, c->ki_min = 0.005, c->end_velocity = 14.0, c->has_end_velocity = 1;
    c->step_cost[0] = 4e-3, c->step_cost[1] = 5e-2, c->step_cost[2] = 0.0;
    c->r_term[0] = 1e-2, c->r_term[1] = 10.0;
    c->final_cost[0] = 1.0, c->final_cost[1] = 0.0, c->final_cost[2] = 0.1;
    c->wheelbase = 2.65, c->width = 1.99, c->delta_max = 0.30;
    c->input_v_min = 8.0, c->input_v_max = 84.0;
    c->rho = 0.1, c->sigma = 1e-6, c->alpha = 1.6;
    c->eps_abs = 1e-3, c->eps_rel = 1e-3, c->eps_prim_inf = 1e-4, c->eps_dual_inf = 1e-4;
    c->adaptive_rho_tolerance = 5.0;
    c->scaling = 10, c->check_termination = 25, c->adaptive_rho = 1, c->adaptive_rho_interval = 50;
    c->check_dualgap = 0;   // OSQP 0.6.x termination test
}

int32_t acmpc_create(const acmpc_config* cfg, int32_t device, acmpc_handle** out)
{
    if (!cfg || !out) return ACMPC_ERR_INVALID;
    *out = nullptr;
    std::string why;
    if (!valid_config(cfg, &why)) return ACMPC_ERR_INVALID;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return ACMPC_ERR_NO_DEVICE;
    }
    acmpc_handle* h = new (std::nothrow) acmpc_handle();
    if (!h) return ACMPC_ERR_INVALID;
    h->cfg = *cfg;
    h->device = device;
    h->d_arena = nullptr, h->arena_bytes = 0;
    h->h_stage = nullptr, h->stage_bytes = 0;
    h->last_launches = h->last_smem = h->last_threads = h->last_ipc = 0;
    cudaDeviceProp prop;
    if (fail(h, cudaSetDevice(device), "cudaSetDevice") ||
        fail(h, cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) {
        delete h;
        return ACMPC_ERR_CUDA;
    }
    if (prop.major != 10) {   // the fatbin only holds sm_100a code
        delete h;
        return ACMPC_ERR_NO_DEVICE;
    }
    h->sm_count = prop.multiProcessorCount;
    h->d_queue = nullptr, h->d_done = nullptr, h->d_cold = nullptr, h->cold_stride = 0;
    h->att_flag = nullptr, h->att_credit_table = nullptr, h->att_credit_n = 0, h->att_flag_value = h->att_credit_value = 0;
    h->att_credit_wait = nullptr, h->att_credit_need = 0;
    for (int k = 0; k < 4; ++k) h->streams[k] = nullptr;
    for (int k = 0; k < kSlots; ++k) h->queue_pos[k] = 0;
    h->d_vel = nullptr, h->vel_bytes = 0;
    h->d_warm = nullptr, h->warm_B = 0;
    for (int k = 0; k < kSlots; ++k) h->d_order[k] = nullptr, h->order_cap[k] = 0, h->order_parity[k] = 0;
    {
        const char* e = getenv("ACMPC_ORDER");
        h->order_on = (e && e[0] == '0') ? 0 : 1;
        const char* m = getenv("ACMPC_ORDER_MIN");
        h->order_min = m ? atoi(m) : 1024;
        if (h->order_min < 1) h->order_min = 1;
    }
    h->profiling = 0, h->ev = nullptr, h->ev_head = 0, h->ev_count = 0;
    {
        const char* e = getenv("ACMPC_PERSISTENT");
        h->persistent = (e && e[0] == '0') ? 0 : 1;
    }
    const size_t smem = smem_bytes_for(cfg->horizon);
    if (smem > (size_t)prop.sharedMemPerBlockOptin ||
        fail(h, cudaFuncSetAttribute(kernel_for(cfg->horizon), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
             "cudaFuncSetAttribute") ||
        fail(h, cudaFuncSetAttribute(speed_kernel_for(cfg->horizon), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)speed_smem_bytes_for(cfg->horizon)),
             "cudaFuncSetAttribute(speed)") ||
        fail(h, cudaFuncSetAttribute(speed_kernel_for(cfg->horizon, true), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)speed_smem_bytes_for(cfg->horizon)),
             "cudaFuncSetAttribute(speed, dense)") ||
        fail(h, cudaFuncSetAttribute(kernel_for(cfg->horizon), cudaFuncAttributePreferredSharedMemoryCarveout,
                                     carveout_percent_for(cfg->horizon, (int)prop.sharedMemPerMultiprocessor)),
             "cudaFuncSetAttribute(carveout)") ||
        fail(h, cudaMalloc(&h->d_queue, kSlots * sizeof(uint32_t)), "cudaMalloc(queue)") ||
        fail(h, cudaMemset(h->d_queue, 0, kSlots * sizeof(uint32_t)), "cudaMemset(queue)") ||
        fail(h, cudaMalloc(&h->d_done, kSlots * sizeof(uint32_t)), "cudaMalloc(done)") ||
        fail(h, cudaMemset(h->d_done, 0, kSlots * sizeof(uint32_t)), "cudaMemset(done)") ||
        fail(h, cudaStreamCreateWithFlags(&h->streams[0], cudaStreamNonBlocking), "cudaStreamCreate") ||
        fail(h, cudaStreamCreateWithFlags(&h->streams[1], cudaStreamNonBlocking), "cudaStreamCreate") ||
        fail(h, cudaStreamCreateWithFlags(&h->streams[2], cudaStreamNonBlocking), "cudaStreamCreate") ||
        fail(h, cudaStreamCreateWithFlags(&h->streams[3], cudaStreamNonBlocking), "cudaStreamCreate")) {
        if (h->d_queue) cudaFree(h->d_queue);
        if (h->d_done) cudaFree(h->d_done);
        delete h;
        return ACMPC_ERR_CUDA;
    }
    h->stream = h->streams[0];
    // CTAs resident per SM = min over shared memory, registers and tensor memory (512 columns per SM).
    // (cudaOccupancyMaxActiveBlocksPerMultiprocessor assumes the default carve-out and under-reports.)
    {
        cudaFuncAttributes fa;
        if (fail(h, cudaFuncGetAttributes(&fa, kernel_for(cfg->horizon)), "cudaFuncGetAttributes")) {
            cudaFree(h->d_queue);
            for (int k = 0; k < 4; ++k) cudaStreamDestroy(h->streams[k]);
            delete h;
            return ACMPC_ERR_CUDA;
        }
        const size_t per_cta = smem + fa.sharedSizeBytes + prop.reservedSharedMemPerBlock;
        const int by_smem = (int)(prop.sharedMemPerMultiprocessor / per_cta);
        const int regs_per_cta = ((fa.numRegs + 7) / 8) * 8 * 32 * kWarpsPerCta;
        const int by_regs = prop.regsPerMultiprocessor / regs_per_cta;
        const int by_tmem = 512 / tmem_cols_for(cfg->horizon);
        int r = by_smem < by_regs ? by_smem : by_regs;
        if (by_tmem < r) r = by_tmem;
        h->ctas_per_sm = r < 1 ? 1 : r;
    }
    // split layout: the cold per-stage fields of every RESIDENT warp live in global memory (the kernel is persistent, so
    // at most sm_count * ctas_per_sm CTAs are ever launched); one region per work-queue slot, launches of different
    // slots may overlap
    if (cold_doubles_for(cfg->horizon) > 0) {
        h->cold_stride = (size_t)h->sm_count * h->ctas_per_sm * kWarpsPerCta * cold_doubles_for(cfg->horizon) * sizeof(double);
        if (fail(h, cudaMalloc(&h->d_cold, kSlots * h->cold_stride), "cudaMalloc(cold fields)")) {
            acmpc_destroy(h);
            return ACMPC_ERR_CUDA;
        }
        h->persistent = 1;
    }
    *out = h;
    return ACMPC_OK;
}

int32_t acmpc_destroy(acmpc_handle* h)
{
    if (!h) return ACMPC_OK;
    cudaSetDevice(h->device);
    if (h->d_arena) cudaFree(h->d_arena);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    if (h->d_queue) cudaFree(h->d_queue);
    if (h->d_done) cudaFree(h->d_done);
    if (h->d_cold) cudaFree(h->d_cold);
    if (h->d_vel) cudaFree(h->d_vel);
    if (h->d_warm) cudaFree(h->d_warm);
    for (int k = 0; k < kSlots; ++k)
        if (h->d_order[k]) cudaFree(h->d_order[k]);
    if (h->ev) {
        for (int i = 0; i < 3 * kEventRing; ++i) cudaEventDestroy(h->ev[i]);
        delete[] h->ev;
    }
    for (int k = 0; k < 4; ++k)
        if (h->streams[k]) cudaStreamDestroy(h->streams[k]);
    delete h;
    return ACMPC_OK;
}

const char* acmpc_last_error(const acmpc_handle* h) { return h ? h->err.c_str() : "null handle"; }

int64_t acmpc_warm_stride(const acmpc_handle* h) { return h ? (int64_t)warm_bytes_for(h->cfg.horizon) : 0; }

int32_t acmpc_solve_batch_device(acmpc_handle* h, int32_t B, const double* d_paths, const double* d_offsets,
                                 const double* d_vmax, int32_t is_localised, void* d_warm, int32_t warm_valid,
                                 const acmpc_outputs* d_out, void* stream)
{
    if (!h) return ACMPC_ERR_INVALID;
    if (B < 0 || !d_out || (B > 0 && !d_paths)) {
        h->err = "bad arguments";
        return ACMPC_ERR_INVALID;
    }
    if (warm_valid && !d_warm) {
        h->err = "warm_valid without a warm-start buffer";
        return ACMPC_ERR_INVALID;
    }
    if (d_warm && (reinterpret_cast<uintptr_t>(d_warm) & 7)) {
        h->err = "warm-start buffer must be 8-byte aligned";
        return ACMPC_ERR_INVALID;
    }
    if (B == 0) return ACMPC_OK;
    if (fail(h, cudaSetDevice(h->device), "cudaSetDevice")) return ACMPC_ERR_CUDA;
    h->last_launches = 0;
    if (!ensure_order(h, kDeviceSlot, B)) return ACMPC_ERR_CUDA;
    double* d_vel = d_out->v_ref;
    if (!d_vel) {   // the caller did not ask for v_ref: hand over through a scratch buffer owned by the handle
        const size_t need = (size_t)B * (h->cfg.horizon - 1) * sizeof(double);
        if (need > h->vel_bytes) {
            if (h->d_vel) {
                if (fail(h, cudaDeviceSynchronize(), "cudaDeviceSynchronize")) return ACMPC_ERR_CUDA;
                cudaFree(h->d_vel);
            }
            h->d_vel = nullptr, h->vel_bytes = 0;
            if (fail(h, cudaMalloc(&h->d_vel, need), "cudaMalloc(vel)")) return ACMPC_ERR_CUDA;
            h->vel_bytes = need;
        }
        d_vel = static_cast<double*>(h->d_vel);
    }
    return launch(h, B, d_paths, d_offsets, d_vmax, is_localised, d_out, d_vel, static_cast<double*>(d_warm),
                  warm_valid, static_cast<cudaStream_t>(stream), kDeviceSlot);
}

int32_t acmpc_solve_batch_host(acmpc_handle* h, int32_t B, const double* paths, const double* offsets,
                               const double* vmax, int32_t is_localised, int32_t keep_warm,
                               const acmpc_outputs* out)
{
    if (!h) return ACMPC_ERR_INVALID;
    if (B < 0 || !out || (B > 0 && !paths)) {
        h->err = "bad arguments";
        return ACMPC_ERR_INVALID;
    }
    if (B == 0) return ACMPC_OK;
    if (fail(h, cudaSetDevice(h->device), "cudaSetDevice")) return ACMPC_ERR_CUDA;
    // keep_warm: the handle plays the reference's persistent solver objects -- one zero-initialised record
    // per instance slot, dropped (cold restart) when the batch size changes or keep_warm is 0
    if (!ensure_warm(h, B, keep_warm)) return ACMPC_ERR_CUDA;
    const int H = h->cfg.horizon;
    const size_t nb = (size_t)B;
    // arena layout (all 256-byte aligned): inputs then one slab per output field
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align(off + bytes); return o; };
    const size_t o_paths = take(nb * 3 * H * 8), o_off = take(nb * 8), o_vmax = take(nb * 8);
    size_t o_f[kNumFields], per_b[kNumFields];
    for (int f = 0; f < kNumFields; ++f) per_b[f] = field_elems(f, H) * field_esize(f), o_f[f] = take(nb * per_b[f]);
    if (off > h->arena_bytes) {
        for (int k = 0; k < 4; ++k)   // an earlier call's copies are done (every call ends with a sync); be explicit
            if (fail(h, cudaStreamSynchronize(h->streams[k]), "cudaStreamSynchronize")) return ACMPC_ERR_CUDA;
        if (h->d_arena) cudaFree(h->d_arena);
        h->d_arena = nullptr, h->arena_bytes = 0;
        if (fail(h, cudaMalloc(&h->d_arena, off), "cudaMalloc(arena)")) return ACMPC_ERR_CUDA;
        h->arena_bytes = off;
    }
    char* base = static_cast<char*>(h->d_arena);
    const size_t wstride = warm_bytes_for(H);
    h->last_launches = 0;
    // device-side output struct of the instances [z, z + ...): requested fields only (v_ref is always the hand-over)
    auto device_outputs = [&](size_t z) {
        acmpc_outputs d;
        memset(&d, 0, sizeof(d));
        for (int f = 0; f < kNumFields; ++f)
            if (field_ptr(*out, f)) field_ptr(d, f) = base + o_f[f] + z * per_b[f];
        return d;
    };
    // Small batches (the B = 1 drop-in call) are ZERO-COPY: inputs and outputs live in ONE pinned staging buffer that
    // mirrors the arena layout and the kernels read / write it directly over PCIe (UVA: pinned host memory is device
    // accessible under the same pointer) -- no cudaMemcpy at all, two launches and one stream synchronisation.  An
    // instance moves 1.2 KB in and 2.4-6 KB out, far too little for a copy engine round trip to pay (the two staged
    // copies cost ~20 us of a 0.24 ms call).  Only the speed-profile hand-over between the kernels stays in the arena.
    if (B <= 64) {
        if (off > h->stage_bytes) {
            if (fail(h, cudaStreamSynchronize(h->streams[0]), "cudaStreamSynchronize")) return ACMPC_ERR_CUDA;
            if (h->h_stage) cudaFreeHost(h->h_stage);
            h->h_stage = nullptr, h->stage_bytes = 0;
            if (fail(h, cudaHostAlloc(&h->h_stage, off, cudaHostAllocMapped), "cudaHostAlloc(stage)")) return ACMPC_ERR_CUDA;
            h->stage_bytes = off;
        }
        char* st = static_cast<char*>(h->h_stage);
        cudaStream_t s = h->streams[0];
        memcpy(st + o_paths, paths, nb * 3 * H * 8);
        if (offsets) memcpy(st + o_off, offsets, nb * 8);
        if (vmax) memcpy(st + o_vmax, vmax, nb * 8);
        acmpc_outputs d;
        memset(&d, 0, sizeof(d));
        for (int f = 0; f < kNumFields; ++f)
            if (field_ptr(*out, f)) field_ptr(d, f) = st + o_f[f];
        int rc = launch(h, B, reinterpret_cast<const double*>(st + o_paths),
                        offsets ? reinterpret_cast<const double*>(st + o_off) : nullptr,
                        vmax ? reinterpret_cast<const double*>(st + o_vmax) : nullptr, is_localised, &d,
                        reinterpret_cast<double*>(base + o_f[kFieldVref]), static_cast<double*>(h->d_warm), 1, s, 0);
        if (rc != ACMPC_OK) return rc;
        if (fail(h, cudaStreamSynchronize(s), "cudaStreamSynchronize")) return ACMPC_ERR_CUDA;
        for (int f = 0; f < kNumFields; ++f)
            if (field_ptr(*out, f)) memcpy(field_ptr(*out, f), st + o_f[f], nb * per_b[f]);
        return ACMPC_OK;
    }
    // Large batches whose buffers are ALL pinned host memory (cudaHostAlloc / cudaHostRegister: what a control loop that
    // cares about latency uses, and what bench.py's e2e leg passes) are zero-copy as well: one launch over the whole
    // batch, the speed kernel pulls the paths over PCIe (4.9 MB per 4096 instances, hidden behind its 16 warps per SM),
    // the kernels push their results straight into the caller's arrays (23 GB/s at 4096 instances per 0.42 ms), and the
    // call ends with one stream synchronisation.  ACMPC_ZEROCOPY=0 forces the staged, chunk-pipelined path below.
    {
        static const int zc_env = [] { const char* e = getenv("ACMPC_ZEROCOPY"); return (e && e[0] == '0') ? 0 : 1; }();
        auto pinned = [](const void* ptr) {
            if (!ptr) return true;
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
                cudaGetLastError();
                return false;
            }
            return at.type == cudaMemoryTypeHost;
        };
        bool zc = zc_env && pinned(paths) && pinned(offsets) && pinned(vmax);
        for (int f = 0; zc && f < kNumFields; ++f) zc = pinned(field_ptr(*out, f));
        if (zc) {
            cudaStream_t s = h->streams[0];
            if (!ensure_order(h, 0, B)) return ACMPC_ERR_CUDA;
            int rc = launch(h, B, paths, offsets, vmax, is_localised, out,
                            reinterpret_cast<double*>(base + o_f[kFieldVref]), static_cast<double*>(h->d_warm), 1, s, 0);
            if (rc != ACMPC_OK) return rc;
            if (fail(h, cudaStreamSynchronize(s), "cudaStreamSynchronize")) return ACMPC_ERR_CUDA;
            return ACMPC_OK;
        }
    }
    // Large batches go through as up to 4 chunks on 4 streams (own ticket queue each), so that the H2D copy of
    // one chunk and the D2H copy of another overlap the kernels of a third.  (Async only from pinned host memory.)
    // 2 chunks from 512 instances, 4 from 2048 (measured at 4096: 1 / 2 / 3 / 4 / 6 / 8 chunks = 4.31 / 5.04 / 5.12 / 5.36 /
    // 4.70 / 4.35 M solves/s end to end), then chunks of at most 16384 instances dealt round-robin to the 4 streams
    // (a 1 M-instance sweep would otherwise wait for a 300 MB H2D before its first kernel)
    int chunks = B >= 2048 ? 4 : (B >= 512 ? 2 : 1);
    if (B > 4 * 16384) chunks = (B + 16383) / 16384;
    int per = ((B + chunks - 1) / chunks + kWarpsPerCta - 1) / kWarpsPerCta * kWarpsPerCta;
    // experiment switch: instances per chunk.  At 4096 (r1q): 4 x 1024 = 5.50 M solves/s end to end; 1184 (one full wave of
    // the control kernel) + a 544 tail 5.34 M; 1216 5.27 M; 1100 5.27 M; 896 (5 chunks) 4.99 M
    if (const char* e = getenv("ACMPC_CHUNK_PER")) {
        const int v = atoi(e) / kWarpsPerCta * kWarpsPerCta;
        if (v >= kWarpsPerCta && chunks > 1) per = v, chunks = (B + per - 1) / per;
    }
    if (h->d_warm && chunks > 1 &&
        fail(h, cudaStreamSynchronize(h->streams[0]), "cudaStreamSynchronize"))   // the zero-fill of new records
        return ACMPC_ERR_CUDA;
    for (int k = 0; k < chunks; ++k) {
        const int c0 = k * per;
        const int cb = (c0 + per <= B) ? per : B - c0;
        if (cb <= 0) break;
        const int qi = k & 3;
        cudaStream_t s = h->streams[qi];
        const size_t z = (size_t)c0, nbk = (size_t)cb;
        if (!ensure_order(h, qi, per)) return ACMPC_ERR_CUDA;
        if (fail(h, cudaMemcpyAsync(base + o_paths + z * 3 * H * 8, paths + z * 3 * H, nbk * 3 * H * 8,
                                    cudaMemcpyHostToDevice, s), "H2D paths"))
            return ACMPC_ERR_CUDA;
        if (offsets && fail(h, cudaMemcpyAsync(base + o_off + z * 8, offsets + z, nbk * 8, cudaMemcpyHostToDevice, s),
                            "H2D offsets"))
            return ACMPC_ERR_CUDA;
        if (vmax && fail(h, cudaMemcpyAsync(base + o_vmax + z * 8, vmax + z, nbk * 8, cudaMemcpyHostToDevice, s),
                         "H2D vmax"))
            return ACMPC_ERR_CUDA;
        const acmpc_outputs d = device_outputs(z);
        int rc = launch(h, cb, reinterpret_cast<const double*>(base + o_paths) + z * 3 * H,
                        offsets ? reinterpret_cast<const double*>(base + o_off) + z : nullptr,
                        vmax ? reinterpret_cast<const double*>(base + o_vmax) + z : nullptr, is_localised, &d,
                        reinterpret_cast<double*>(base + o_f[kFieldVref]) + z * (H - 1),
                        h->d_warm ? reinterpret_cast<double*>(static_cast<char*>(h->d_warm) + z * wstride) : nullptr, 1, s, qi);
        if (rc != ACMPC_OK) return rc;
        for (int f = 0; f < kNumFields; ++f)
            if (field_ptr(*out, f) &&
                fail(h, cudaMemcpyAsync(static_cast<char*>(field_ptr(*out, f)) + z * per_b[f], field_ptr(d, f), nbk * per_b[f],
                                        cudaMemcpyDeviceToHost, s), "D2H outputs"))
                return ACMPC_ERR_CUDA;
    }
    for (int k = 0; k < 4; ++k)
        if (fail(h, cudaStreamSynchronize(h->streams[k]), "cudaStreamSynchronize")) return ACMPC_ERR_CUDA;
    return ACMPC_OK;
}

int32_t acmpc_speed_profile_batch_device(acmpc_handle* h, int32_t B, double* d_waypoints, const double* d_vmax,
                                         int32_t is_localised, int32_t has_end_vel, double end_vel, void* d_warm,
                                         int32_t warm_valid, double* d_solution, int32_t* d_status, int32_t* d_iters,
                                         int32_t* d_rho_updates, void* stream)
{
    if (!h) return ACMPC_ERR_INVALID;
    if (B < 0 || (B > 0 && !d_waypoints) || (warm_valid && !d_warm) || (d_warm && (reinterpret_cast<uintptr_t>(d_warm) & 7))) {
        h->err = "bad arguments";
        return ACMPC_ERR_INVALID;
    }
    if (B == 0) return ACMPC_OK;
    if (fail(h, cudaSetDevice(h->device), "cudaSetDevice")) return ACMPC_ERR_CUDA;
    return launch_speed_only(h, B, d_waypoints, d_vmax, is_localised, has_end_vel, end_vel, d_solution, d_status, d_iters,
                             d_rho_updates, static_cast<double*>(d_warm), warm_valid, static_cast<cudaStream_t>(stream));
}

int32_t acmpc_speed_profile_batch_host(acmpc_handle* h, int32_t B, double* waypoints, const double* vmax,
                                       int32_t is_localised, int32_t has_end_vel, double end_vel, int32_t keep_warm,
                                       double* solution, int32_t* status, int32_t* iters, int32_t* rho_updates)
{
    if (!h) return ACMPC_ERR_INVALID;
    if (B < 0 || (B > 0 && !waypoints)) {
        h->err = "bad arguments";
        return ACMPC_ERR_INVALID;
    }
    if (B == 0) return ACMPC_OK;
    if (fail(h, cudaSetDevice(h->device), "cudaSetDevice")) return ACMPC_ERR_CUDA;
    if (!ensure_warm(h, B, keep_warm)) return ACMPC_ERR_CUDA;
    const int n = h->cfg.horizon - 1;
    const size_t nb = (size_t)B, way_b = nb * 7 * n * 8;
    Staged s(h, way_b + nb * (8 + (size_t)n * 8 + 4 + 8 + 8) + 8 * 256);
    double* d_way = (double*)s.in(waypoints, way_b);
    const double* d_vm = (const double*)s.in(vmax, nb * 8);
    double* d_sol = (double*)s.scratch(nb * n * 8);
    int32_t* d_st = (int32_t*)s.scratch(nb * 4);
    int32_t* d_it = (int32_t*)s.scratch(nb * 8);      // the kernels write [B,2] (speed QP, control QP) pairs
    int32_t* d_ru = (int32_t*)s.scratch(nb * 8);
    if (s.rc == ACMPC_OK) {
        const int32_t rc = launch_speed_only(h, B, d_way, d_vm, is_localised, has_end_vel, end_vel, d_sol, d_st, d_it, d_ru,
                                             static_cast<double*>(h->d_warm), 1, h->stream);
        if (rc != ACMPC_OK) return rc;
    }
    // only the velocities row can have changed
    std::vector<int32_t> it2(iters || rho_updates ? 2 * nb : 0), ru2(rho_updates ? 2 * nb : 0);
    if (s.rc == ACMPC_OK &&
        fail(h, cudaMemcpy2DAsync(waypoints + 6 * (size_t)n, 7 * (size_t)n * 8, d_way + 6 * (size_t)n, 7 * (size_t)n * 8,
                                  (size_t)n * 8, nb, cudaMemcpyDeviceToHost, h->stream), "D2H(velocities)"))
        s.rc = ACMPC_ERR_CUDA;
    s.back(solution, d_sol, nb * n * 8);
    s.back(status, d_st, nb * 4);
    if (iters) s.back(it2.data(), d_it, nb * 8);
    if (rho_updates) s.back(ru2.data(), d_ru, nb * 8);
    const int32_t rc = s.finish(1, 32);
    if (rc == ACMPC_OK)
        for (size_t b = 0; b < nb; ++b) {
            if (iters) iters[b] = it2[2 * b];
            if (rho_updates) rho_updates[b] = ru2[2 * b];
        }
    return rc;
}

int32_t acmpc_attach_completion(acmpc_handle* h, uint32_t* d_flag, uint32_t flag_value, const uint64_t* d_credit_table,
                                int32_t credit_n, uint32_t credit_value, const uint32_t* d_credit_wait, uint32_t credit_need)
{
    if (!h || credit_n < 0 || credit_n > 32 || (credit_n > 0 && !d_credit_table)) return ACMPC_ERR_INVALID;
    h->att_credit_wait = d_credit_wait, h->att_credit_need = credit_need;
    h->att_flag = d_flag, h->att_flag_value = flag_value;
    h->att_credit_table = reinterpret_cast<const unsigned long long*>(d_credit_table), h->att_credit_n = credit_n;
    h->att_credit_value = credit_value;
    return ACMPC_OK;
}

int32_t acmpc_stream_wait_value32(acmpc_handle* h, const uint32_t* d_addr, uint32_t value, void* stream)
{
    if (!h || !d_addr || (reinterpret_cast<uintptr_t>(d_addr) & 3)) return ACMPC_ERR_INVALID;
    if (fail(h, cudaSetDevice(h->device), "cudaSetDevice")) return ACMPC_ERR_CUDA;
    // cuStreamWaitValue32 (driver API) through the runtime's entry-point query: the library does not link libcuda, so it
    // still loads on a box without a driver (the CPU-side tests)
    typedef int (*wait_fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
    static wait_fn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (fail(h, cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &st), "cudaGetDriverEntryPoint") ||
            st != cudaDriverEntryPointSuccess || !f) {
            h->err = "cuStreamWaitValue32 is not available from this driver";
            return ACMPC_ERR_CUDA;
        }
        fn = reinterpret_cast<wait_fn>(f);
    }
    const int rc = fn(static_cast<cudaStream_t>(stream), (unsigned long long)reinterpret_cast<uintptr_t>(d_addr), value,
                      0x0 /* CU_STREAM_WAIT_VALUE_GEQ */);
    if (rc != 0) {
        h->err = "cuStreamWaitValue32 failed with CUresult " + std::to_string(rc);
        return ACMPC_ERR_CUDA;
    }
    return ACMPC_OK;
}

#ifdef ACMPC_PHASE_TIMING
// experiment builds only: read and reset the phase clocks of the control kernel
extern "C" int32_t acmpc_exp_phase_cycles(unsigned long long* out16)
{
    if (cudaMemcpyFromSymbol(out16, acmpc::g_phase, 16 * sizeof(unsigned long long)) != cudaSuccess) return ACMPC_ERR_CUDA;
    unsigned long long z[16] = {0};
    return cudaMemcpyToSymbol(acmpc::g_phase, z, sizeof(z)) == cudaSuccess ? ACMPC_OK : ACMPC_ERR_CUDA;
}
#endif

int32_t acmpc_last_launch_info(const acmpc_handle* h, int32_t* n_launches, int32_t* smem_bytes,
                               int32_t* threads_per_cta, int32_t* instances_per_cta)
{
    if (!h) return ACMPC_ERR_INVALID;
    if (n_launches) *n_launches = h->last_launches;
    if (smem_bytes) *smem_bytes = h->last_smem;
    if (threads_per_cta) *threads_per_cta = h->last_threads;
    if (instances_per_cta) *instances_per_cta = h->last_ipc;
    return ACMPC_OK;
}

int32_t acmpc_set_profiling(acmpc_handle* h, int32_t on)
{
    if (!h) return ACMPC_ERR_INVALID;
    if (fail(h, cudaSetDevice(h->device), "cudaSetDevice")) return ACMPC_ERR_CUDA;
    if (on && !h->ev) {
        h->ev = new (std::nothrow) cudaEvent_t[3 * kEventRing];
        if (!h->ev) return ACMPC_ERR_INVALID;
        for (int i = 0; i < 3 * kEventRing; ++i)
            if (fail(h, cudaEventCreate(&h->ev[i]), "cudaEventCreate")) return ACMPC_ERR_CUDA;
    }
    h->profiling = on ? 1 : 0;
    h->ev_head = 0, h->ev_count = 0;
    return ACMPC_OK;
}

int32_t acmpc_collect_kernel_ms(acmpc_handle* h, double* speed_ms, double* control_ms, int32_t* launches)
{
    if (!h) return ACMPC_ERR_INVALID;
    double s = 0.0, c = 0.0;
    const int cnt = h->ev_count;
    for (int k = 0; k < cnt; ++k) {
        const int slot = ((h->ev_head - 1 - k) % kEventRing + kEventRing) % kEventRing;
        cudaEvent_t* ev = h->ev + 3 * slot;
        if (fail(h, cudaEventSynchronize(ev[2]), "cudaEventSynchronize")) return ACMPC_ERR_CUDA;
        float a = 0.f, b = 0.f;
        if (fail(h, cudaEventElapsedTime(&a, ev[0], ev[1]), "cudaEventElapsedTime") ||
            fail(h, cudaEventElapsedTime(&b, ev[1], ev[2]), "cudaEventElapsedTime"))
            return ACMPC_ERR_CUDA;
        s += a, c += b;
    }
    if (speed_ms) *speed_ms = s;
    if (control_ms) *control_ms = c;
    if (launches) *launches = cnt;
    h->ev_head = 0, h->ev_count = 0;
    return ACMPC_OK;
}

int32_t acmpc_fp64_peak_tflops(int32_t device, double* tflops)
{
    if (!tflops) return ACMPC_ERR_INVALID;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        return ACMPC_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return ACMPC_ERR_CUDA;
    double* sink = nullptr;
    if (cudaMalloc(&sink, 8) != cudaSuccess) return ACMPC_ERR_CUDA;
    const int iters = 1 << 14, threads = 256, blocks = prop.multiProcessorCount * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, threads>>>(sink, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) {
            cudaFree(sink);
            return ACMPC_ERR_CUDA;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    cudaFree(sink);
    const double flops = 2.0 * 8.0 * (double)iters * threads * (double)blocks;
    *tflops = flops / (best * 1e-3) / 1e12;
    return ACMPC_OK;
}

}  // extern "C"

// ---- whole-track speed profile (map_profile.cuh) -------------------------------------------------------------
namespace {

// one launch of the map kernel.  h_track [M,3] or nullptr; h_way [7,n] in/out; do_solve 0 = waypoints only
int32_t run_map(acmpc_handle* h, int M, int n, const double* h_track, double* h_way, int do_solve, double v_max,
                double ay_max, double a_min, int max_iter, double* h_solution, acmpc_map_info* info)
{
    using namespace acmpc::mapqp;
    if (!h || !h_way || n < 2 || (h_track && M != n + 1)) return ACMPC_ERR_INVALID;
    const int ctas = (n + kThreads - 1) / kThreads;
    if (ctas > h->sm_count || ctas > kMaxCtas) {
        h->err = "track too long for one cooperative launch (n > 512 * SM count)";
        return ACMPC_ERR_INVALID;
    }
    if (fail(h, cudaSetDevice(h->device), "cudaSetDevice")) return ACMPC_ERR_CUDA;
    MapParams p;
    memset(&p, 0, sizeof(p));
    p.cfg = h->cfg;
    p.cfg.v_max = v_max, p.cfg.ay_max = ay_max, p.cfg.a_min = a_min;
    p.cfg.max_iter = max_iter > 0 ? max_iter : ACMPC_MAP_MAX_ITER;
    p.cfg.has_end_velocity = 0;   // compute_map_speed_profile passes no end velocity (spatial_mpc.py:71)
    p.M = M, p.n = n, p.do_solve = do_solve;
    const size_t way_b = (size_t)7 * n * 8, trk_b = h_track ? (size_t)3 * M * 8 : 0, k_b = (size_t)ctas * kThreads * 8;
    const size_t slots_b = (size_t)2 * ctas * ctas * sizeof(ulonglong4), halo_b = (size_t)2 * 2 * kMaxCtas * kWarps * 8,
                 red_b = (size_t)2 * kMaxCtas * kRed * 8;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t total = up(way_b) + up(trk_b) + up((size_t)n * 8) + 2 * up(k_b) + up(slots_b) + up(halo_b) +
                         up(red_b) + up(64) + 256;
    char* base = nullptr;
    if (fail(h, cudaMalloc((void**)&base, total), "cudaMalloc(map profile)")) return ACMPC_ERR_CUDA;
    char* q = base;
    auto take = [&](size_t b) { char* r = q; q += up(b); return r; };
    p.waypoints = (double*)take(way_b);
    double* d_track = (double*)take(trk_b);
    p.track = h_track ? d_track : nullptr;
    p.solution = (double*)take((size_t)n * 8);
    p.kd = (double*)take(k_b), p.ko = (double*)take(k_b);
    p.slots = (ulonglong4*)take(slots_b);
    p.halo = (double*)take(halo_b);
    p.red = (double*)take(red_b);
    p.info = (double*)take(64);
    p.counter = (unsigned*)take(256);
    cudaStream_t st = h->stream;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int32_t rc = ACMPC_OK;
    double h_info[8] = {0};
    float ms = 0.f;
    void* args[] = {(void*)&p};
    if (fail(h, cudaMemsetAsync(base, 0, total, st), "cudaMemsetAsync(map profile)") ||
        (h_track && fail(h, cudaMemcpyAsync(d_track, h_track, trk_b, cudaMemcpyHostToDevice, st), "H2D(track)")) ||
        (!h_track && fail(h, cudaMemcpyAsync(p.waypoints, h_way, way_b, cudaMemcpyHostToDevice, st), "H2D(waypoints)")) ||
        fail(h, cudaEventCreate(&e0), "cudaEventCreate") || fail(h, cudaEventCreate(&e1), "cudaEventCreate") ||
        fail(h, cudaEventRecord(e0, st), "cudaEventRecord") ||
        fail(h, cudaLaunchCooperativeKernel((const void*)acmpc_map_profile_kernel, dim3(ctas), dim3(kThreads), args, 0, st),
             "cudaLaunchCooperativeKernel(map profile)") ||
        fail(h, cudaEventRecord(e1, st), "cudaEventRecord") ||
        fail(h, cudaMemcpyAsync(h_way, p.waypoints, way_b, cudaMemcpyDeviceToHost, st), "D2H(waypoints)") ||
        (h_solution && do_solve &&
         fail(h, cudaMemcpyAsync(h_solution, p.solution, (size_t)n * 8, cudaMemcpyDeviceToHost, st), "D2H(solution)")) ||
        fail(h, cudaMemcpyAsync(h_info, p.info, 64, cudaMemcpyDeviceToHost, st), "D2H(info)") ||
        fail(h, cudaStreamSynchronize(st), "map profile kernel") ||
        fail(h, cudaEventElapsedTime(&ms, e0, e1), "cudaEventElapsedTime"))
        rc = ACMPC_ERR_CUDA;
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(base);
    if (rc == ACMPC_OK && h_info[7] != 0.0) {
        h->err = "map profile kernel: grid barrier timed out";
        rc = ACMPC_ERR_CUDA;
    }
    if (rc == ACMPC_OK) {
        h->last_launches = 1, h->last_smem = (int)sizeof(Shared), h->last_threads = kThreads, h->last_ipc = 1;
        if (info) {
            info->status = do_solve ? (int32_t)h_info[0] : ACMPC_UNSOLVED;
            info->iters = (int32_t)h_info[1], info->rho_updates = (int32_t)h_info[2], info->ctas = ctas;
            info->pri_res = h_info[3], info->dua_res = h_info[4], info->obj_val = h_info[5], info->rho = h_info[6];
            info->kernel_ms = ms;
        }
    }
    return rc;
}

// weights of the least-squares cubic over x = -10 .. 10 evaluated at x = e  (scipy.signal.savgol_coeffs /
// the polynomial fit of mode "interp")
void savgol_weights(double e, double* w)
{
    using namespace acmpc::mapqp;
    long double G[4][5];
    for (int a = 0; a < 4; ++a) {
        for (int b = 0; b < 4; ++b) {
            long double t = 0;
            for (int k = -kSgHalf; k <= kSgHalf; ++k) t += powl((long double)k, a + b);
            G[a][b] = t;
        }
        G[a][4] = powl((long double)e, a);
    }
    // solve G c = (e^a): weights w_k = sum_b c_b k^b
    for (int i = 0; i < 4; ++i) {
        int piv = i;
        for (int r = i + 1; r < 4; ++r)
            if (fabsl(G[r][i]) > fabsl(G[piv][i])) piv = r;
        for (int c = 0; c < 5; ++c) {
            long double t = G[i][c];
            G[i][c] = G[piv][c], G[piv][c] = t;
        }
        for (int r = 0; r < 4; ++r) {
            if (r == i) continue;
            const long double f = G[r][i] / G[i][i];
            for (int c = i; c < 5; ++c) G[r][c] -= f * G[i][c];
        }
    }
    for (int k = -kSgHalf; k <= kSgHalf; ++k) {
        long double t = 0;
        for (int b = 0; b < 4; ++b) t += (G[b][4] / G[b][b]) * powl((long double)k, b);
        w[k + kSgHalf] = (double)t;
    }
}

}  // namespace

extern "C" {

int32_t acmpc_construct_waypoints_host(acmpc_handle* h, int32_t M, const double* track, double* waypoints)
{
    if (!track) return ACMPC_ERR_INVALID;
    return run_map(h, M, M - 1, track, waypoints, 0, 0.0, 0.0, 0.0, 0, nullptr, nullptr);
}

int32_t acmpc_map_speed_profile_host(acmpc_handle* h, int32_t n, double* waypoints, double v_max, double ay_max,
                                     double a_min, int32_t max_iter, double* solution, acmpc_map_info* info)
{
    return run_map(h, n + 1, n, nullptr, waypoints, 1, v_max, ay_max, a_min, max_iter, solution, info);
}

int32_t acmpc_track_speed_profile_host(acmpc_handle* h, int32_t M, const double* track, double v_max, double ay_max,
                                       double a_min, int32_t max_iter, double* waypoints, double* solution,
                                       acmpc_map_info* info)
{
    if (!track) return ACMPC_ERR_INVALID;
    return run_map(h, M, M - 1, track, waypoints, 1, v_max, ay_max, a_min, max_iter, solution, info);
}

int32_t acmpc_reference_speeds_host(acmpc_handle* h, int32_t n, const double* velocities, int32_t behind,
                                    int32_t ahead, double* smoothed, double* window_mean)
{
    using namespace acmpc::mapqp;
    if (!h || !velocities || n < kSgWindow || behind < 0 || ahead < 0 || behind + ahead < 1) return ACMPC_ERR_INVALID;
    if (fail(h, cudaSetDevice(h->device), "cudaSetDevice")) return ACMPC_ERR_CUDA;
    SavgolCoeffs c;
    savgol_weights(0.0, c.interior);
    for (int i = 0; i < kSgHalf; ++i) savgol_weights((double)(i - kSgHalf), c.edge[i]);
    double* d = nullptr;
    const size_t nb = (size_t)n * 8;
    if (fail(h, cudaMalloc((void**)&d, 3 * nb), "cudaMalloc(reference speeds)")) return ACMPC_ERR_CUDA;
    double *d_v = d, *d_s = d + n, *d_w = d + 2 * (size_t)n;
    cudaStream_t st = h->stream;
    const int threads = 128, blocks = (n + threads - 1) / threads;
    int32_t rc = ACMPC_OK;
    if (fail(h, cudaMemcpyAsync(d_v, velocities, nb, cudaMemcpyHostToDevice, st), "H2D(velocities)")) rc = ACMPC_ERR_CUDA;
    if (rc == ACMPC_OK) {
        acmpc_savgol_kernel<<<blocks, threads, 0, st>>>(d_v, n, c, d_s);
        acmpc_window_mean_kernel<<<blocks, threads, 0, st>>>(d_s, n, behind, ahead, d_w);
        if (fail(h, cudaGetLastError(), "reference speed kernels") ||
            (smoothed && fail(h, cudaMemcpyAsync(smoothed, d_s, nb, cudaMemcpyDeviceToHost, st), "D2H(smoothed)")) ||
            (window_mean && fail(h, cudaMemcpyAsync(window_mean, d_w, nb, cudaMemcpyDeviceToHost, st), "D2H(window mean)")) ||
            fail(h, cudaStreamSynchronize(st), "reference speed kernels"))
            rc = ACMPC_ERR_CUDA;
    }
    cudaFree(d);
    if (rc == ACMPC_OK) h->last_launches = 2, h->last_smem = 0, h->last_threads = threads, h->last_ipc = 1;
    return rc;
}

}  // extern "C"

// ---- caller side of the step (publish.cuh) --------------------------------------------------------------------
namespace {

template <typename T>
int32_t select_commands(acmpc_handle* h, int32_t B, int32_t n, const T* cum_time, const T* commands, const double* elapsed,
                        int32_t mode, T* out, int32_t* indices)
{
    if (!h || B < 1 || n < 1 || !cum_time || !commands || !elapsed || !out || (mode != 0 && mode != 1)) return ACMPC_ERR_INVALID;
    const size_t nb = (size_t)B * n * sizeof(T);
    Staged s(h, 3 * nb + (size_t)B * (8 + 2 * sizeof(T) + 8) + 8 * 256);
    const T* d_ct = (const T*)s.in(cum_time, nb);
    const T* d_cm = (const T*)s.in(commands, 2 * nb);
    const double* d_el = (const double*)s.in(elapsed, (size_t)B * 8);
    T* d_out = (T*)s.out(out, (size_t)B * 2 * sizeof(T));
    int32_t* d_idx = (int32_t*)s.out(indices, (size_t)B * 8);
    if (s.rc == ACMPC_OK)
        acmpc::pub::select_commands_kernel<T><<<(B + 127) / 128, 128, 0, h->stream>>>(d_ct, d_cm, d_el, B, n, mode, d_out, d_idx);
    s.back(out, d_out, (size_t)B * 2 * sizeof(T));
    s.back(indices, d_idx, (size_t)B * 8);
    return s.finish(1, 128);
}

}  // namespace

extern "C" {

int32_t acmpc_reference_paths_host(acmpc_handle* h, int32_t B, int32_t P, const float* centrelines, double* paths)
{
    if (!h || B < 1 || !centrelines || !paths) return ACMPC_ERR_INVALID;
    const int H = h->cfg.horizon;
    if (P < H) return ACMPC_ERR_INVALID;
    const int ds = P / H;                          // int(len(centreline) / horizon), controller.py:260
    if ((P + ds - 1) / ds != H) {                  // centreline[0::ds] has another length: np.stack raises
        h->err = "reference path: len(centreline[0::ds]) != horizon (the reference's np.stack raises here)";
        return ACMPC_ERR_INVALID;
    }
    const size_t in_b = (size_t)B * P * 2 * sizeof(float), out_b = (size_t)B * H * 3 * 8;
    Staged s(h, in_b + out_b);
    const float* d_c = (const float*)s.in(centrelines, in_b);
    double* d_p = (double*)s.out(paths, out_b);
    if (s.rc == ACMPC_OK)
        acmpc::pub::reference_paths_kernel<<<(B * H + 127) / 128, 128, 0, h->stream>>>(d_c, B, P, H, ds, d_p);
    s.back(paths, d_p, out_b);
    return s.finish(1, 128);
}

int32_t acmpc_publish_host(acmpc_handle* h, int32_t B, const double* controls, const double* cum_time,
                           const double* prediction, float* control_inputs, float* control_cumtime,
                           float* predicted_locations)
{
    if (!h || B < 1) return ACMPC_ERR_INVALID;
    const int n = h->cfg.horizon - 1;
    const size_t e = (size_t)B * n;
    Staged s(h, e * (2 + 1 + 2) * (8 + 4) + 8 * 256);
    const double* d_c = (const double*)s.in(control_inputs ? controls : nullptr, e * 2 * 8);
    const double* d_t = (const double*)s.in(control_cumtime ? cum_time : nullptr, e * 8);
    const double* d_p = (const double*)s.in(predicted_locations ? prediction : nullptr, e * 2 * 8);
    float* o_c = (float*)s.out(d_c ? control_inputs : nullptr, e * 2 * 4);
    float* o_t = (float*)s.out(d_t ? control_cumtime : nullptr, e * 4);
    float* o_p = (float*)s.out(d_p ? predicted_locations : nullptr, e * 2 * 4);
    if (s.rc == ACMPC_OK)
        acmpc::pub::publish_kernel<<<((int)e + 127) / 128, 128, 0, h->stream>>>(d_c, d_t, d_p, B, n, o_c, o_t, o_p);
    if (o_c) s.back(control_inputs, o_c, e * 2 * 4);
    if (o_t) s.back(control_cumtime, o_t, e * 4);
    if (o_p) s.back(predicted_locations, o_p, e * 2 * 4);
    return s.finish(1, 128);
}

int32_t acmpc_select_commands_f32_host(acmpc_handle* h, int32_t B, int32_t n, const float* cum_time, const float* commands,
                                       const double* elapsed, int32_t mode, float* out, int32_t* indices)
{
    return select_commands<float>(h, B, n, cum_time, commands, elapsed, mode, out, indices);
}

int32_t acmpc_select_commands_f64_host(acmpc_handle* h, int32_t B, int32_t n, const double* cum_time, const double* commands,
                                       const double* elapsed, int32_t mode, double* out, int32_t* indices)
{
    return select_commands<double>(h, B, n, cum_time, commands, elapsed, mode, out, indices);
}

// ---- track side of the step: track_prep.cuh ---------------------------------------

int32_t acmpc_remove_near_duplicates_cols_host(acmpc_handle* h, int32_t M, int32_t cols, const double* rows, double tol,
                                               double* out, int32_t* kept)
{
    if (!h || M < 0 || cols < 2 || !kept || (M > 0 && (!rows || !out))) return ACMPC_ERR_INVALID;
    if (M == 0) {
        *kept = 0;
        return ACMPC_OK;
    }
    const int ctas = (M + acmpc::trk::kDupThreads - 1) / acmpc::trk::kDupThreads;
    const size_t b = (size_t)M * cols * 8;
    Staged s(h, 2 * b + (size_t)ctas * 4 + 4 * 256);
    const double* d_in = (const double*)s.in(rows, b);
    double* d_out = (double*)s.out(out, b);
    int* d_cnt = (int*)s.scratch((size_t)ctas * 4);
    int* d_kept = (int*)s.scratch(4);
    if (s.rc == ACMPC_OK) {
        acmpc::trk::near_duplicate_count_kernel<<<ctas, acmpc::trk::kDupThreads, 0, h->stream>>>(d_in, M, cols, tol, d_cnt);
        acmpc::trk::near_duplicate_scatter_kernel<<<ctas, acmpc::trk::kDupThreads, 0, h->stream>>>(d_in, M, cols, tol, d_cnt, d_out, d_kept);
    }
    s.back(kept, d_kept, 4);
    s.back(out, d_out, b);     // rows [kept, M) of `out` are unspecified
    return s.finish(2, acmpc::trk::kDupThreads);
}

int32_t acmpc_remove_near_duplicates_host(acmpc_handle* h, int32_t M, const double* xy, double tol, double* out, int32_t* kept)
{
    return acmpc_remove_near_duplicates_cols_host(h, M, 2, xy, tol, out, kept);
}

// ---- SpatialBicycleModel / update_prediction as stand-alone entry points (model.cuh) ------------------------------------
int32_t acmpc_t2s_host(acmpc_handle* h, int32_t B, const double* waypoints, const double* states, double* out)
{
    if (!h || B < 1 || !waypoints || !states || !out) return ACMPC_ERR_INVALID;
    const size_t b = (size_t)B * 3 * 8;
    Staged s(h, 3 * b + 4 * 256);
    const double* d_w = (const double*)s.in(waypoints, b);
    const double* d_s = (const double*)s.in(states, b);
    double* d_o = (double*)s.out(out, b);
    if (s.rc == ACMPC_OK) acmpc::model::t2s_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(d_w, d_s, B, d_o);
    s.back(out, d_o, b);
    return s.finish(1, 128);
}

int32_t acmpc_s2t_host(acmpc_handle* h, int32_t B, int32_t n, const double* waypoints, const double* states, double* out,
                       double* prediction)
{
    if (!h || B < 1 || n < 1 || !waypoints || !states || (!out && !prediction)) return ACMPC_ERR_INVALID;
    const size_t e = (size_t)B * n;
    Staged s(h, e * (7 + 3 + 3 + 2) * 8 + 6 * 256);
    const double* d_w = (const double*)s.in(waypoints, e * 7 * 8);
    const double* d_s = (const double*)s.in(states, e * 3 * 8);
    double* d_o = (double*)s.out(out, e * 3 * 8);
    double* d_p = (double*)s.out(prediction, e * 2 * 8);
    if (s.rc == ACMPC_OK) acmpc::model::s2t_kernel<<<(unsigned)((e + 127) / 128), 128, 0, h->stream>>>(d_w, d_s, B, n, d_o, d_p);
    s.back(out, d_o, e * 3 * 8);
    s.back(prediction, d_p, e * 2 * 8);
    return s.finish(1, 128);
}

int32_t acmpc_linearise_host(acmpc_handle* h, int32_t B, int32_t n, const double* waypoints, double* f, double* A, double* Bm)
{
    if (!h || B < 1 || n < 1 || !waypoints || (!f && !A && !Bm)) return ACMPC_ERR_INVALID;
    const size_t e = (size_t)B * n;
    Staged s(h, e * (7 + 3 + 9 + 6) * 8 + 6 * 256);
    const double* d_w = (const double*)s.in(waypoints, e * 7 * 8);
    double* d_f = (double*)s.out(f, e * 3 * 8);
    double* d_a = (double*)s.out(A, e * 9 * 8);
    double* d_b = (double*)s.out(Bm, e * 6 * 8);
    if (s.rc == ACMPC_OK) acmpc::model::linearise_kernel<<<(unsigned)((e + 127) / 128), 128, 0, h->stream>>>(d_w, B, n, d_f, d_a, d_b);
    s.back(f, d_f, e * 3 * 8);
    s.back(A, d_a, e * 9 * 8);
    s.back(Bm, d_b, e * 6 * 8);
    return s.finish(1, 128);
}

static int32_t polyfit_tracks(acmpc_handle* h, int32_t B, const int32_t* offsets, const double* points, const double* points_b,
                              int32_t num_points, int32_t degree, int32_t pad_origin, double* out, int32_t* status,
                              int32_t* start_index)
{
    if (!h || B < 1 || !offsets || num_points < 1 || degree < 0 || degree > acmpc::trk::kMaxDegree || !out) return ACMPC_ERR_INVALID;
    if (offsets[0] != 0) return ACMPC_ERR_INVALID;
    for (int b = 0; b < B; ++b)
        if (offsets[b + 1] < offsets[b]) return ACMPC_ERR_INVALID;
    const size_t total = (size_t)offsets[B];
    if (total > 0 && !points) return ACMPC_ERR_INVALID;
    const size_t in_b = total * 2 * 8, out_b = (size_t)B * num_points * 2 * 8;
    Staged s(h, (points_b ? 3 : 1) * in_b + out_b + (size_t)(3 * B + 1) * 4 + 8 * 256);
    const int* d_off = (const int*)s.in(offsets, (size_t)(B + 1) * 4);
    const double* d_pts = total ? (const double*)s.in(points, in_b) : nullptr;
    int launches = 1;
    if (points_b && total) {                       // centre track: the fit runs on (points + points_b) / 2
        const double* d_b = (const double*)s.in(points_b, in_b);
        double* d_mid = (double*)s.scratch(in_b);
        if (s.rc == ACMPC_OK)
            acmpc::trk::midline_kernel<<<(unsigned)((total * 2 + 255) / 256), 256, 0, h->stream>>>(d_pts, d_b, total * 2, d_mid);
        d_pts = d_mid, ++launches;
    }
    double* d_out = (double*)s.out(out, out_b);
    int* d_st = (int*)s.out(status, (size_t)B * 4);
    int* d_si = (int*)s.out(start_index, (size_t)B * 4);
    if (s.rc == ACMPC_OK)
        acmpc::trk::polyfit_resample_kernel<<<(B + 3) / 4, 128, 0, h->stream>>>(d_pts, d_off, B, num_points, degree, pad_origin,
                                                                               d_out, d_st, d_si);
    s.back(out, d_out, out_b);
    s.back(status, d_st, (size_t)B * 4);
    s.back(start_index, d_si, (size_t)B * 4);
    return s.finish(launches, 128);
}

int32_t acmpc_smooth_tracks_polyfit_host(acmpc_handle* h, int32_t B, const int32_t* offsets, const double* points,
                                         int32_t num_points, int32_t degree, double* out, int32_t* status, int32_t* start_index)
{
    return polyfit_tracks(h, B, offsets, points, nullptr, num_points, degree, 0, out, status, start_index);
}

int32_t acmpc_centre_tracks_host(acmpc_handle* h, int32_t B, int32_t N, const double* left, const double* right,
                                 int32_t num_points, double* centre, int32_t* status)
{
    if (!h || B < 1 || N < 1 || !left || !right) return ACMPC_ERR_INVALID;
    std::vector<int32_t> off((size_t)B + 1);
    for (int b = 0; b <= B; ++b) off[b] = b * N;
    // tracks.py:247-252: 10 origin points (x of the first centre point, y = 0) in front, degree 2
    return polyfit_tracks(h, B, off.data(), left, right, num_points, 2, 10, centre, status, nullptr);
}

int32_t acmpc_extract_paths_device(acmpc_handle* h, int32_t M, const double* d_centreline, int32_t B, const int32_t* d_index,
                                   const double* d_offset_lat, const double* d_offset_psi, double lookahead, double ds,
                                   double* d_paths, void* stream)
{
    if (!h || M < 2 || B < 1 || !d_centreline || !d_index || !d_paths || !(ds > 0.0) || !(lookahead >= 0.0)) return ACMPC_ERR_INVALID;
    if (reinterpret_cast<uintptr_t>(d_centreline) & 15) {
        h->err = "centre line must be 16-byte aligned (rows are read as double2)";
        return ACMPC_ERR_INVALID;
    }
    if (fail(h, cudaSetDevice(h->device), "cudaSetDevice")) return ACMPC_ERR_CUDA;
    const int H = h->cfg.horizon;
    cudaStream_t st = static_cast<cudaStream_t>(stream);   // as acmpc_solve_batch_device: NULL = the legacy default stream
    const double div = H > 1 ? (double)(H - 1) : 1.0;
    const size_t smem = (size_t)H * 24 * (1 + acmpc::trk::kExtractThreads / 32);   // step table + one [H,3] tile per warp
    if (smem > 48 * 1024 &&
        fail(h, cudaFuncSetAttribute(acmpc::trk::extract_paths_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
             "extract_paths_kernel: horizon too long for the shared-memory tiles"))
        return ACMPC_ERR_CUDA;
    acmpc::trk::extract_paths_kernel<<<(B + acmpc::trk::kExtractGroup - 1) / acmpc::trk::kExtractGroup, acmpc::trk::kExtractThreads,
                                       smem, st>>>(d_centreline, M, d_index, d_offset_lat, d_offset_psi, B, H, lookahead, ds,
                                                   lookahead / div, (6.0 - 10.0) / div, d_paths);
    if (fail(h, cudaGetLastError(), "extract_paths_kernel")) return ACMPC_ERR_CUDA;
    return ACMPC_OK;
}

int32_t acmpc_extract_paths_host(acmpc_handle* h, int32_t M, const double* centreline, int32_t B, const int32_t* index,
                                 const double* offset_lat, const double* offset_psi, double lookahead, double ds, double* paths)
{
    if (!h || M < 2 || B < 1 || !centreline || !index || !paths) return ACMPC_ERR_INVALID;
    const int H = h->cfg.horizon;
    const size_t out_b = (size_t)B * H * 3 * 8;
    Staged s(h, (size_t)M * 16 + (size_t)B * 20 + out_b + 8 * 256);
    const double* d_cl = (const double*)s.in(centreline, (size_t)M * 16);
    const int* d_i = (const int*)s.in(index, (size_t)B * 4);
    const double* d_l = (const double*)s.in(offset_lat, (size_t)B * 8);
    const double* d_p = (const double*)s.in(offset_psi, (size_t)B * 8);
    double* d_out = (double*)s.out(paths, out_b);
    if (s.rc == ACMPC_OK) {
        const int32_t rc = acmpc_extract_paths_device(h, M, d_cl, B, d_i, d_l, d_p, lookahead, ds, d_out, h->stream);
        if (rc != ACMPC_OK) return rc;
    }
    s.back(paths, d_out, out_b);
    return s.finish(1, 128);
}

}  // extern "C"
