// tests/_emul/emul.cpp -- TEST-ONLY 32-lane lock-step CPU emulation of the kernel body.
//
// Compiles ac_mpc_b200/csrc/mpc_warp.cuh with -DACMPC_EMULATE: the per-lane value types of simt.cuh
// become 32-wide arrays and every shuffle / reduction is applied lane by lane, so the SAME warp-parallel
// algorithm (lane ownership, shuffles, scans) runs in a GPU-less container and can be compared with the
// oracle.  It is NOT part of the product: the package never loads it and the C ABI does not expose it.
#define ACMPC_EMULATE 1
#include "../../ac_mpc_b200/csrc/mpc_warp.cuh"

#include <stdlib.h>
#include <string.h>

namespace {

template <int C>
void run(const acmpc_config* cfg, int B, const double* paths, const double* offsets, const double* vmax,
         int is_localised, const acmpc_outputs* out, double* warm, int use_warm)
{
    const int H = cfg->horizon, n = H - 1;
    using L = acmpc::Layout<C>;
    const size_t nd = (size_t)acmpc::smem_doubles<C>();
    const size_t nt = (size_t)L::kTmemDoubles * 32;   // tensor-memory model: [double column][lane]
    double* smem = (double*)malloc(sizeof(double) * nd);
    double* cold = (double*)malloc(sizeof(double) * (size_t)L::kColdDoubles);   // split layout: global memory
    double* tmem = (double*)malloc(sizeof(double) * nt);
    double* vel = (double*)malloc(sizeof(double) * (size_t)H);
    for (int b = 0; b < B; ++b) {
        memset(smem, 0xff, sizeof(double) * nd);   // NaN-poison
        memset(tmem, 0xff, sizeof(double) * nt);
        acmpc::Ctx<C> c;
        c.S = L::kColdGlobal ? cold : smem;
        c.W = L::kColdGlobal ? smem : smem + L::kColdDoubles;
        c.HS = c.W + L::kScratch;
        c.tm.p = tmem, c.H = H, c.n = n, c.cfg = cfg;
        c.lane = acmpc::lane_iota();
        double* raw = c.scratch();
        memcpy(raw, paths + (size_t)b * 3 * H, sizeof(double) * 3 * (size_t)H);
        acmpc::InstanceOut o;
        o.controls = out->controls ? out->controls + (size_t)b * 2 * n : nullptr;
        o.prediction = out->prediction ? out->prediction + (size_t)b * 2 * n : nullptr;
        o.cum_time = out->cum_time ? out->cum_time + (size_t)b * n : nullptr;
        o.states = out->states ? out->states + (size_t)b * 3 * H : nullptr;
        o.v_ref = out->v_ref ? out->v_ref + (size_t)b * n : nullptr;
        o.cost = out->cost ? out->cost + b : nullptr;
        o.pri_res = out->pri_res ? out->pri_res + b : nullptr;
        o.dua_res = out->dua_res ? out->dua_res + b : nullptr;
        o.status = out->status ? out->status + b : nullptr;
        o.status_speed = out->status_speed ? out->status_speed + b : nullptr;
        o.iters = out->iters ? out->iters + (size_t)b * 2 : nullptr;
        o.rho_updates = out->rho_updates ? out->rho_updates + (size_t)b * 2 : nullptr;
        o.waypoints = out->waypoints ? out->waypoints + (size_t)b * 7 * n : nullptr;
        o.derived = out->derived ? out->derived + (size_t)b * 3 * (n - 1) : nullptr;
        // the two phases are two kernels in the product; the hand-over is the speed profile
        double* wrec = warm ? warm + (size_t)b * acmpc::Layout<C>::kWarmDoubles : nullptr;
        acmpc::speed_instance<C>(c, raw, vmax ? vmax[b] : cfg->v_max, is_localised, vel, o, wrec, use_warm != 0);
        memset(smem, 0xff, sizeof(double) * nd);
        memset(cold, 0xff, sizeof(double) * (size_t)L::kColdDoubles);
        memcpy(raw, paths + (size_t)b * 3 * H, sizeof(double) * 3 * (size_t)H);
        acmpc::control_instance<C>(c, raw, vel, offsets ? offsets[b] : 0.0, o, wrec, use_warm != 0);
    }
    free(smem);
    free(cold);
    free(tmem);
    free(vel);
}

// stand-alone speed profile on ReferencePath rows (speed_instance's `way` mode)
template <int C>
void run_speed(const acmpc_config* cfg, int B, double* way, const double* vmax, int is_localised, double* sol, int* status,
               int* iters, int* rho_updates, double* warm, int use_warm)
{
    const int H = cfg->horizon, n = H - 1;
    const size_t nd = (size_t)acmpc::Layout<C>::kSpeedDoubles;
    double* smem = (double*)malloc(sizeof(double) * nd);
    for (int b = 0; b < B; ++b) {
        memset(smem, 0xff, sizeof(double) * nd);
        acmpc::Ctx<C> c;
        c.S = nullptr, c.HS = nullptr, c.W = smem, c.tm.p = nullptr, c.H = H, c.n = n, c.cfg = cfg;
        c.lane = acmpc::lane_iota();
        acmpc::InstanceOut o;
        memset(&o, 0, sizeof(o));
        int32_t it2[2] = {0, 0}, ru2[2] = {0, 0}, st = 0;
        o.status_speed = &st, o.iters = it2, o.rho_updates = ru2;
        double* wrec = warm ? warm + (size_t)b * acmpc::Layout<C>::kWarmDoubles : nullptr;
        acmpc::speed_instance<C>(c, nullptr, vmax ? vmax[b] : cfg->v_max, is_localised, sol ? sol + (size_t)b * n : nullptr, o,
                                 wrec, use_warm != 0, way + (size_t)b * 7 * n);
        status[b] = st, iters[b] = it2[0], rho_updates[b] = ru2[0];
    }
    free(smem);
}

}  // namespace

extern "C" int acmpc_emul_speed_profile(const acmpc_config* cfg, int B, double* way, const double* vmax, int is_localised,
                                        double* sol, int* status, int* iters, int* rho_updates, double* warm, int use_warm)
{
    switch ((cfg->horizon + 31) / 32) {
        case 1: run_speed<1>(cfg, B, way, vmax, is_localised, sol, status, iters, rho_updates, warm, use_warm); break;
        case 2: run_speed<2>(cfg, B, way, vmax, is_localised, sol, status, iters, rho_updates, warm, use_warm); break;
        case 3: run_speed<3>(cfg, B, way, vmax, is_localised, sol, status, iters, rho_updates, warm, use_warm); break;
        default: run_speed<4>(cfg, B, way, vmax, is_localised, sol, status, iters, rho_updates, warm, use_warm); break;
    }
    return 0;
}

extern "C" int acmpc_emul_warm_doubles(int H)
{
    switch ((H + 31) / 32) {
        case 1: return acmpc::Layout<1>::kWarmDoubles;
        case 2: return acmpc::Layout<2>::kWarmDoubles;
        case 3: return acmpc::Layout<3>::kWarmDoubles;
        default: return acmpc::Layout<4>::kWarmDoubles;
    }
}

// `warm`: NULL or [B, acmpc_emul_warm_doubles(H)] doubles (zero-initialised by the caller before first use)
extern "C" int acmpc_emul_solve_batch(const acmpc_config* cfg, int B, const double* paths,
                                      const double* offsets, const double* vmax, int is_localised,
                                      const acmpc_outputs* out, double* warm, int use_warm)
{
    const int H = cfg->horizon;
    if (H < ACMPC_MIN_HORIZON || H > ACMPC_MAX_HORIZON) return ACMPC_ERR_INVALID;
    switch ((H + 31) / 32) {
        case 1: run<1>(cfg, B, paths, offsets, vmax, is_localised, out, warm, use_warm); break;
        case 2: run<2>(cfg, B, paths, offsets, vmax, is_localised, out, warm, use_warm); break;
        case 3: run<3>(cfg, B, paths, offsets, vmax, is_localised, out, warm, use_warm); break;
        default: run<4>(cfg, B, paths, offsets, vmax, is_localised, out, warm, use_warm); break;
    }
    return 0;
}
