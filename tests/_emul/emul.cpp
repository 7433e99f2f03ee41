// tests/_emul/emul.cpp -- TEST-ONLY one-lane CPU emulation of the kernel body.
//
// Compiles ac_mpc_b200/csrc/mpc_body.cuh with -DACMPC_EMULATE (ACMPC_LANES == 1, no shuffles, no
// barriers) so the arithmetic of the CUDA path can be debugged in a GPU-less container against the
// oracle.  It is NOT part of the product: the package never loads it, the C ABI does not expose it,
// and it proves nothing about the parallel execution (that is what the -m gpu tests are for).
#define ACMPC_EMULATE 1
#include "../../ac_mpc_b200/csrc/mpc_body.cuh"

#include <stdlib.h>
#include <string.h>

extern "C" int acmpc_emul_smem_doubles(int H) { return acmpc::smem_doubles(H); }

extern "C" int acmpc_emul_solve_batch(const acmpc_config* cfg, int B, const double* paths,
                                      const double* offsets, const double* vmax, int is_localised,
                                      const acmpc_outputs* out)
{
    const int H = cfg->horizon, n = H - 1;
    if (H < ACMPC_MIN_HORIZON || H > ACMPC_MAX_HORIZON) return ACMPC_ERR_INVALID;
    double* smem = (double*)malloc(sizeof(double) * (size_t)acmpc::smem_doubles(H));
    for (int b = 0; b < B; ++b) {
        memset(smem, 0xff, sizeof(double) * (size_t)acmpc::smem_doubles(H));  // NaN-poison
        acmpc::Ctx c;
        c.S = smem, c.H = H, c.n = n, c.Hs = H, c.lane = 0, c.cfg = cfg;
        double* raw = c.f(acmpc::F_PATH_END);
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
        acmpc::solve_instance(c, raw, offsets ? offsets[b] : 0.0, vmax ? vmax[b] : cfg->v_max,
                              is_localised, o);
    }
    free(smem);
    return 0;
}
