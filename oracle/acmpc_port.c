/*
 * oracle/acmpc_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, FP64) of one ac-mpc MPC step, SpatialMPC.get_control
 * (/root/reference/src/acmpc/control/spatial_mpc.py:170-217), built on the OSQP
 * restatement in osqp_port.c.  It assembles the two QPs as general sparse CSC
 * matrices exactly as the reference does (no structure exploitation), so it is an
 * independent check of the structure-exploiting CUDA path.
 *
 * PARITY UNPINNED for the solver part (see osqp_port.h); the assembly part is pinned
 * against the reference's own Python run in this container (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may link or call this.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "../include/acmpc_b200.h"
#include "osqp_port.h"

#define NX 3
#define NU 2

typedef struct acmpc_port {
    acmpc_config cfg;
    int H, n;
    /* scratch for one instance */
    double *wp;                          /* (7,n) ReferencePath: xs ys psis kappas distances widths velocities */
    /* speed QP */
    int sn, sm;
    int *sPp, *sPi, *sAp, *sAi;
    double *sPx, *sAx, *sq, *sl, *su, *sx;
    opq_workspace *speed_ws[2];          /* [0] unlocalised, [1] localised (separate objects, spatial_mpc.py:55-56) */
    /* control QP */
    int cn, cm;
    int *cPp, *cPi, *cAp, *cAi;
    double *cPx, *cAx, *cq, *cl, *cu, *cx;
    opq_workspace *ctrl_ws;
} acmpc_port;

static double np_mod(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
    return r;
}

static void port_settings(const acmpc_config *c, opq_settings *s)
{
    opq_default_settings(s);
    s->rho = c->rho, s->sigma = c->sigma, s->alpha = c->alpha;
    s->eps_abs = c->eps_abs, s->eps_rel = c->eps_rel;
    s->eps_prim_inf = c->eps_prim_inf, s->eps_dual_inf = c->eps_dual_inf;
    s->adaptive_rho_tolerance = c->adaptive_rho_tolerance;
    s->scaling = c->scaling, s->max_iter = c->max_iter;
    s->check_termination = c->check_termination;
    s->adaptive_rho = c->adaptive_rho, s->adaptive_rho_interval = c->adaptive_rho_interval;
    s->warm_start = 1;
    s->check_dualgap = c->check_dualgap;
}

/* spatial_mpc.py:125-154 */
static void construct_waypoints(const acmpc_port *p, const double *W, double *wp)
{
    int n = p->n, H = p->H;
    double *xs = wp, *ys = wp + n, *psis = wp + 2 * n, *kap = wp + 3 * n, *dist = wp + 4 * n,
           *wid = wp + 5 * n, *vel = wp + 6 * n;
    for (int i = 0; i < n; i++) {
        const double *cur = W + 3 * i, *nxt = W + 3 * (i + 1);
        const double *prv = (i == 0) ? W + 3 * (H - 1) : W + 3 * (i - 1);
        double ax = nxt[0] - cur[0], ay = nxt[1] - cur[1];
        double bx = cur[0] - prv[0], by = cur[1] - prv[1];
        xs[i] = cur[0];
        ys[i] = cur[1];
        wid[i] = nxt[2];
        psis[i] = atan2(ay, ax);
        dist[i] = sqrt(ax * ax + ay * ay);
        double behind = atan2(by, bx);
        double dang = np_mod(psis[i] - behind + M_PI, 2.0 * M_PI) - M_PI;
        kap[i] = dang / (dist[i] + 1e-12) + 1e-12;
        vel[i] = 0.0;
    }
    kap[0] = kap[1];
}

/* speed_profile.py:26-59 (unlocalised) and :131-150 (localised) */
static void assemble_speed_qp(acmpc_port *p, const double *wp, double v_max_live, int localised)
{
    const acmpc_config *c = &p->cfg;
    int n = p->n;
    const double *kap = wp + 3 * n, *dist = wp + 4 * n;
    double *vmaxs = p->su + (n - 1);
    for (int i = 0; i < n; i++) {
        if (localised) {
            vmaxs[i] = v_max_live;
        } else {
            double ak = fabs(kap[i]);
            double vdyn = sqrt(c->ay_max / (ak + 1e-12));
            if (ak < c->ki_min) vdyn = v_max_live;
            double v = vdyn < v_max_live ? vdyn : v_max_live;
            v = c->v_min > v ? c->v_min : v;
            vmaxs[i] = v + 2.0;
        }
    }
    if (!localised && c->has_end_velocity) vmaxs[n - 1] = c->end_velocity;
    for (int i = 0; i < n - 1; i++) p->sl[i] = c->a_min, p->su[i] = c->a_max;
    for (int i = 0; i < n; i++) p->sl[n - 1 + i] = c->v_min, p->sq[i] = -1.0 * vmaxs[i];
    /* A = [D1 ; I], column j: row j-1 (+1/(2 d_{j-1})), row j (-1/(2 d_j)), row n-1+j (1) */
    int nz = 0;
    for (int j = 0; j < n; j++) {
        p->sAp[j] = nz;
        if (j >= 1) p->sAi[nz] = j - 1, p->sAx[nz] = 1.0 / (2.0 * dist[j - 1]), nz++;
        if (j <= n - 2) p->sAi[nz] = j, p->sAx[nz] = -1.0 / (2.0 * dist[j]), nz++;
        p->sAi[nz] = n - 1 + j, p->sAx[nz] = 1.0, nz++;
    }
    p->sAp[n] = nz;
}

/* dynamics.py:65-103 + solvers/control.py:26-79,121-158 */
static void assemble_control_qp(acmpc_port *p, const double *wp, const double x0[3])
{
    const acmpc_config *c = &p->cfg;
    int n = p->n, H = p->H;
    const double *kap = wp + 3 * n, *dist = wp + 4 * n, *wid = wp + 5 * n, *vel = wp + 6 * n;
    const double eps = 1e-12, inf = INFINITY;
    double margin = c->width / 2.0;
    int nz = 0, col = 0;
    /* columns of the states x_k */
    for (int k = 0; k < H; k++) {
        for (int s = 0; s < NX; s++, col++) {
            p->cAp[col] = nz;
            p->cAi[nz] = 3 * k + s, p->cAx[nz] = -1.0, nz++;
            if (k < n) {
                double d = dist[k], ka = kap[k], v = vel[k];
                int r0 = 3 * (k + 1);
                if (s == 0) {
                    p->cAi[nz] = r0 + 0, p->cAx[nz] = 1.0, nz++;
                    p->cAi[nz] = r0 + 1, p->cAx[nz] = -(ka * ka) * d, nz++;
                    p->cAi[nz] = r0 + 2, p->cAx[nz] = -ka / (v * d + eps), nz++;
                } else if (s == 1) {
                    p->cAi[nz] = r0 + 0, p->cAx[nz] = d, nz++;
                    p->cAi[nz] = r0 + 1, p->cAx[nz] = 1.0, nz++;
                } else {
                    p->cAi[nz] = r0 + 2, p->cAx[nz] = 1.0, nz++;
                }
            }
            p->cAi[nz] = 3 * H + col, p->cAx[nz] = 1.0, nz++;
        }
    }
    /* columns of the inputs u_k */
    for (int k = 0; k < n; k++) {
        double d = dist[k], v = vel[k];
        int r0 = 3 * (k + 1);
        p->cAp[col] = nz;
        p->cAi[nz] = r0 + 2, p->cAx[nz] = -1.0 / (v * v * d + eps), nz++;
        p->cAi[nz] = 3 * H + col, p->cAx[nz] = 1.0, nz++;
        col++;
        p->cAp[col] = nz;
        p->cAi[nz] = r0 + 1, p->cAx[nz] = d, nz++;
        p->cAi[nz] = 3 * H + col, p->cAx[nz] = 1.0, nz++;
        col++;
    }
    p->cAp[col] = nz;
    /* bounds: equality block then identity block */
    double *l = p->cl, *u = p->cu;
    for (int s = 0; s < NX; s++) l[s] = u[s] = -x0[s];
    for (int k = 0; k < n; k++) {
        double d = dist[k], ka = kap[k], v = vel[k];
        double b31 = -1.0 / (v * v * d + eps), f3 = 1.0 / (v * d + eps);
        double uq0 = 0.0 * v + 0.0 * ka - 0.0;
        double uq1 = 0.0 * v + d * ka - 0.0;
        double uq2 = b31 * v + 0.0 * ka - f3;
        int r0 = 3 * (k + 1);
        l[r0] = u[r0] = uq0;
        l[r0 + 1] = u[r0 + 1] = uq1;
        l[r0 + 2] = u[r0 + 2] = uq2;
    }
    int b0 = 3 * H;
    for (int k = 0; k < H; k++) {
        if (k == 0) {
            l[b0] = u[b0] = x0[0];
        } else {
            l[b0 + 3 * k] = (-wid[k - 1] / 2.0) + margin;
            u[b0 + 3 * k] = (wid[k - 1] / 2.0) - margin;
        }
        l[b0 + 3 * k + 1] = -inf, u[b0 + 3 * k + 1] = inf;
        l[b0 + 3 * k + 2] = 0.01, u[b0 + 3 * k + 2] = inf;
    }
    double kmax = tan(c->delta_max) / c->wheelbase;
    for (int k = 0; k < n; k++) {
        int r = b0 + 3 * H + 2 * k;
        l[r] = c->input_v_min - 0.1, u[r] = c->input_v_max + 0.1;
        l[r + 1] = -kmax, u[r + 1] = kmax;
    }
    /* cost: q = [-Q*xr (xr == 0), -QN*xr_N, -R*urs] */
    for (int j = 0; j < 3 * H; j++) p->cq[j] = 0.0;
    for (int k = 0; k < n; k++) {
        p->cq[3 * H + 2 * k] = -c->r_term[0] * vel[k];
        p->cq[3 * H + 2 * k + 1] = -c->r_term[1] * kap[k];
    }
}

acmpc_port *acmpc_port_create(const acmpc_config *cfg)
{
    if (cfg->horizon < ACMPC_MIN_HORIZON) return NULL;
    acmpc_port *p = (acmpc_port *)calloc(1, sizeof(*p));
    p->cfg = *cfg;
    int H = p->H = cfg->horizon, n = p->n = H - 1;
    p->wp = (double *)calloc((size_t)(7 * n), sizeof(double));
    /* speed QP storage */
    p->sn = n, p->sm = 2 * n - 1;
    p->sPp = (int *)calloc((size_t)n + 1, sizeof(int));
    p->sPi = (int *)calloc((size_t)n, sizeof(int));
    p->sPx = (double *)calloc((size_t)n, sizeof(double));
    for (int j = 0; j < n; j++) p->sPp[j] = j, p->sPi[j] = j, p->sPx[j] = 1.0;
    p->sPp[n] = n;
    p->sAp = (int *)calloc((size_t)n + 1, sizeof(int));
    p->sAi = (int *)calloc((size_t)(3 * n), sizeof(int));
    p->sAx = (double *)calloc((size_t)(3 * n), sizeof(double));
    p->sq = (double *)calloc((size_t)n, sizeof(double));
    p->sl = (double *)calloc((size_t)p->sm, sizeof(double));
    p->su = (double *)calloc((size_t)p->sm, sizeof(double));
    p->sx = (double *)calloc((size_t)n, sizeof(double));
    /* control QP storage */
    p->cn = 5 * H - 2, p->cm = 8 * H - 2;
    p->cPp = (int *)calloc((size_t)p->cn + 1, sizeof(int));
    p->cPi = (int *)calloc((size_t)p->cn, sizeof(int));
    p->cPx = (double *)calloc((size_t)p->cn, sizeof(double));
    for (int j = 0; j < p->cn; j++) {
        p->cPp[j] = j, p->cPi[j] = j;
        if (j < 3 * n) p->cPx[j] = cfg->step_cost[j % 3];
        else if (j < 3 * H) p->cPx[j] = cfg->final_cost[j - 3 * n];
        else p->cPx[j] = cfg->r_term[(j - 3 * H) % 2];
    }
    p->cPp[p->cn] = p->cn;
    p->cAp = (int *)calloc((size_t)p->cn + 1, sizeof(int));
    p->cAi = (int *)calloc((size_t)(16 * H), sizeof(int));
    p->cAx = (double *)calloc((size_t)(16 * H), sizeof(double));
    p->cq = (double *)calloc((size_t)p->cn, sizeof(double));
    p->cl = (double *)calloc((size_t)p->cm, sizeof(double));
    p->cu = (double *)calloc((size_t)p->cm, sizeof(double));
    p->cx = (double *)calloc((size_t)p->cn, sizeof(double));
    return p;
}

void acmpc_port_destroy(acmpc_port *p)
{
    if (!p) return;
    opq_free(p->speed_ws[0]), opq_free(p->speed_ws[1]), opq_free(p->ctrl_ws);
    free(p->wp);
    free(p->sPp), free(p->sPi), free(p->sPx), free(p->sAp), free(p->sAi), free(p->sAx);
    free(p->sq), free(p->sl), free(p->su), free(p->sx);
    free(p->cPp), free(p->cPi), free(p->cPx), free(p->cAp), free(p->cAi), free(p->cAx);
    free(p->cq), free(p->cl), free(p->cu), free(p->cx);
    free(p);
}

static int solve_ws(opq_workspace **ws, const opq_settings *st, int warm, int n, int m,
                    const int *Pp, const int *Pi, const double *Px, const double *q,
                    const int *Ap, const int *Ai, const double *Ax, const double *l,
                    const double *u, double *x, opq_info *info)
{
    if (*ws == NULL) {
        *ws = opq_setup(n, m, Pp, Pi, Px, q, Ap, Ai, Ax, l, u, st);
        if (*ws == NULL) {   /* data rejected (l > u): the reference's osqp.setup raises */
            memset(info, 0, sizeof(*info));
            info->status = OPQ_UNSOLVED;
            return OPQ_UNSOLVED;
        }
    } else {
        if (!warm) opq_cold_start(*ws);
        if (opq_update(*ws, q, l, u, Ax) != 0) {
            memset(info, 0, sizeof(*info));
            info->status = OPQ_UNSOLVED;
            return OPQ_UNSOLVED;
        }
    }
    return opq_solve(*ws, x, NULL, info);
}

/* One get_control.  warm != 0 keeps the OSQP objects' iterates and rho between calls like the
 * reference's persistent solver objects; warm == 0 is a cold start (fresh setup semantics).
 * `o` holds pointers to ONE instance's slices. */
int acmpc_port_step(acmpc_port *p, const double *path, double offset, double v_max_live,
                    int is_localised, int warm, const acmpc_outputs *o)
{
    const acmpc_config *c = &p->cfg;
    int n = p->n, H = p->H;
    double *wp = p->wp;
    opq_settings st;
    opq_info si, ci;
    port_settings(c, &st);
    construct_waypoints(p, path, wp);
    /* speed profile (spatial_mpc.py:89-123) */
    int loc = is_localised ? 1 : 0;
    assemble_speed_qp(p, wp, v_max_live, loc);
    int sstat = solve_ws(&p->speed_ws[loc], &st, warm, p->sn, p->sm, p->sPp, p->sPi, p->sPx, p->sq,
                         p->sAp, p->sAi, p->sAx, p->sl, p->su, p->sx, &si);
    double *vel = wp + 6 * n;
    if (sstat == OPQ_SOLVED) memcpy(vel, p->sx, sizeof(double) * (size_t)n);
    /* initial spatial state (spatial_mpc.py:186-189, dynamics.py:23-40) */
    double psi0 = wp[2 * n], x0[3];
    x0[0] = cos(psi0) * (0.0 - wp[n]) - sin(psi0) * (offset - wp[0]);
    x0[1] = np_mod((M_PI / 2.0 - psi0) + M_PI, 2.0 * M_PI) - M_PI;
    x0[2] = 0.0;
    assemble_control_qp(p, wp, x0);
    int cstat = solve_ws(&p->ctrl_ws, &st, warm, p->cn, p->cm, p->cPp, p->cPi, p->cPx, p->cq,
                         p->cAp, p->cAi, p->cAx, p->cl, p->cu, p->cx, &ci);
    /* unpack (spatial_mpc.py:193-212) + rollout (dynamics.py:42-63) */
    const double *x = p->cx;
    if (o->controls)
        for (int k = 0; k < n; k++) {
            o->controls[k] = x[3 * H + 2 * k];
            o->controls[n + k] = atan(x[3 * H + 2 * k + 1] * c->wheelbase);
        }
    if (o->prediction)
        for (int k = 0; k < n; k++) {
            double ey = x[3 * k], psi = wp[2 * n + k];
            o->prediction[2 * k] = wp[k] - ey * sin(psi);
            o->prediction[2 * k + 1] = wp[n + k] + ey * cos(psi);
        }
    if (o->cum_time)
        for (int k = 0; k < n; k++) o->cum_time[k] = x[3 * k + 2];
    if (o->derived)   /* spatial_mpc.py:208-211: times, accelerations (sic: e_y column), steer_rates */
        for (int k = 0; k + 1 < n; k++) {
            double dt = x[3 * (k + 1) + 2] - x[3 * k + 2];
            o->derived[k] = dt;
            o->derived[(n - 1) + k] = (x[3 * (k + 1)] - x[3 * k]) / dt;
            o->derived[2 * (n - 1) + k] = (x[3 * (k + 1) + 1] - x[3 * k + 1]) / dt;
        }
    if (o->states) memcpy(o->states, x, sizeof(double) * (size_t)(3 * H));
    if (o->v_ref) memcpy(o->v_ref, vel, sizeof(double) * (size_t)n);
    if (o->waypoints) memcpy(o->waypoints, wp, sizeof(double) * (size_t)(7 * n));
    if (o->cost) *o->cost = ci.obj_val;
    if (o->pri_res) *o->pri_res = ci.pri_res;
    if (o->dua_res) *o->dua_res = ci.dua_res;
    if (o->status) *o->status = cstat;
    if (o->status_speed) *o->status_speed = sstat;
    if (o->iters) o->iters[0] = si.iter, o->iters[1] = ci.iter;
    if (o->rho_updates) o->rho_updates[0] = si.rho_updates, o->rho_updates[1] = ci.rho_updates;
    return 0;
}

/* SpatialMPC.compute_speed_profile(reference_path, is_localised, end_vel) (spatial_mpc.py:89-123) on the object's
 * persistent speed solvers: `wp` (7,n) ReferencePath rows, velocities row written only when "solved".
 * has_end_vel / end_vel = the call's end_vel argument (None <=> has_end_vel == 0). */
int acmpc_port_speed_profile(acmpc_port *p, double *wp, double v_max_live, int is_localised, int has_end_vel,
                             double end_vel, int warm, double *x_out, int *iters, int *rho_updates)
{
    int n = p->n, loc = is_localised ? 1 : 0;
    opq_settings st;
    opq_info si;
    port_settings(&p->cfg, &st);
    int he = p->cfg.has_end_velocity;
    double ev = p->cfg.end_velocity;
    p->cfg.has_end_velocity = has_end_vel, p->cfg.end_velocity = end_vel;
    assemble_speed_qp(p, wp, v_max_live, loc);
    p->cfg.has_end_velocity = he, p->cfg.end_velocity = ev;
    int sstat = solve_ws(&p->speed_ws[loc], &st, warm, p->sn, p->sm, p->sPp, p->sPi, p->sPx, p->sq,
                         p->sAp, p->sAi, p->sAx, p->sl, p->su, p->sx, &si);
    if (sstat == OPQ_SOLVED) memcpy(wp + 6 * n, p->sx, sizeof(double) * (size_t)n);
    if (x_out) memcpy(x_out, p->sx, sizeof(double) * (size_t)n);
    if (iters) *iters = si.iter;
    if (rho_updates) *rho_updates = si.rho_updates;
    return sstat;
}

/* introspection for the assembly tests: the QP data of the last step */
int acmpc_port_get_qp(const acmpc_port *p, int which, int *n, int *m, const int **Ap,
                      const int **Ai, const double **Ax, const double **Pdiag, const double **q,
                      const double **l, const double **u)
{
    if (which == 0) {
        *n = p->sn, *m = p->sm, *Ap = p->sAp, *Ai = p->sAi, *Ax = p->sAx;
        *Pdiag = p->sPx, *q = p->sq, *l = p->sl, *u = p->su;
    } else {
        *n = p->cn, *m = p->cm, *Ap = p->cAp, *Ai = p->cAi, *Ax = p->cAx;
        *Pdiag = p->cPx, *q = p->cq, *l = p->cl, *u = p->cu;
    }
    return 0;
}

const double *acmpc_port_waypoints(const acmpc_port *p) { return p->wp; }

/* Batched cold-start solve of B independent instances on `nthreads` host threads (pthreads,
 * chunks of 16 instances handed out through an atomic counter): the CPU baseline of bench.py.
 * Mirrors acmpc_solve_batch_host. */
typedef struct {
    const acmpc_config *cfg;
    int B, is_localised;
    const double *paths, *offsets, *vmax;
    const acmpc_outputs *out;
    int *next;
    int *fail;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    const acmpc_outputs *out = j->out;
    int H = j->cfg->horizon, n = H - 1;
    acmpc_port *p = acmpc_port_create(j->cfg);
    if (!p) {
        __atomic_store_n(j->fail, 1, __ATOMIC_SEQ_CST);
        return NULL;
    }
    for (;;) {
        int b0 = __atomic_fetch_add(j->next, 16, __ATOMIC_SEQ_CST);
        if (b0 >= j->B) break;
        int b1 = b0 + 16 < j->B ? b0 + 16 : j->B;
        for (int b = b0; b < b1; b++) {
            acmpc_outputs o;
            memset(&o, 0, sizeof(o));
            if (out->controls) o.controls = out->controls + (size_t)b * 2 * n;
            if (out->prediction) o.prediction = out->prediction + (size_t)b * 2 * n;
            if (out->cum_time) o.cum_time = out->cum_time + (size_t)b * n;
            if (out->states) o.states = out->states + (size_t)b * 3 * H;
            if (out->v_ref) o.v_ref = out->v_ref + (size_t)b * n;
            if (out->cost) o.cost = out->cost + b;
            if (out->pri_res) o.pri_res = out->pri_res + b;
            if (out->dua_res) o.dua_res = out->dua_res + b;
            if (out->status) o.status = out->status + b;
            if (out->status_speed) o.status_speed = out->status_speed + b;
            if (out->iters) o.iters = out->iters + (size_t)b * 2;
            if (out->rho_updates) o.rho_updates = out->rho_updates + (size_t)b * 2;
            if (out->waypoints) o.waypoints = out->waypoints + (size_t)b * 7 * n;
            if (out->derived) o.derived = out->derived + (size_t)b * 3 * (n - 1);
            acmpc_port_step(p, j->paths + (size_t)b * 3 * H, j->offsets ? j->offsets[b] : 0.0,
                            j->vmax ? j->vmax[b] : j->cfg->v_max, j->is_localised, 0, &o);
        }
    }
    acmpc_port_destroy(p);
    return NULL;
}

int acmpc_port_solve_batch(const acmpc_config *cfg, int B, const double *paths,
                           const double *offsets, const double *vmax, int is_localised,
                           int nthreads, const acmpc_outputs *out)
{
    int next = 0, fail = 0;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    batch_job job = {cfg, B, is_localised, paths, offsets, vmax, out, &next, &fail};
    pthread_t th[256];
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, batch_worker, &job);
    batch_worker(&job);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    return fail;
}

/* Same defaults as the product's acmpc_default_config, restated so the oracle does not
 * link the product library. */
void acmpc_port_default_config(acmpc_config *c)
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
}
