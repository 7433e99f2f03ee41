/*
 * oracle/osqp_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see osqp_port.h).
 *
 * CPU restatement of OSQP 0.6.x for general sparse QPs
 *      min 1/2 x'Px + q'x   s.t.  l <= Ax <= u
 * following the published algorithm (Stellato et al. 2020, sections 3-5) and the
 * behaviour the reference relies on at
 *   /root/reference/src/acmpc/control/solvers/control.py:81-106   (setup / update / solve)
 *   /root/reference/src/acmpc/control/solvers/speed_profile.py:61-86,146
 *   /root/reference/src/acmpc/control/spatial_mpc.py:115,193      (status == "solved")
 *
 * Deliberate, documented deviations from the real library:
 *   - adaptive_rho_interval is a FIXED iteration count (the library's default is
 *     derived from wall-clock timing and is not reproducible);
 *   - the pristine (unscaled) problem data are kept next to the scaled copy, so an
 *     `update` re-equilibrates from exact data instead of un-scaling in place
 *     (differs from the library by rounding only);
 *   - fill-reducing ordering is a plain greedy minimum-degree, not AMD (changes the
 *     rounding of the KKT solve at the 1e-13 level, not the mathematics).
 *
 * PARITY UNPINNED (no OSQP golden vectors exist in the reference, no wheel here).
 */
#include "osqp_port.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

struct opq_workspace {
    int n, m, N;
    opq_settings s;
    double rho0;
    /* problem data: pristine (…0) and scaled copies */
    int *Pp, *Pi, Pnnz;
    double *Px0, *Px;
    int *Ap, *Ai, Annz;
    double *Ax0, *Ax;
    double *q0, *l0, *u0, *q, *l, *u;
    /* scaling */
    double *D, *E, *Dinv, *Einv, c, cinv;
    double *Dtmp, *DtmpA, *Etmp;
    /* rho */
    int *ctype;
    double *rho_vec, *rho_inv_vec;
    /* iterates */
    double *x, *z, *y, *x_prev, *z_prev, *xz;
    double *delta_x, *delta_y, *Axv, *Pxv, *Aty, *Atdy, *Adx, *Pdx;
    /* KKT, permuted upper triangle in CSC, plus scatter map from the sources */
    int *perm, *iperm;
    int *Kp, *Ki, Knnz;
    double *Kx;
    int ntrip, *trip_dst;
    int *etree, *Lnz, *Lp, *Li;
    double *Lx, *Dv, *Dvinv;
    int *iwork;
    double *fwork, *sol;
    /* info */
    opq_info info;
    int rho_updates;
    double xtPx, qtx, sc;   /* scaled x'Px, q'x, SC(y) of the last update_info (check_dualgap) */
};

void opq_default_settings(opq_settings *s)
{
    s->rho = 0.1;
    s->sigma = 1e-6;
    s->alpha = 1.6;
    s->eps_abs = 1e-3;
    s->eps_rel = 1e-3;
    s->eps_prim_inf = 1e-4;
    s->eps_dual_inf = 1e-4;
    s->adaptive_rho_tolerance = 5.0;
    s->scaling = 10;
    s->max_iter = 4000;
    s->check_termination = 25;
    s->adaptive_rho = 1;
    s->adaptive_rho_interval = 50;
    s->warm_start = 1;
    s->scaled_termination = 0;
    s->check_dualgap = 0;
}

const char *opq_status_string(int status)
{
    switch (status) {
    case OPQ_SOLVED: return "solved";
    case OPQ_SOLVED_INACCURATE: return "solved inaccurate";
    case OPQ_PRIMAL_INFEASIBLE: return "primal infeasible";
    case OPQ_PRIMAL_INFEASIBLE_INACCURATE: return "primal infeasible inaccurate";
    case OPQ_DUAL_INFEASIBLE: return "dual infeasible";
    case OPQ_DUAL_INFEASIBLE_INACCURATE: return "dual infeasible inaccurate";
    case OPQ_MAX_ITER_REACHED: return "maximum iterations reached";
    case OPQ_NON_CVX: return "problem non convex";
    default: return "unsolved";
    }
}

/* ------------------------------------------------------------------ helpers */

static double *dalloc(int n) { return (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }
static int *ialloc(int n) { return (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int)); }

static double norm_inf(const double *v, int n)
{
    double r = 0.0;
    for (int i = 0; i < n; i++) {
        double a = fabs(v[i]);
        if (a > r) r = a;
    }
    return r;
}

static double scaled_norm_inf(const double *s, const double *v, int n)
{
    double r = 0.0;
    for (int i = 0; i < n; i++) {
        double a = fabs(s[i] * v[i]);
        if (a > r) r = a;
    }
    return r;
}

/* y (+)= A x for CSC A (m rows, n cols) */
static void csc_mv(int n, const int *Ap, const int *Ai, const double *Ax, const double *x,
                   double *y, int m, int accumulate)
{
    if (!accumulate) memset(y, 0, sizeof(double) * (size_t)m);
    for (int j = 0; j < n; j++) {
        double xj = x[j];
        for (int p = Ap[j]; p < Ap[j + 1]; p++) y[Ai[p]] += Ax[p] * xj;
    }
}

/* y (+)= A' x ; skip_diag leaves out entries with row == col (upper-triangular P) */
static void csc_tmv(int n, const int *Ap, const int *Ai, const double *Ax, const double *x,
                    double *y, int accumulate, int skip_diag)
{
    for (int j = 0; j < n; j++) {
        double acc = accumulate ? y[j] : 0.0;
        for (int p = Ap[j]; p < Ap[j + 1]; p++) {
            if (skip_diag && Ai[p] == j) continue;
            acc += Ax[p] * x[Ai[p]];
        }
        y[j] = acc;
    }
}

/* full symmetric product with upper-triangular storage: y = P x */
static void sym_mv(const opq_workspace *w, const double *x, double *y)
{
    csc_mv(w->n, w->Pp, w->Pi, w->Px, x, y, w->n, 0);
    csc_tmv(w->n, w->Pp, w->Pi, w->Px, x, y, 1, 1);
}

static void limit_scaling(double *v, int n)
{
    for (int i = 0; i < n; i++) {
        if (v[i] < OPQ_MIN_SCALING) v[i] = 1.0;
        if (v[i] > OPQ_MAX_SCALING) v[i] = OPQ_MAX_SCALING;
    }
}

/* ------------------------------------------------- Ruiz equilibration (sec 5.1) */

static void scale_data(opq_workspace *w)
{
    int n = w->n, m = w->m;
    memcpy(w->Px, w->Px0, sizeof(double) * (size_t)w->Pnnz);
    memcpy(w->Ax, w->Ax0, sizeof(double) * (size_t)w->Annz);
    memcpy(w->q, w->q0, sizeof(double) * (size_t)n);
    memcpy(w->l, w->l0, sizeof(double) * (size_t)m);
    memcpy(w->u, w->u0, sizeof(double) * (size_t)m);
    w->c = 1.0;
    for (int i = 0; i < n; i++) w->D[i] = 1.0;
    for (int i = 0; i < m; i++) w->E[i] = 1.0;

    for (int pass = 0; pass < w->s.scaling; pass++) {
        /* inf-norms of the columns of the KKT matrix [P A'; A 0] */
        for (int j = 0; j < n; j++) w->Dtmp[j] = 0.0, w->DtmpA[j] = 0.0;
        for (int i = 0; i < m; i++) w->Etmp[i] = 0.0;
        for (int j = 0; j < n; j++) {
            for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) {
                double a = fabs(w->Px[p]);
                int i = w->Pi[p];
                if (a > w->Dtmp[j]) w->Dtmp[j] = a;
                if (i != j && a > w->Dtmp[i]) w->Dtmp[i] = a;
            }
            for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) {
                double a = fabs(w->Ax[p]);
                if (a > w->DtmpA[j]) w->DtmpA[j] = a;
                if (a > w->Etmp[w->Ai[p]]) w->Etmp[w->Ai[p]] = a;
            }
        }
        for (int j = 0; j < n; j++)
            if (w->DtmpA[j] > w->Dtmp[j]) w->Dtmp[j] = w->DtmpA[j];
        limit_scaling(w->Dtmp, n);
        limit_scaling(w->Etmp, m);
        for (int j = 0; j < n; j++) w->Dtmp[j] = 1.0 / sqrt(w->Dtmp[j]);
        for (int i = 0; i < m; i++) w->Etmp[i] = 1.0 / sqrt(w->Etmp[i]);
        /* P <- D P D ; A <- E A D ; q <- D q */
        for (int j = 0; j < n; j++) {
            for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) {
                w->Px[p] *= w->Dtmp[w->Pi[p]];
                w->Px[p] *= w->Dtmp[j];
            }
            for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) {
                w->Ax[p] *= w->Etmp[w->Ai[p]];
                w->Ax[p] *= w->Dtmp[j];
            }
            w->q[j] *= w->Dtmp[j];
            w->D[j] *= w->Dtmp[j];
        }
        for (int i = 0; i < m; i++) w->E[i] *= w->Etmp[i];
        /* cost normalisation */
        for (int j = 0; j < n; j++) w->Dtmp[j] = 0.0;
        for (int j = 0; j < n; j++)
            for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) {
                double a = fabs(w->Px[p]);
                int i = w->Pi[p];
                if (a > w->Dtmp[j]) w->Dtmp[j] = a;
                if (i != j && a > w->Dtmp[i]) w->Dtmp[i] = a;
            }
        double c_tmp = 0.0;
        for (int j = 0; j < n; j++) c_tmp += w->Dtmp[j];
        c_tmp /= (double)n;
        double nq = norm_inf(w->q, n);
        limit_scaling(&nq, 1);
        if (nq > c_tmp) c_tmp = nq;
        limit_scaling(&c_tmp, 1);
        c_tmp = 1.0 / c_tmp;
        for (int p = 0; p < w->Pnnz; p++) w->Px[p] *= c_tmp;
        for (int j = 0; j < n; j++) w->q[j] *= c_tmp;
        w->c *= c_tmp;
    }
    w->cinv = 1.0 / w->c;
    for (int j = 0; j < n; j++) w->Dinv[j] = 1.0 / w->D[j];
    for (int i = 0; i < m; i++) {
        w->Einv[i] = 1.0 / w->E[i];
        w->l[i] *= w->E[i];
        w->u[i] *= w->E[i];
    }
}

/* ------------------------------------------------------ rho vector (sec 5.2) */

static void set_rho_vec(opq_workspace *w)
{
    if (w->s.rho < OPQ_RHO_MIN) w->s.rho = OPQ_RHO_MIN;
    if (w->s.rho > OPQ_RHO_MAX) w->s.rho = OPQ_RHO_MAX;
    for (int i = 0; i < w->m; i++) {
        if (w->l[i] < -OPQ_INFTY * OPQ_MIN_SCALING && w->u[i] > OPQ_INFTY * OPQ_MIN_SCALING) {
            w->ctype[i] = -1;
            w->rho_vec[i] = OPQ_RHO_MIN;
        } else if (w->u[i] - w->l[i] < OPQ_RHO_TOL) {
            w->ctype[i] = 1;
            w->rho_vec[i] = OPQ_RHO_EQ_OVER_RHO_INEQ * w->s.rho;
        } else {
            w->ctype[i] = 0;
            w->rho_vec[i] = w->s.rho;
        }
        w->rho_inv_vec[i] = 1.0 / w->rho_vec[i];
    }
}

/* --------------------------------------------- ordering + sparse LDL' of the KKT */

/* greedy minimum degree on the symmetric pattern given as an edge list */
static void min_degree_order(int N, int nedge, const int *ea, const int *eb, int *perm)
{
    int *deg = ialloc(N), *cap = ialloc(N), *alive = ialloc(N), *mark = ialloc(N);
    int **adj = (int **)calloc((size_t)N, sizeof(int *));
    for (int e = 0; e < nedge; e++)
        if (ea[e] != eb[e]) deg[ea[e]]++, deg[eb[e]]++;
    for (int i = 0; i < N; i++) {
        cap[i] = deg[i] + 4;
        adj[i] = ialloc(cap[i]);
        deg[i] = 0;
        alive[i] = 1;
        mark[i] = -1;
    }
    for (int e = 0; e < nedge; e++) {
        int a = ea[e], b = eb[e];
        if (a == b) continue;
        adj[a][deg[a]++] = b;
        adj[b][deg[b]++] = a;
    }
    /* remove duplicate edges */
    for (int i = 0; i < N; i++) {
        int k = 0;
        for (int t = 0; t < deg[i]; t++) {
            int j = adj[i][t];
            if (mark[j] != i) mark[j] = i, adj[i][k++] = j;
        }
        deg[i] = k;
    }
    for (int i = 0; i < N; i++) mark[i] = -1;
    int stamp = 0;
    for (int step = 0; step < N; step++) {
        int p = -1;
        for (int i = 0; i < N; i++)
            if (alive[i] && (p < 0 || deg[i] < deg[p])) p = i;
        perm[step] = p;
        alive[p] = 0;
        for (int t = 0; t < deg[p]; t++) {
            int i = adj[p][t];
            /* adj[i] <- (adj[i] U adj[p]) \ {i, p} */
            stamp++;
            int k = 0;
            for (int s = 0; s < deg[i]; s++) {
                int j = adj[i][s];
                if (j == p) continue;
                mark[j] = stamp;
                adj[i][k++] = j;
            }
            mark[i] = stamp;
            int need = k + deg[p];
            if (need > cap[i]) {
                cap[i] = 2 * need;
                adj[i] = (int *)realloc(adj[i], sizeof(int) * (size_t)cap[i]);
            }
            for (int s = 0; s < deg[p]; s++) {
                int j = adj[p][s];
                if (mark[j] != stamp) mark[j] = stamp, adj[i][k++] = j;
            }
            deg[i] = k;
        }
    }
    for (int i = 0; i < N; i++) free(adj[i]);
    free(adj), free(deg), free(cap), free(alive), free(mark);
}

/* Source triplets of the KKT matrix, in this order:
 *   P entries (Pnnz) | sigma on the n diagonals | A entries (Annz) | -1/rho (m)     */
static void kkt_symbolic(opq_workspace *w)
{
    int n = w->n, m = w->m, N = n + m;
    int nt = w->Pnnz + n + w->Annz + m;
    int *ti = ialloc(nt), *tj = ialloc(nt);
    int t = 0;
    for (int j = 0; j < n; j++)
        for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) ti[t] = w->Pi[p], tj[t] = j, t++;
    for (int j = 0; j < n; j++) ti[t] = j, tj[t] = j, t++;
    for (int j = 0; j < n; j++)
        for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) ti[t] = j, tj[t] = n + w->Ai[p], t++;
    for (int i = 0; i < m; i++) ti[t] = n + i, tj[t] = n + i, t++;
    w->ntrip = nt;
    w->perm = ialloc(N);
    w->iperm = ialloc(N);
    min_degree_order(N, nt, ti, tj, w->perm);
    for (int k = 0; k < N; k++) w->iperm[w->perm[k]] = k;
    /* permute, force upper triangle, bucket by column, merge duplicates */
    int *cnt = ialloc(N + 1);
    for (t = 0; t < nt; t++) {
        int a = w->iperm[ti[t]], b = w->iperm[tj[t]];
        if (a > b) { int s = a; a = b; b = s; }
        ti[t] = a, tj[t] = b;
        cnt[b + 1]++;
    }
    for (int j = 0; j < N; j++) cnt[j + 1] += cnt[j];
    int *ord = ialloc(nt), *pos = ialloc(N + 1);
    memcpy(pos, cnt, sizeof(int) * (size_t)(N + 1));
    for (t = 0; t < nt; t++) ord[pos[tj[t]]++] = t;
    /* sort rows inside every column (insertion sort; columns are short) */
    for (int j = 0; j < N; j++)
        for (int a = cnt[j] + 1; a < cnt[j + 1]; a++) {
            int v = ord[a], b = a - 1;
            while (b >= cnt[j] && ti[ord[b]] > ti[v]) ord[b + 1] = ord[b], b--;
            ord[b + 1] = v;
        }
    w->Kp = ialloc(N + 1);
    w->Ki = ialloc(nt);
    w->trip_dst = ialloc(nt);
    int nz = 0;
    for (int j = 0; j < N; j++) {
        w->Kp[j] = nz;
        for (int a = cnt[j]; a < cnt[j + 1]; a++) {
            int tt = ord[a];
            if (a > cnt[j] && ti[ord[a - 1]] == ti[tt]) {
                w->trip_dst[tt] = nz - 1;
            } else {
                w->Ki[nz] = ti[tt];
                w->trip_dst[tt] = nz;
                nz++;
            }
        }
    }
    w->Kp[N] = nz;
    w->Knnz = nz;
    w->Kx = dalloc(nz);
    free(ti), free(tj), free(cnt), free(ord), free(pos);

    /* elimination tree and column counts of L (row-subtree walk) */
    w->etree = ialloc(N);
    w->Lnz = ialloc(N);
    w->Lp = ialloc(N + 1);
    int *flag = ialloc(N);
    for (int k = 0; k < N; k++) {
        w->etree[k] = -1;
        flag[k] = k;
        for (int p = w->Kp[k]; p < w->Kp[k + 1]; p++) {
            int i = w->Ki[p];
            while (i < k && flag[i] != k) {
                if (w->etree[i] < 0) w->etree[i] = k;
                w->Lnz[i]++;
                flag[i] = k;
                i = w->etree[i];
            }
        }
    }
    w->Lp[0] = 0;
    for (int k = 0; k < N; k++) w->Lp[k + 1] = w->Lp[k] + w->Lnz[k];
    w->Li = ialloc(w->Lp[N]);
    w->Lx = dalloc(w->Lp[N]);
    w->Dv = dalloc(N);
    w->Dvinv = dalloc(N);
    w->iwork = ialloc(4 * N);
    w->fwork = dalloc(N);
    w->sol = dalloc(N);
    free(flag);
}

static void kkt_numeric_values(opq_workspace *w)
{
    int n = w->n, m = w->m;
    memset(w->Kx, 0, sizeof(double) * (size_t)w->Knnz);
    int t = 0;
    for (int p = 0; p < w->Pnnz; p++) w->Kx[w->trip_dst[t++]] += w->Px[p];
    for (int j = 0; j < n; j++) w->Kx[w->trip_dst[t++]] += w->s.sigma;
    for (int p = 0; p < w->Annz; p++) w->Kx[w->trip_dst[t++]] += w->Ax[p];
    for (int i = 0; i < m; i++) w->Kx[w->trip_dst[t++]] += -w->rho_inv_vec[i];
}

/* up-looking LDL' : row k of L is obtained from a sparse triangular solve whose
 * pattern is the reach of column k of K in the elimination tree */
static int kkt_factor(opq_workspace *w)
{
    int N = w->N;
    int *fill = w->iwork, *stack = w->iwork + N, *pat = w->iwork + 2 * N, *flag = w->iwork + 3 * N;
    double *yv = w->fwork;
    kkt_numeric_values(w);
    for (int k = 0; k < N; k++) fill[k] = 0, yv[k] = 0.0, flag[k] = -1;
    for (int k = 0; k < N; k++) {
        int top = N;
        flag[k] = k;
        double dk = 0.0;
        for (int p = w->Kp[k]; p < w->Kp[k + 1]; p++) {
            int i = w->Ki[p];
            if (i == k) { dk = w->Kx[p]; continue; }
            yv[i] = w->Kx[p];
            int len = 0;
            while (flag[i] != k) {
                stack[len++] = i;
                flag[i] = k;
                i = w->etree[i];
            }
            while (len > 0) pat[--top] = stack[--len];
        }
        for (; top < N; top++) {
            int i = pat[top];
            double yi = yv[i];
            yv[i] = 0.0;
            int pend = w->Lp[i] + fill[i];
            for (int p = w->Lp[i]; p < pend; p++) yv[w->Li[p]] -= w->Lx[p] * yi;
            double lki = yi * w->Dvinv[i];
            dk -= yi * lki;
            w->Li[pend] = k;
            w->Lx[pend] = lki;
            fill[i]++;
        }
        if (dk == 0.0) return -1;
        w->Dv[k] = dk;
        w->Dvinv[k] = 1.0 / dk;
    }
    return 0;
}

/* solve K s = b in place on b = [b_x ; b_z] (unpermuted order) */
static void kkt_solve(opq_workspace *w, double *b)
{
    int N = w->N;
    double *s = w->sol;
    for (int k = 0; k < N; k++) s[k] = b[w->perm[k]];
    for (int k = 0; k < N; k++) {
        double v = s[k];
        for (int p = w->Lp[k]; p < w->Lp[k + 1]; p++) s[w->Li[p]] -= w->Lx[p] * v;
    }
    for (int k = 0; k < N; k++) s[k] *= w->Dvinv[k];
    for (int k = N - 1; k >= 0; k--) {
        double v = s[k];
        for (int p = w->Lp[k]; p < w->Lp[k + 1]; p++) v -= w->Lx[p] * s[w->Li[p]];
        s[k] = v;
    }
    for (int k = 0; k < N; k++) b[w->perm[k]] = s[k];
}

/* ------------------------------------------------------------------ setup */

opq_workspace *opq_setup(int n, int m, const int *Pp, const int *Pi, const double *Px,
                         const double *q, const int *Ap, const int *Ai, const double *Ax,
                         const double *l, const double *u, const opq_settings *settings)
{
    /* validate_data: "Lower bound must be lower than or equal to upper bound" -- osqp_setup fails (the Python
     * wrapper raises ValueError); osqp_update_bounds applies the same test (opq_update below) */
    for (int i = 0; i < m; i++)
        if (l[i] > u[i]) return NULL;
    opq_workspace *w = (opq_workspace *)calloc(1, sizeof(*w));
    w->n = n, w->m = m, w->N = n + m;
    w->s = *settings;
    w->rho0 = settings->rho;
    w->Pnnz = Pp[n];
    w->Annz = Ap[n];
    w->Pp = ialloc(n + 1), w->Pi = ialloc(w->Pnnz);
    w->Ap = ialloc(n + 1), w->Ai = ialloc(w->Annz);
    memcpy(w->Pp, Pp, sizeof(int) * (size_t)(n + 1));
    memcpy(w->Pi, Pi, sizeof(int) * (size_t)w->Pnnz);
    memcpy(w->Ap, Ap, sizeof(int) * (size_t)(n + 1));
    memcpy(w->Ai, Ai, sizeof(int) * (size_t)w->Annz);
    w->Px0 = dalloc(w->Pnnz), w->Px = dalloc(w->Pnnz);
    w->Ax0 = dalloc(w->Annz), w->Ax = dalloc(w->Annz);
    memcpy(w->Px0, Px, sizeof(double) * (size_t)w->Pnnz);
    memcpy(w->Ax0, Ax, sizeof(double) * (size_t)w->Annz);
    w->q0 = dalloc(n), w->q = dalloc(n);
    w->l0 = dalloc(m), w->l = dalloc(m), w->u0 = dalloc(m), w->u = dalloc(m);
    memcpy(w->q0, q, sizeof(double) * (size_t)n);
    for (int i = 0; i < m; i++) {
        w->l0[i] = l[i] < -OPQ_INFTY ? -OPQ_INFTY : l[i];
        w->u0[i] = u[i] > OPQ_INFTY ? OPQ_INFTY : u[i];
    }
    w->D = dalloc(n), w->Dinv = dalloc(n), w->Dtmp = dalloc(n), w->DtmpA = dalloc(n);
    w->E = dalloc(m), w->Einv = dalloc(m), w->Etmp = dalloc(m);
    w->ctype = ialloc(m), w->rho_vec = dalloc(m), w->rho_inv_vec = dalloc(m);
    w->x = dalloc(n), w->x_prev = dalloc(n), w->delta_x = dalloc(n);
    w->z = dalloc(m), w->z_prev = dalloc(m), w->y = dalloc(m), w->delta_y = dalloc(m);
    w->xz = dalloc(n + m);
    w->Axv = dalloc(m), w->Pxv = dalloc(n), w->Aty = dalloc(n);
    w->Atdy = dalloc(n), w->Adx = dalloc(m), w->Pdx = dalloc(n);
    if (w->s.scaling > 0) {
        scale_data(w);
    } else {
        int keep = w->s.scaling;
        w->s.scaling = 0;
        scale_data(w);
        w->s.scaling = keep;
    }
    set_rho_vec(w);
    kkt_symbolic(w);
    if (kkt_factor(w) != 0) {
        opq_free(w);
        return NULL;
    }
    w->info.status = OPQ_UNSOLVED;
    return w;
}

void opq_free(opq_workspace *w)
{
    if (!w) return;
    free(w->Pp), free(w->Pi), free(w->Px0), free(w->Px);
    free(w->Ap), free(w->Ai), free(w->Ax0), free(w->Ax);
    free(w->q0), free(w->l0), free(w->u0), free(w->q), free(w->l), free(w->u);
    free(w->D), free(w->E), free(w->Dinv), free(w->Einv);
    free(w->Dtmp), free(w->DtmpA), free(w->Etmp);
    free(w->ctype), free(w->rho_vec), free(w->rho_inv_vec);
    free(w->x), free(w->z), free(w->y), free(w->x_prev), free(w->z_prev), free(w->xz);
    free(w->delta_x), free(w->delta_y), free(w->Axv), free(w->Pxv), free(w->Aty);
    free(w->Atdy), free(w->Adx), free(w->Pdx);
    free(w->perm), free(w->iperm), free(w->Kp), free(w->Ki), free(w->Kx), free(w->trip_dst);
    free(w->etree), free(w->Lnz), free(w->Lp), free(w->Li), free(w->Lx), free(w->Dv), free(w->Dvinv);
    free(w->iwork), free(w->fwork), free(w->sol);
    free(w);
}

int opq_update(opq_workspace *w, const double *q, const double *l, const double *u,
               const double *Ax)
{
    if (q) memcpy(w->q0, q, sizeof(double) * (size_t)w->n);
    for (int i = 0; i < w->m; i++) {
        if (l) w->l0[i] = l[i] < -OPQ_INFTY ? -OPQ_INFTY : l[i];
        if (u) w->u0[i] = u[i] > OPQ_INFTY ? OPQ_INFTY : u[i];
    }
    for (int i = 0; i < w->m; i++)
        if (w->l0[i] > w->u0[i]) return 1;
    if (Ax) memcpy(w->Ax0, Ax, sizeof(double) * (size_t)w->Annz);
    scale_data(w);
    set_rho_vec(w);
    w->info.status = OPQ_UNSOLVED;
    return kkt_factor(w);
}

void opq_cold_start(opq_workspace *w)
{
    memset(w->x, 0, sizeof(double) * (size_t)w->n);
    memset(w->z, 0, sizeof(double) * (size_t)w->m);
    memset(w->y, 0, sizeof(double) * (size_t)w->m);
    if (w->s.rho != w->rho0) {
        w->s.rho = w->rho0;
        set_rho_vec(w);
        kkt_factor(w);
    }
}

void opq_warm_start(opq_workspace *w, const double *x, const double *y)
{
    for (int j = 0; j < w->n; j++) w->x[j] = x[j] * w->Dinv[j];
    for (int i = 0; i < w->m; i++) w->y[i] = y[i] * w->Einv[i] * w->c;
    csc_mv(w->n, w->Ap, w->Ai, w->Ax, w->x, w->z, w->m, 0);
}

/* ------------------------------------------------------------ ADMM pieces */

static double compute_pri_res(opq_workspace *w)
{
    /* z_prev doubles as the residual work vector, as in the library */
    csc_mv(w->n, w->Ap, w->Ai, w->Ax, w->x, w->Axv, w->m, 0);
    for (int i = 0; i < w->m; i++) w->z_prev[i] = w->Axv[i] - w->z[i];
    if (w->s.scaling && !w->s.scaled_termination) return scaled_norm_inf(w->Einv, w->z_prev, w->m);
    return norm_inf(w->z_prev, w->m);
}

static double compute_dua_res(opq_workspace *w)
{
    sym_mv(w, w->x, w->Pxv);
    for (int j = 0; j < w->n; j++) w->x_prev[j] = w->q[j] + w->Pxv[j];
    if (w->m > 0) {
        csc_tmv(w->n, w->Ap, w->Ai, w->Ax, w->y, w->Aty, 0, 0);
        for (int j = 0; j < w->n; j++) w->x_prev[j] += w->Aty[j];
    }
    if (w->s.scaling && !w->s.scaled_termination)
        return w->cinv * scaled_norm_inf(w->Dinv, w->x_prev, w->n);
    return norm_inf(w->x_prev, w->n);
}

static double compute_obj_val(opq_workspace *w)
{
    /* 1/2 x'Px + q'x with upper-triangular P */
    double quad = 0.0, lin = 0.0;
    for (int j = 0; j < w->n; j++) {
        for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) {
            int i = w->Pi[p];
            if (i == j) quad += 0.5 * w->Px[p] * w->x[i] * w->x[j];
            else quad += w->Px[p] * w->x[i] * w->x[j];
        }
        lin += w->q[j] * w->x[j];
    }
    double v = quad + lin;
    if (w->s.scaling) v *= w->cinv;
    return v;
}

static void update_info(opq_workspace *w, int iter)
{
    w->info.iter = iter;
    w->info.obj_val = compute_obj_val(w);
    w->info.pri_res = w->m ? compute_pri_res(w) : 0.0;
    w->info.dua_res = compute_dua_res(w);
    if (w->s.check_dualgap) {
        /* OSQP 1.x compute_obj_val_dual_gap: SC(y) = u'max(y,0) + l'min(y,0); sides beyond OSQP_INFTY * MIN_SCALING
         * are infinite bounds and contribute nothing */
        const double big = OPQ_INFTY * OPQ_MIN_SCALING;
        double a = 0.0, b = 0.0, sc = 0.0;
        for (int j = 0; j < w->n; j++) a += w->Pxv[j] * w->x[j], b += w->q[j] * w->x[j];
        for (int i = 0; i < w->m; i++) {
            double y = w->y[i];
            if (w->u[i] < big && y > 0.0) sc += w->u[i] * y;
            if (w->l[i] > -big && y < 0.0) sc += w->l[i] * y;
        }
        w->xtPx = a, w->qtx = b, w->sc = sc;
        w->info.duality_gap = (a + b + sc) * (w->s.scaling ? w->cinv : 1.0);
    }
}

static double pri_tol(const opq_workspace *w, double eps_abs, double eps_rel)
{
    double a, b;
    if (w->s.scaling && !w->s.scaled_termination) {
        a = scaled_norm_inf(w->Einv, w->z, w->m);
        b = scaled_norm_inf(w->Einv, w->Axv, w->m);
    } else {
        a = norm_inf(w->z, w->m);
        b = norm_inf(w->Axv, w->m);
    }
    return eps_abs + eps_rel * (a > b ? a : b);
}

static double dua_tol(const opq_workspace *w, double eps_abs, double eps_rel)
{
    double r, t;
    if (w->s.scaling && !w->s.scaled_termination) {
        r = scaled_norm_inf(w->Dinv, w->q, w->n);
        t = scaled_norm_inf(w->Dinv, w->Aty, w->n);
        if (t > r) r = t;
        t = scaled_norm_inf(w->Dinv, w->Pxv, w->n);
        if (t > r) r = t;
        r *= w->cinv;
    } else {
        r = norm_inf(w->q, w->n);
        t = norm_inf(w->Aty, w->n);
        if (t > r) r = t;
        t = norm_inf(w->Pxv, w->n);
        if (t > r) r = t;
    }
    return eps_abs + eps_rel * r;
}

static int is_primal_infeasible(opq_workspace *w, double eps)
{
    int m = w->m;
    const double big = OPQ_INFTY * OPQ_MIN_SCALING;
    for (int i = 0; i < m; i++) {
        if (w->u[i] > big) {
            if (w->l[i] < -big) w->delta_y[i] = 0.0;
            else if (w->delta_y[i] > 0.0) w->delta_y[i] = 0.0;
        } else if (w->l[i] < -big) {
            if (w->delta_y[i] < 0.0) w->delta_y[i] = 0.0;
        }
    }
    double nrm;
    if (w->s.scaling && !w->s.scaled_termination) nrm = scaled_norm_inf(w->E, w->delta_y, m);
    else nrm = norm_inf(w->delta_y, m);
    if (nrm > eps) {
        double lhs = 0.0;
        for (int i = 0; i < m; i++) {
            double dy = w->delta_y[i];
            lhs += w->u[i] * (dy > 0.0 ? dy : 0.0) + w->l[i] * (dy < 0.0 ? dy : 0.0);
        }
        if (lhs < -eps * nrm) {
            csc_tmv(w->n, w->Ap, w->Ai, w->Ax, w->delta_y, w->Atdy, 0, 0);
            if (w->s.scaling && !w->s.scaled_termination)
                for (int j = 0; j < w->n; j++) w->Atdy[j] *= w->Dinv[j];
            return norm_inf(w->Atdy, w->n) < eps * nrm;
        }
    }
    return 0;
}

static int is_dual_infeasible(opq_workspace *w, double eps)
{
    int n = w->n, m = w->m;
    const double big = OPQ_INFTY * OPQ_MIN_SCALING;
    double nrm, cs;
    if (w->s.scaling && !w->s.scaled_termination) {
        nrm = scaled_norm_inf(w->D, w->delta_x, n);
        cs = w->c;
    } else {
        nrm = norm_inf(w->delta_x, n);
        cs = 1.0;
    }
    if (nrm > eps) {
        double qdx = 0.0;
        for (int j = 0; j < n; j++) qdx += w->q[j] * w->delta_x[j];
        if (qdx < -cs * eps * nrm) {
            sym_mv(w, w->delta_x, w->Pdx);
            if (w->s.scaling && !w->s.scaled_termination)
                for (int j = 0; j < n; j++) w->Pdx[j] *= w->Dinv[j];
            if (norm_inf(w->Pdx, n) < cs * eps * nrm) {
                csc_mv(n, w->Ap, w->Ai, w->Ax, w->delta_x, w->Adx, m, 0);
                if (w->s.scaling && !w->s.scaled_termination)
                    for (int i = 0; i < m; i++) w->Adx[i] *= w->Einv[i];
                for (int i = 0; i < m; i++) {
                    if ((w->u[i] < big && w->Adx[i] > eps * nrm) ||
                        (w->l[i] > -big && w->Adx[i] < -eps * nrm))
                        return 0;
                }
                return 1;
            }
        }
    }
    return 0;
}

static int check_termination(opq_workspace *w, int approximate)
{
    double eps_abs = w->s.eps_abs, eps_rel = w->s.eps_rel;
    double eps_pinf = w->s.eps_prim_inf, eps_dinf = w->s.eps_dual_inf;
    int prim_ok = 0, dual_ok = 0, prim_inf = 0, dual_inf = 0, gap_ok = 1;
    if (w->info.pri_res > OPQ_INFTY || w->info.dua_res > OPQ_INFTY) {
        w->info.status = OPQ_NON_CVX;
        w->info.obj_val = NAN;
        return 1;
    }
    if (approximate) eps_abs *= 10, eps_rel *= 10, eps_pinf *= 10, eps_dinf *= 10;
    if (w->m == 0) {
        prim_ok = 1;
    } else {
        if (w->info.pri_res < pri_tol(w, eps_abs, eps_rel)) prim_ok = 1;
        else prim_inf = is_primal_infeasible(w, eps_pinf);
    }
    if (w->info.dua_res < dua_tol(w, eps_abs, eps_rel)) dual_ok = 1;
    else dual_inf = is_dual_infeasible(w, eps_dinf);

    if (w->s.check_dualgap) {
        double rel = fabs(w->xtPx), t = fabs(w->qtx);
        if (t > rel) rel = t;
        t = fabs(w->sc);
        if (t > rel) rel = t;
        if (w->s.scaling) rel *= w->cinv;
        gap_ok = fabs(w->info.duality_gap) < eps_abs + eps_rel * rel;
    }
    if (prim_ok && dual_ok && gap_ok) {
        w->info.status = approximate ? OPQ_SOLVED_INACCURATE : OPQ_SOLVED;
        return 1;
    }
    if (prim_inf) {
        w->info.status = approximate ? OPQ_PRIMAL_INFEASIBLE_INACCURATE : OPQ_PRIMAL_INFEASIBLE;
        w->info.obj_val = OPQ_INFTY;
        return 1;
    }
    if (dual_inf) {
        w->info.status = approximate ? OPQ_DUAL_INFEASIBLE_INACCURATE : OPQ_DUAL_INFEASIBLE;
        w->info.obj_val = -OPQ_INFTY;
        return 1;
    }
    return 0;
}

static double rho_estimate(const opq_workspace *w)
{
    /* z_prev / x_prev hold the (scaled) residual vectors left by update_info */
    int n = w->n, m = w->m;
    double pri = norm_inf(w->z_prev, m), dua = norm_inf(w->x_prev, n);
    double a = norm_inf(w->z, m), b = norm_inf(w->Axv, m);
    pri /= ((a > b ? a : b) + 1e-10);
    a = norm_inf(w->q, n);
    b = norm_inf(w->Aty, n);
    if (b > a) a = b;
    b = norm_inf(w->Pxv, n);
    if (b > a) a = b;
    dua /= (a + 1e-10);
    double r = w->s.rho * sqrt(pri / (dua + 1e-10));
    if (r < OPQ_RHO_MIN) r = OPQ_RHO_MIN;
    if (r > OPQ_RHO_MAX) r = OPQ_RHO_MAX;
    return r;
}

static void update_rho(opq_workspace *w, double rho_new)
{
    if (rho_new < OPQ_RHO_MIN) rho_new = OPQ_RHO_MIN;
    if (rho_new > OPQ_RHO_MAX) rho_new = OPQ_RHO_MAX;
    w->s.rho = rho_new;
    for (int i = 0; i < w->m; i++) {
        if (w->ctype[i] == 0) w->rho_vec[i] = rho_new;
        else if (w->ctype[i] == 1) w->rho_vec[i] = OPQ_RHO_EQ_OVER_RHO_INEQ * rho_new;
        else continue;
        w->rho_inv_vec[i] = 1.0 / w->rho_vec[i];
    }
    kkt_factor(w);
}

int opq_solve(opq_workspace *w, double *x_out, double *y_out, opq_info *info_out)
{
    int n = w->n, m = w->m;
    double alpha = w->s.alpha, sigma = w->s.sigma;
    int iter, can_check = 0;
    if (!w->s.warm_start) opq_cold_start(w);
    w->info.status = OPQ_UNSOLVED;
    w->rho_updates = 0;
    for (iter = 1; iter <= w->s.max_iter; iter++) {
        double *t;
        t = w->x, w->x = w->x_prev, w->x_prev = t;
        t = w->z, w->z = w->z_prev, w->z_prev = t;
        /* (x~, nu) from the KKT system, then z~ = z_prev + (nu - y)/rho */
        for (int j = 0; j < n; j++) w->xz[j] = sigma * w->x_prev[j] - w->q[j];
        for (int i = 0; i < m; i++) w->xz[n + i] = w->z_prev[i] - w->rho_inv_vec[i] * w->y[i];
        {
            double *b = w->xz;
            double *s;
            kkt_solve(w, b);
            s = b; /* b now holds [x~ ; nu] */
            for (int i = 0; i < m; i++)
                s[n + i] = (w->z_prev[i] - w->rho_inv_vec[i] * w->y[i]) + w->rho_inv_vec[i] * s[n + i];
        }
        for (int j = 0; j < n; j++) {
            w->x[j] = alpha * w->xz[j] + (1.0 - alpha) * w->x_prev[j];
            w->delta_x[j] = w->x[j] - w->x_prev[j];
        }
        for (int i = 0; i < m; i++) {
            double zi = alpha * w->xz[n + i] + (1.0 - alpha) * w->z_prev[i] + w->rho_inv_vec[i] * w->y[i];
            if (zi < w->l[i]) zi = w->l[i];
            if (zi > w->u[i]) zi = w->u[i];
            w->z[i] = zi;
        }
        for (int i = 0; i < m; i++) {
            double dy = alpha * w->xz[n + i] + (1.0 - alpha) * w->z_prev[i] - w->z[i];
            dy *= w->rho_vec[i];
            w->delta_y[i] = dy;
            w->y[i] += dy;
        }
        can_check = w->s.check_termination && (iter % w->s.check_termination == 0);
        if (can_check) {
            update_info(w, iter);
            if (check_termination(w, 0)) break;
        }
        if (w->s.adaptive_rho && w->s.adaptive_rho_interval &&
            (iter % w->s.adaptive_rho_interval == 0)) {
            if (!can_check) update_info(w, iter);
            double rn = rho_estimate(w);
            w->info.rho_estimate = rn;
            if (rn > w->s.rho * w->s.adaptive_rho_tolerance ||
                rn < w->s.rho / w->s.adaptive_rho_tolerance) {
                update_rho(w, rn);
                w->rho_updates++;
            }
        }
    }
    if (iter > w->s.max_iter) iter = w->s.max_iter;
    if (!can_check) {
        update_info(w, iter);
        check_termination(w, 0);
    }
    if (w->info.status == OPQ_UNSOLVED) {
        if (!check_termination(w, 1)) w->info.status = OPQ_MAX_ITER_REACHED;
    }
    w->info.iter = iter;
    w->info.rho_updates = w->rho_updates;
    w->info.rho_estimate = rho_estimate(w);
    if (x_out)
        for (int j = 0; j < n; j++) x_out[j] = w->D[j] * w->x[j];
    if (y_out)
        for (int i = 0; i < m; i++) y_out[i] = w->E[i] * w->y[i] * w->cinv;
    if (info_out) *info_out = w->info;
    return w->info.status;
}

const double *opq_get_vec(const opq_workspace *w, const char *name, int *len)
{
    int n = w->n, m = w->m;
#define RET(field, ln) do { if (len) *len = (ln); return w->field; } while (0)
    if (!strcmp(name, "x")) RET(x, n);
    if (!strcmp(name, "z")) RET(z, m);
    if (!strcmp(name, "y")) RET(y, m);
    if (!strcmp(name, "D")) RET(D, n);
    if (!strcmp(name, "E")) RET(E, m);
    if (!strcmp(name, "q")) RET(q, n);
    if (!strcmp(name, "l")) RET(l, m);
    if (!strcmp(name, "u")) RET(u, m);
    if (!strcmp(name, "rho_vec")) RET(rho_vec, m);
    if (!strcmp(name, "Ax")) RET(Ax, w->Annz);
    if (!strcmp(name, "Px")) RET(Px, w->Pnnz);
#undef RET
    if (len) *len = 0;
    return NULL;
}

double opq_get_scalar(const opq_workspace *w, const char *name)
{
    if (!strcmp(name, "c")) return w->c;
    if (!strcmp(name, "rho")) return w->s.rho;
    if (!strcmp(name, "L_nnz")) return (double)w->Lp[w->N];
    return NAN;
}
