/*
 * oracle/osqp_port.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, FP64) of the OSQP ADMM QP solver, the third-party
 * dependency the reference calls at
 *   /root/reference/src/acmpc/control/solvers/control.py:88-106
 *   /root/reference/src/acmpc/control/solvers/speed_profile.py:68-86,146
 * (`osqp` on PyPI, un-pinned in /root/reference/requirements.txt:2, absent from
 * /root/reference and from this image).  The algorithm restated is OSQP 0.6.x
 * (Stellato et al., "OSQP: an operator splitting solver for quadratic programs",
 * Math. Prog. Comp. 2020): Ruiz equilibration + cost scaling, per-constraint rho,
 * quasi-definite KKT solved by a sparse LDL^T, alpha-relaxed ADMM, unscaled
 * residual termination test every `check_termination` iterations, primal/dual
 * infeasibility certificates, adaptive rho at a FIXED iteration interval.
 *
 * PARITY UNPINNED: the reference ships no golden vector for this path and no
 * OSQP wheel can be run here, so this port is pinned only against (a) an
 * independent dense numpy ADMM written from the same paper and (b) the exact
 * optimum from HiGHS at tight tolerances (tests/test_oracle_*.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may link or call this.
 */
#ifndef ORACLE_OSQP_PORT_H
#define ORACLE_OSQP_PORT_H

#ifdef __cplusplus
extern "C" {
#endif

#define OPQ_INFTY 1e30
#define OPQ_MIN_SCALING 1e-4
#define OPQ_MAX_SCALING 1e4
#define OPQ_RHO_MIN 1e-6
#define OPQ_RHO_MAX 1e6
#define OPQ_RHO_EQ_OVER_RHO_INEQ 1e3
#define OPQ_RHO_TOL 1e-4

enum {
    OPQ_SOLVED = 1,
    OPQ_SOLVED_INACCURATE = 2,
    OPQ_PRIMAL_INFEASIBLE_INACCURATE = 3,
    OPQ_DUAL_INFEASIBLE_INACCURATE = 4,
    OPQ_MAX_ITER_REACHED = -2,
    OPQ_PRIMAL_INFEASIBLE = -3,
    OPQ_DUAL_INFEASIBLE = -4,
    OPQ_NON_CVX = -7,
    OPQ_UNSOLVED = -10
};

typedef struct {
    double rho, sigma, alpha;
    double eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
    double adaptive_rho_tolerance;
    int scaling;               /* Ruiz passes (10) */
    int max_iter;              /* 4000 */
    int check_termination;     /* 25 */
    int adaptive_rho;          /* 1 */
    int adaptive_rho_interval; /* fixed interval; OSQP's 0 = "timing based" is not reproducible */
    int warm_start;            /* 1: keep x,z,y between solves (OSQP default) */
    int scaled_termination;    /* 0 */
    int check_dualgap;         /* 0 = OSQP 0.6.x termination; 1 = OSQP 1.x adds the duality-gap test (restated from the
                                  1.0 sources as remembered -- UNVERIFIED until tools/pin_osqp.py meets a 1.x wheel) */
} opq_settings;

typedef struct {
    int status;
    int iter;
    int rho_updates;
    double obj_val, pri_res, dua_res, rho_estimate;
    double duality_gap;        /* x'Px + q'x + SC(y), unscaled (only meaningful with check_dualgap) */
} opq_info;

typedef struct opq_workspace opq_workspace;

void opq_default_settings(opq_settings *s);

/* P: upper triangle (incl. diagonal) in CSC; A: CSC.  Data are copied. */
opq_workspace *opq_setup(int n, int m,
                         const int *Pp, const int *Pi, const double *Px,
                         const double *q,
                         const int *Ap, const int *Ai, const double *Ax,
                         const double *l, const double *u,
                         const opq_settings *settings);
void opq_free(opq_workspace *w);

/* osqp.update(q=, l=, u=, Ax=): any pointer may be NULL (= keep).  Ax must
 * have the setup sparsity.  Re-equilibrates and refactors like osqp_update_A. */
int opq_update(opq_workspace *w, const double *q, const double *l, const double *u,
               const double *Ax);
/* forget iterates (x=z=y=0) and restore rho to the setup value */
void opq_cold_start(opq_workspace *w);
/* osqp.warm_start(x=, y=): unscaled values */
void opq_warm_start(opq_workspace *w, const double *x, const double *y);

/* returns status; x (n), y (m) may be NULL */
int opq_solve(opq_workspace *w, double *x, double *y, opq_info *info);

/* introspection for tests: scaled iterates / scaling vectors (pointers into w) */
const double *opq_get_vec(const opq_workspace *w, const char *name, int *len);
double opq_get_scalar(const opq_workspace *w, const char *name);
const char *opq_status_string(int status);

#ifdef __cplusplus
}
#endif
#endif
