/*
 * mpc_oracle.h — CPU FP64 restatement of the ros2_mpc per-control-step NMPC solve.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ros2_mpc_b200/) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker /
 * the timed CPU baseline.
 *
 * PARITY STATUS: the NLP (cost, dynamics, bounds) is pinned against the
 * reference's own unmodified Python sources executed here through a tracing
 * casadi stand-in (tests/golden/make_golden.py); the optimum is pinned
 * against scipy SLSQP/trust-constr.  The *solver* (CasADi 3.6.3 -> IPOPT ->
 * MUMPS, requirements.txt:1) is NOT installable offline, so the interior
 * point iteration below restates IPOPT's published algorithm (Waechter &
 * Biegler 2006) and documented defaults: "solver parity unpinned".
 *
 * Reference files restated (under /root/reference):
 *   ros2_mpc/planner/local_planner_point_stabilization.py:99-178  (variant B)
 *   ros2_mpc/mpc_point_stabilization.py:46-149                    (variant A)
 *   ros2_mpc/planner/local_planner_tracking.py:55-178             (variant C)
 *   config/params.yaml:1-11
 */
#ifndef MPC_ORACLE_H
#define MPC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_N 256

/* integrator */
#define ORC_RK4 0
#define ORC_EULER 1
/* obstacle cost form */
#define ORC_OBS_NONE 0
#define ORC_OBS_GAUSS 1  /* c*exp(-s)   : variant B form, local_planner_point_stabilization.py:60-67 */
#define ORC_OBS_EXPLOG 2 /* exp(c/s)    : variant A/C form, mpc_point_stabilization.py:46-53 */
/* reference kind */
#define ORC_REF_GOAL 0 /* xref[3]                       (A, B) */
#define ORC_REF_TRAJ 1 /* xref[3N], uref[2N]            (C)    */

/* IPOPT ApplicationReturnStatus values that can occur */
#define ORC_SOLVE_SUCCEEDED 0
#define ORC_SOLVED_TO_ACCEPTABLE_LEVEL 1
#define ORC_SEARCH_DIRECTION_TOO_SMALL 3
#define ORC_MAXITER_EXCEEDED (-1)
#define ORC_RESTORATION_FAILED (-2)
#define ORC_ERROR_IN_STEP_COMPUTATION (-3)
#define ORC_INVALID_NUMBER_DETECTED (-13)

typedef struct {
    int N;          /* horizon (params.yaml:2) */
    int M;          /* obstacle slots, int(costmap_size*2/resolution)*2 */
    double dt;      /* params.yaml:1 */
    int integrator; /* ORC_RK4 / ORC_EULER */
    double Q[3];    /* diagonal state weights */
    double R[2];    /* diagonal control weights */
    double kappa;   /* exponent of the reverse penalty (1/exp(v))**kappa */
    int ref_kind;
    int obs_form;
    double obs_c;   /* obstacle cost factor */
    double obs_r;   /* inflation radius */
    int obs_k0, obs_k1; /* stages k0..k1 (inclusive) carry the obstacle sum */
    double u_lo[2], u_hi[2];
    /* solver options (IPOPT names) */
    double tol;             /* 1e-8 */
    int max_iter;           /* 3000 */
    double acceptable_tol;  /* 1e-6 */
    int acceptable_iter;    /* 15 */
    double mu_init;         /* 0.1 */
    int max_soc;            /* 4 */
    int linear_solver;      /* 0 = stage-wise Riccati, 1 = dense Bunch-Kaufman LDL^T on the full KKT */
} orc_params;

typedef struct {
    int iters;      /* accepted interior-point iterations */
    int ls_extra;   /* trial-point evaluations beyond the first of each iteration */
    int n_soc;      /* second-order-correction solves */
    int n_reg;      /* Hessian regularisation (inertia correction) re-factorisations */
    int n_resto;    /* rollout restorations */
    double mu;      /* final barrier parameter */
    double err;     /* final scaled NLP error E_0 */
    double obj_scale;
} orc_stats;

void orc_default_options(orc_params *p);

/* One solve.  x0[3]; xref[3] or [3N]; uref[2N] or NULL; obs_x/obs_y[M] (NULL when obs_form==NONE);
 * u_init[2N] (NULL = zeros); x_init[3(N+1)] stage-major or NULL (= zeros, the Opti default).
 * Outputs: X_out[3(N+1)] stage-major (X_out[3k+i] = X[i,k], X_out[0..2] == x0), U_out[2N], cost_out[1].
 * Returns the IPOPT-style status. */
int orc_solve(const orc_params *p, const double *x0, const double *xref, const double *uref,
              const double *obs_x, const double *obs_y, const double *u_init, const double *x_init,
              double *X_out, double *U_out, double *cost_out, orc_stats *stats);

/* Batch of B independent solves on nthreads host threads.  Arrays are problem-major
 * (problem b's x0 at x0 + 3b, ...).  obs_stride = M (per-problem lists) or 0 (shared list). */
int orc_solve_batch(const orc_params *p, int B, const double *x0, const double *xref, const double *uref,
                    const double *obs_x, const double *obs_y, int obs_stride, const double *u_init,
                    double *X_out, double *U_out, double *cost_out, int *status_out, int *iters_out,
                    int *ls_out, int nthreads);

/* NLP function evaluation at a point (for derivative / restatement tests).
 * X[3(N+1)] (X[0..2] must be x0), U[2N], lam[3N] multipliers of c_k = X_k - F(X_{k-1},U_{k-1}), k=1..N.
 * f_out: objective; c_out[3N]: defects; grad_out[5N] = (dX_1..dX_N (3N), dU_0..dU_{N-1} (2N));
 * stage_out[(N+1)*36]: per stage k: a13,a23,b11,b12,b21,b22, H[25] row-major (x,y,th,v,w) of the
 * Lagrangian Hessian, c_{k+1}[3], pad[2]. Any output pointer may be NULL. */
void orc_eval(const orc_params *p, const double *x0, const double *xref, const double *uref,
              const double *obs_x, const double *obs_y, const double *X, const double *U,
              const double *lam, double obj_scale, double *f_out, double *c_out, double *grad_out,
              double *stage_out);

/* Dense symmetric-indefinite LDL^T (Bunch-Kaufman) used by linear_solver=1; exported for its own test.
 * A: n*n row-major, lower triangle used, overwritten. Returns 0 ok / k+1 if exactly singular at k.
 * inertia[3] = (#positive, #negative, #zero). b[n] overwritten with the solution. */
int orc_ldl_solve(int n, double *A, double *b, int *inertia);

#ifdef __cplusplus
}
#endif
#endif
