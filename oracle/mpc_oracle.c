/*
 * mpc_oracle.c — CPU FP64 restatement of the ros2_mpc NMPC solve (see mpc_oracle.h header:
 * TEST INFRASTRUCTURE ONLY; solver parity unpinned because CasADi/IPOPT is not installable here).
 *
 * Part 1 restates the NLP the reference builds with casadi.Opti:
 *   dynamics      get_system_function  local_planner_point_stabilization.py:159-178
 *   RK4 defects   rk4                  local_planner_point_stabilization.py:136-148
 *   Euler defects euler_integration    local_planner_tracking.py:132-137
 *   stage cost    define_cost_function local_planner_point_stabilization.py:104-127,
 *                                      mpc_point_stabilization.py:85-100, local_planner_tracking.py:106-130
 *   obstacle cost define_obstacles_cost_function  mpc_point_stabilization.py:46-53 (exp(c*exp(-log s))),
 *                                      local_planner_point_stabilization.py:60-67 (c*exp(-s))
 *   bounds        constraints          local_planner_point_stabilization.py:99-102 (two-sided inequalities on U)
 * Part 2 restates the solver behind opti.solver("ipopt") / opti.solve()
 *   (local_planner_point_stabilization.py:56-57,84): a primal-dual interior point method with IPOPT's
 *   slack formulation, monotone barrier update, fraction-to-the-boundary rule, filter line search with
 *   second-order correction, inertia correction and IPOPT's default constants.
 *   Deviations from IPOPT proper are listed in DESIGN.md ("oracle deviations").
 */
#include "mpc_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* Part 1: the NLP                                                                             */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    const orc_params *p;
    const double *x0, *xref, *uref, *ox, *oy;
} prob_t;

typedef struct {
    double a13, a23, b11, b12, b21, b22; /* A = I + a13 e1e3' + a23 e2e3',  B = [[b11,b12],[b21,b22],[0,dt]] */
    double H[5][5];                      /* Lagrangian Hessian block on (x,y,th,v,w) */
    double g[5];                         /* (scaled) objective gradient */
    double c[3];                         /* c_{k+1} = X_{k+1} - F(X_k,U_k) */
} stage_t;

/* continuous unicycle rhs, get_system_function :174 */
static void f_cont(const double x[3], const double u[2], double o[3]) {
    o[0] = u[0] * cos(x[2]);
    o[1] = u[0] * sin(x[2]);
    o[2] = u[1];
}

/* one integration step, literally as the reference stages it (rk4 :139-146 / euler :134) */
static void dyn_step(const orc_params *p, const double x[3], const double u[2], double xn[3]) {
    double dt = p->dt;
    if (p->integrator == ORC_EULER) {
        double k1[3];
        f_cont(x, u, k1);
        for (int i = 0; i < 3; i++) xn[i] = x[i] + dt * k1[i];
        return;
    }
    double k1[3], k2[3], k3[3], k4[3], t[3];
    f_cont(x, u, k1);
    for (int i = 0; i < 3; i++) t[i] = x[i] + dt / 2 * k1[i];
    f_cont(t, u, k2);
    for (int i = 0; i < 3; i++) t[i] = x[i] + dt / 2 * k2[i];
    f_cont(t, u, k3);
    for (int i = 0; i < 3; i++) t[i] = x[i] + dt * k3[i];
    f_cont(t, u, k4);
    for (int i = 0; i < 3; i++) xn[i] = x[i] + dt / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
}

/* first and second derivatives of the step map (closed forms).
 * d2x / d2y: second derivatives of F_x, F_y w.r.t. (th,v,w) in the order tt,tv,tw,vv,vw,ww. */
static void dyn_derivs(const orc_params *p, const double x[3], const double u[2], stage_t *s, double d2x[6],
                       double d2y[6]) {
    double dt = p->dt, th = x[2], v = u[0], w = u[1];
    if (p->integrator == ORC_EULER) {
        double c = cos(th), sn = sin(th);
        s->a13 = -dt * v * sn;
        s->a23 = dt * v * c;
        s->b11 = dt * c;
        s->b21 = dt * sn;
        s->b12 = 0;
        s->b22 = 0;
        d2x[0] = -dt * v * c; d2x[1] = -dt * sn; d2x[2] = 0; d2x[3] = 0; d2x[4] = 0; d2x[5] = 0;
        d2y[0] = -dt * v * sn; d2y[1] = dt * c; d2y[2] = 0; d2y[3] = 0; d2y[4] = 0; d2y[5] = 0;
        return;
    }
    double tm = th + dt * w / 2, te = th + dt * w;
    double c0 = cos(th), s0 = sin(th), cm = cos(tm), sm = sin(tm), ce = cos(te), se = sin(te);
    double C = c0 + 4 * cm + ce, S = s0 + 4 * sm + se;
    double C1 = 2 * cm + ce, S1 = 2 * sm + se;
    double C2 = cm + ce, S2 = sm + se;
    double h = dt / 6;
    s->a13 = -h * v * S;
    s->a23 = h * v * C;
    s->b11 = h * C;
    s->b21 = h * S;
    s->b12 = -h * dt * v * S1;
    s->b22 = h * dt * v * C1;
    d2x[0] = -h * v * C;       d2y[0] = -h * v * S;
    d2x[1] = -h * S;           d2y[1] = h * C;
    d2x[2] = -h * dt * v * C1; d2y[2] = -h * dt * v * S1;
    d2x[3] = 0;                d2y[3] = 0;
    d2x[4] = -h * dt * S1;     d2y[4] = h * dt * C1;
    d2x[5] = -h * dt * dt * v * C2; d2y[5] = -h * dt * dt * v * S2;
}

/* obstacle sum at one predicted position; g2 = d/d(x,y), h3 = (xx,xy,yy); either may be NULL */
static double obstacle_sum(const prob_t *q, double x, double y, double *g2, double *h3) {
    const orc_params *p = q->p;
    double r = p->obs_r, c = p->obs_c, r2 = r * r, val = 0;
    double gx = 0, gy = 0, hxx = 0, hxy = 0, hyy = 0;
    for (int j = 0; j < p->M; j++) {
        double dx = x - q->ox[j], dy = y - q->oy[j];
        if (p->obs_form == ORC_OBS_EXPLOG) {
            /* mpc_point_stabilization.py:50-52: hxy = log((dx/r)^2+(dy/r)^2); obj += exp(c*exp(-hxy)) */
            double s = (dx / r) * (dx / r) + (dy / r) * (dy / r);
            double e = exp(c * exp(-log(s)));
            val += e;
            if (g2 || h3) {
                double p1 = -(c / (s * s)) * e;
                double p2 = c * (2 * s + c) / (s * s * s * s) * e;
                double sx = 2 * dx / r2, sy = 2 * dy / r2;
                gx += p1 * sx;
                gy += p1 * sy;
                hxx += p2 * sx * sx + p1 * 2 / r2;
                hxy += p2 * sx * sy;
                hyy += p2 * sy * sy + p1 * 2 / r2;
            }
        } else {
            /* local_planner_point_stabilization.py:64-66 */
            double s = (dx * dx + dy * dy) / r2;
            double e = c * exp(-s);
            val += e;
            if (g2 || h3) {
                double sx = 2 * dx / r2, sy = 2 * dy / r2;
                gx += -e * sx;
                gy += -e * sy;
                hxx += e * sx * sx - e * 2 / r2;
                hxy += e * sx * sy;
                hyy += e * sy * sy - e * 2 / r2;
            }
        }
    }
    if (g2) { g2[0] = gx; g2[1] = gy; }
    if (h3) { h3[0] = hxx; h3[1] = hxy; h3[2] = hyy; }
    return val;
}

static void ref_at(const prob_t *q, int k, double r[3], double ub[2]) {
    if (q->p->ref_kind == ORC_REF_GOAL) {
        r[0] = q->xref[0]; r[1] = q->xref[1]; r[2] = q->xref[2];
        ub[0] = 0; ub[1] = 0;
    } else {
        /* local_planner_tracking.py:118-122: state ref P_X[3(k+1):3(k+1)+3] = pf[3k:3k+3], control ref P_U[2k:2k+2] */
        for (int i = 0; i < 3; i++) r[i] = q->xref[3 * k + i];
        ub[0] = q->uref[2 * k]; ub[1] = q->uref[2 * k + 1];
    }
}

static int obs_stage(const orc_params *p, int k) {
    return p->obs_form != ORC_OBS_NONE && k >= p->obs_k0 && k <= p->obs_k1;
}

/* objective value (unscaled) */
static double eval_f(const prob_t *q, const double *X, const double *U) {
    const orc_params *p = q->p;
    double f = 0;
    for (int k = 0; k < p->N; k++) {
        double r[3], ub[2];
        ref_at(q, k, r, ub);
        const double *x = X + 3 * k, *u = U + 2 * k;
        double t = 0;
        for (int i = 0; i < 3; i++) t += (x[i] - r[i]) * p->Q[i] * (x[i] - r[i]);
        for (int i = 0; i < 2; i++) t += (u[i] - ub[i]) * p->R[i] * (u[i] - ub[i]);
        t += pow(1.0 / exp(u[0]), p->kappa); /* (1/exp(v))**kappa */
        f += t;
    }
    for (int k = 0; k <= p->N; k++)
        if (obs_stage(p, k)) f += obstacle_sum(q, X[3 * k], X[3 * k + 1], NULL, NULL);
    return f;
}

/* defects c_{k+1} = X_{k+1} - F(X_k,U_k), k=0..N-1, written to c[3(k+1)..] (c[0..2] unused) */
static void eval_c(const prob_t *q, const double *X, const double *U, double *c) {
    const orc_params *p = q->p;
    for (int k = 0; k < p->N; k++) {
        double xn[3];
        dyn_step(p, X + 3 * k, U + 2 * k, xn);
        for (int i = 0; i < 3; i++) c[3 * (k + 1) + i] = X[3 * (k + 1) + i] - xn[i];
    }
}

/* full stage data at a point: gradient of df*f, defects, Jacobians, Hessian of df*f + sum lam'c.
 * lam indexed lam[3k..] for k=1..N. */
static void eval_stages(const prob_t *q, const double *X, const double *U, const double *lam, double df,
                        stage_t *st) {
    const orc_params *p = q->p;
    int N = p->N;
    for (int k = 0; k <= N; k++) {
        stage_t *s = &st[k];
        memset(s, 0, sizeof(*s));
        const double *x = X + 3 * k;
        if (k < N) {
            const double *u = U + 2 * k;
            const double *l = lam + 3 * (k + 1);
            double r[3], ub[2], d2x[6], d2y[6], xn[3];
            ref_at(q, k, r, ub);
            dyn_derivs(p, x, u, s, d2x, d2y);
            dyn_step(p, x, u, xn);
            for (int i = 0; i < 3; i++) s->c[i] = X[3 * (k + 1) + i] - xn[i];
            for (int i = 0; i < 3; i++) {
                s->g[i] = df * 2 * p->Q[i] * (x[i] - r[i]);
                s->H[i][i] = df * 2 * p->Q[i];
            }
            double e = exp(-p->kappa * u[0]);
            s->g[3] = df * (2 * p->R[0] * (u[0] - ub[0]) - p->kappa * e);
            s->g[4] = df * 2 * p->R[1] * (u[1] - ub[1]);
            s->H[3][3] = df * (2 * p->R[0] + p->kappa * p->kappa * e);
            s->H[4][4] = df * 2 * p->R[1];
            /* - lam_{k+1}' d2F */
            double tt = -(l[0] * d2x[0] + l[1] * d2y[0]);
            double tv = -(l[0] * d2x[1] + l[1] * d2y[1]);
            double tw = -(l[0] * d2x[2] + l[1] * d2y[2]);
            double vv = -(l[0] * d2x[3] + l[1] * d2y[3]);
            double vw = -(l[0] * d2x[4] + l[1] * d2y[4]);
            double ww = -(l[0] * d2x[5] + l[1] * d2y[5]);
            s->H[2][2] += tt;
            s->H[2][3] += tv; s->H[3][2] += tv;
            s->H[2][4] += tw; s->H[4][2] += tw;
            s->H[3][3] += vv;
            s->H[3][4] += vw; s->H[4][3] += vw;
            s->H[4][4] += ww;
        }
        if (obs_stage(p, k)) {
            double g2[2], h3[3];
            obstacle_sum(q, x[0], x[1], g2, h3);
            s->g[0] += df * g2[0];
            s->g[1] += df * g2[1];
            s->H[0][0] += df * h3[0];
            s->H[0][1] += df * h3[1];
            s->H[1][0] += df * h3[1];
            s->H[1][1] += df * h3[2];
        }
    }
}

void orc_eval(const orc_params *p, const double *x0, const double *xref, const double *uref,
              const double *obs_x, const double *obs_y, const double *X, const double *U, const double *lam,
              double obj_scale, double *f_out, double *c_out, double *grad_out, double *stage_out) {
    prob_t q = {p, x0, xref, uref, obs_x, obs_y};
    int N = p->N;
    if (f_out) *f_out = eval_f(&q, X, U);
    if (c_out) {
        double *c = (double *)calloc(3 * (N + 1), sizeof(double));
        eval_c(&q, X, U, c);
        memcpy(c_out, c + 3, sizeof(double) * 3 * N);
        free(c);
    }
    if (grad_out || stage_out) {
        stage_t *st = (stage_t *)malloc(sizeof(stage_t) * (N + 1));
        double *l = (double *)calloc(3 * (N + 1), sizeof(double));
        if (lam) memcpy(l + 3, lam, sizeof(double) * 3 * N);
        eval_stages(&q, X, U, l, obj_scale, st);
        if (grad_out) {
            for (int k = 1; k <= N; k++)
                for (int i = 0; i < 3; i++) grad_out[3 * (k - 1) + i] = st[k].g[i];
            for (int k = 0; k < N; k++)
                for (int i = 0; i < 2; i++) grad_out[3 * N + 2 * k + i] = st[k].g[3 + i];
        }
        if (stage_out) {
            for (int k = 0; k <= N; k++) {
                double *o = stage_out + 36 * k;
                o[0] = st[k].a13; o[1] = st[k].a23; o[2] = st[k].b11; o[3] = st[k].b12;
                o[4] = st[k].b21; o[5] = st[k].b22;
                for (int i = 0; i < 5; i++)
                    for (int j = 0; j < 5; j++) o[6 + 5 * i + j] = st[k].H[i][j];
                for (int i = 0; i < 3; i++) o[31 + i] = st[k].c[i];
                o[34] = o[35] = 0;
            }
        }
        free(st);
        free(l);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Dense Bunch-Kaufman LDL^T (stand-in for MUMPS: factor, inertia, solve)                      */
/* ------------------------------------------------------------------------------------------ */

#define AT(i, j) A[(size_t)(i) * n + (j)]

static int bk_factor(int n, double *A, int *ipiv) {
    const double alpha = (1.0 + sqrt(17.0)) / 8.0;
    int info = 0, k = 0;
    while (k < n) {
        int kstep = 1, kp = k, imax = k;
        double absakk = fabs(AT(k, k)), colmax = 0;
        for (int i = k + 1; i < n; i++)
            if (fabs(AT(i, k)) > colmax) { colmax = fabs(AT(i, k)); imax = i; }
        if (fmax(absakk, colmax) == 0.0) {
            if (!info) info = k + 1;
            kp = k;
        } else {
            if (absakk >= alpha * colmax) {
                kp = k;
            } else {
                double rowmax = 0;
                for (int j = k; j < imax; j++) rowmax = fmax(rowmax, fabs(AT(imax, j)));
                for (int i = imax + 1; i < n; i++) rowmax = fmax(rowmax, fabs(AT(i, imax)));
                if (absakk >= alpha * colmax * (colmax / rowmax)) kp = k;
                else if (fabs(AT(imax, imax)) >= alpha * rowmax) kp = imax;
                else { kp = imax; kstep = 2; }
            }
            int kk = k + kstep - 1;
            if (kp != kk) {
                for (int i = kp + 1; i < n; i++) { double t = AT(i, kk); AT(i, kk) = AT(i, kp); AT(i, kp) = t; }
                for (int j = kk + 1; j < kp; j++) { double t = AT(j, kk); AT(j, kk) = AT(kp, j); AT(kp, j) = t; }
                { double t = AT(kk, kk); AT(kk, kk) = AT(kp, kp); AT(kp, kp) = t; }
                if (kstep == 2) { double t = AT(k + 1, k); AT(k + 1, k) = AT(kp, k); AT(kp, k) = t; }
            }
            if (kstep == 1) {
                double d11 = 1.0 / AT(k, k);
                for (int j = k + 1; j < n; j++) {
                    double wj = d11 * AT(j, k);
                    if (wj != 0.0)
                        for (int i = j; i < n; i++) AT(i, j) -= AT(i, k) * wj;
                }
                for (int i = k + 1; i < n; i++) AT(i, k) *= d11;
            } else if (k < n - 2) {
                double d21 = AT(k + 1, k);
                double d11 = AT(k + 1, k + 1) / d21, d22 = AT(k, k) / d21;
                double t = 1.0 / (d11 * d22 - 1.0);
                d21 = t / d21;
                for (int j = k + 2; j < n; j++) {
                    double wk = d21 * (d11 * AT(j, k) - AT(j, k + 1));
                    double wkp1 = d21 * (d22 * AT(j, k + 1) - AT(j, k));
                    for (int i = j; i < n; i++) AT(i, j) -= AT(i, k) * wk + AT(i, k + 1) * wkp1;
                    AT(j, k) = wk;
                    AT(j, k + 1) = wkp1;
                }
            }
        }
        if (kstep == 1) ipiv[k] = kp;
        else { ipiv[k] = -(kp + 1); ipiv[k + 1] = -(kp + 1); }
        k += kstep;
    }
    return info;
}

static void bk_inertia(int n, const double *A, const int *ipiv, int inertia[3]) {
    inertia[0] = inertia[1] = inertia[2] = 0;
    int k = 0;
    while (k < n) {
        if (ipiv[k] >= 0) {
            double d = AT(k, k);
            if (d > 0) inertia[0]++; else if (d < 0) inertia[1]++; else inertia[2]++;
            k++;
        } else {
            double a = AT(k, k), b = AT(k + 1, k), c = AT(k + 1, k + 1);
            double det = a * c - b * b, tr = a + c;
            if (det < 0) { inertia[0]++; inertia[1]++; }
            else if (det > 0) { if (tr > 0) inertia[0] += 2; else inertia[1] += 2; }
            else { inertia[2]++; if (tr > 0) inertia[0]++; else if (tr < 0) inertia[1]++; else inertia[2]++; }
            k += 2;
        }
    }
}

static void bk_solve(int n, const double *A, const int *ipiv, double *b) {
    int k = 0;
    while (k < n) {
        if (ipiv[k] >= 0) {
            int kp = ipiv[k];
            if (kp != k) { double t = b[k]; b[k] = b[kp]; b[kp] = t; }
            for (int i = k + 1; i < n; i++) b[i] -= AT(i, k) * b[k];
            b[k] /= AT(k, k);
            k++;
        } else {
            int kp = -ipiv[k] - 1;
            if (kp != k + 1) { double t = b[k + 1]; b[k + 1] = b[kp]; b[kp] = t; }
            for (int i = k + 2; i < n; i++) b[i] -= AT(i, k) * b[k] + AT(i, k + 1) * b[k + 1];
            double akm1k = AT(k + 1, k);
            double akm1 = AT(k, k) / akm1k, ak = AT(k + 1, k + 1) / akm1k;
            double denom = akm1 * ak - 1.0;
            double bkm1 = b[k] / akm1k, bk = b[k + 1] / akm1k;
            b[k] = (ak * bkm1 - bk) / denom;
            b[k + 1] = (akm1 * bk - bkm1) / denom;
            k += 2;
        }
    }
    k = n - 1;
    while (k >= 0) {
        if (ipiv[k] >= 0) {
            double s = 0;
            for (int i = k + 1; i < n; i++) s += AT(i, k) * b[i];
            b[k] -= s;
            int kp = ipiv[k];
            if (kp != k) { double t = b[k]; b[k] = b[kp]; b[kp] = t; }
            k--;
        } else {
            double s0 = 0, s1 = 0;
            for (int i = k + 1; i < n; i++) { s0 += AT(i, k) * b[i]; s1 += AT(i, k - 1) * b[i]; }
            b[k] -= s0;
            b[k - 1] -= s1;
            int kp = -ipiv[k] - 1;
            if (kp != k) { double t = b[k]; b[k] = b[kp]; b[kp] = t; }
            k -= 2;
        }
    }
}
#undef AT

int orc_ldl_solve(int n, double *A, double *b, int *inertia) {
    int *ipiv = (int *)malloc(sizeof(int) * n);
    int info = bk_factor(n, A, ipiv);
    if (inertia) bk_inertia(n, A, ipiv, inertia);
    if (!info && b) bk_solve(n, A, ipiv, b);
    free(ipiv);
    return info;
}

/* ------------------------------------------------------------------------------------------ */
/* Part 2: interior point method                                                               */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    double *dX, *dlam; /* 3(N+1), index 3k+i, k=1..N used */
    double *dU, *dS, *dyd; /* 2N */
} step_t;

typedef struct {
    int N;
    /* iterate */
    double *X, *U, *S, *lam, *yd, *vL, *vU;
    /* primal-dual steps: the Newton step and the second-order-corrected one */
    step_t main, soc;
    double *dvL, *dvU;
    /* trial */
    double *Xt, *Ut, *St, *ct;
    /* stage data and right-hand sides */
    stage_t *st;
    double *rxX, *rxU, *rs, *rc, *rd, *Dsig, *csoc, *dsoc;
    /* Riccati factors */
    double *K, *kf, *P, *pv;
    /* dense */
    double *Kd, *bd;
    int *ipiv;
    double sL[2], sU[2];
    void *block;
} ws_t;

static double *carve(double **cur, size_t n) {
    double *r = *cur;
    *cur += n;
    return r;
}

static void ws_alloc(ws_t *w, int N, int dense) {
    size_t n3 = 3 * (size_t)(N + 1), n2 = 2 * (size_t)N;
    size_t total = 2 * n3 + 5 * n2 + 2 * (2 * n3 + 3 * n2) + 2 * n2 + 2 * n3 + 2 * n2 + 3 * n3 + 5 * n2 +
                   (6 + 2 + 9 + 3) * (size_t)(N + 1) + 64;
    double *blk = (double *)calloc(total, sizeof(double));
    double *c = blk;
    w->block = blk;
    w->N = N;
    w->X = carve(&c, n3); w->lam = carve(&c, n3);
    w->U = carve(&c, n2); w->S = carve(&c, n2); w->yd = carve(&c, n2); w->vL = carve(&c, n2); w->vU = carve(&c, n2);
    step_t *sp[2] = {&w->main, &w->soc};
    for (int i = 0; i < 2; i++) {
        sp[i]->dX = carve(&c, n3); sp[i]->dlam = carve(&c, n3);
        sp[i]->dU = carve(&c, n2); sp[i]->dS = carve(&c, n2); sp[i]->dyd = carve(&c, n2);
    }
    w->dvL = carve(&c, n2); w->dvU = carve(&c, n2);
    w->Xt = carve(&c, n3); w->ct = carve(&c, n3); w->Ut = carve(&c, n2); w->St = carve(&c, n2);
    w->rxX = carve(&c, n3); w->rc = carve(&c, n3); w->csoc = carve(&c, n3);
    w->rxU = carve(&c, n2); w->rs = carve(&c, n2); w->rd = carve(&c, n2); w->Dsig = carve(&c, n2); w->dsoc = carve(&c, n2);
    w->K = carve(&c, 6 * (size_t)(N + 1)); w->kf = carve(&c, 2 * (size_t)(N + 1));
    w->P = carve(&c, 9 * (size_t)(N + 1)); w->pv = carve(&c, 3 * (size_t)(N + 1));
    w->st = (stage_t *)calloc(N + 1, sizeof(stage_t));
    w->Kd = NULL; w->bd = NULL; w->ipiv = NULL;
    if (dense) {
        size_t n = 12 * (size_t)N;
        w->Kd = (double *)malloc(sizeof(double) * n * n);
        w->bd = (double *)malloc(sizeof(double) * n);
        w->ipiv = (int *)malloc(sizeof(int) * n);
    }
}

static void ws_free(ws_t *w) {
    free(w->block);
    free(w->st);
    free(w->Kd);
    free(w->bd);
    free(w->ipiv);
}

/*
 * Augmented system (IPOPT PDFullSpaceSolver after eliminating the bound multipliers):
 *   (useW*W + dreg) dx + Jc' dyc + Jd' dyd = -rx
 *   Dsig ds - dyd                          = -rs        (Dsig = Sigma_s + dreg, given)
 *   Jc dx                                  = -rc
 *   Jd dx - ds                             = -rd
 * x = (X_1..X_N, U_0..U_{N-1}), Jc rows c_{k+1} = X_{k+1} - F(X_k,U_k), Jd = selector of U.
 * Returns 0 when the inertia is (7N, 5N, 0), 1 otherwise.
 *
 * Riccati backend: eliminate ds = dU + rd, dyd = Dsig ds + rs  => Huu += Dsig, qu = rxU + Dsig rd + rs;
 * the remaining equality-constrained QP in (dX,dU) is solved by the backward/forward recursion and the
 * inertia condition is "every condensed Quu_k is positive definite".
 */
static int kkt_riccati(ws_t *w, const orc_params *p, int useW, double dreg, const double *rc, const double *rd,
                       step_t *o) {
    int N = w->N;
    double dt = p->dt;
    double P[3][3], pv[3];
    {
        const stage_t *s = &w->st[N];
        for (int i = 0; i < 3; i++) {
            for (int j = 0; j < 3; j++) P[i][j] = useW ? s->H[i][j] : 0.0;
            P[i][i] += dreg;
            pv[i] = w->rxX[3 * N + i];
        }
        memcpy(w->P + 9 * N, P, sizeof(P));
        memcpy(w->pv + 3 * N, pv, sizeof(pv));
    }
    for (int k = N - 1; k >= 0; k--) {
        const stage_t *s = &w->st[k];
        double A[3][3] = {{1, 0, s->a13}, {0, 1, s->a23}, {0, 0, 1}};
        double B[3][2] = {{s->b11, s->b12}, {s->b21, s->b22}, {0, dt}};
        double H[5][5];
        for (int i = 0; i < 5; i++) {
            for (int j = 0; j < 5; j++) H[i][j] = useW ? s->H[i][j] : 0.0;
            H[i][i] += dreg;
        }
        double q[5];
        for (int i = 0; i < 3; i++) q[i] = (k >= 1) ? w->rxX[3 * k + i] : 0.0;
        for (int i = 0; i < 2; i++) {
            H[3 + i][3 + i] += w->Dsig[2 * k + i];
            q[3 + i] = w->rxU[2 * k + i] + w->Dsig[2 * k + i] * rd[2 * k + i] + w->rs[2 * k + i];
        }
        double d[3], Pd[3];
        for (int i = 0; i < 3; i++) d[i] = -rc[3 * (k + 1) + i];
        for (int i = 0; i < 3; i++) {
            Pd[i] = pv[i];
            for (int j = 0; j < 3; j++) Pd[i] += P[i][j] * d[j];
        }
        double PA[3][3], PB[3][2];
        for (int i = 0; i < 3; i++) {
            for (int j = 0; j < 3; j++) {
                PA[i][j] = 0;
                for (int l = 0; l < 3; l++) PA[i][j] += P[i][l] * A[l][j];
            }
            for (int j = 0; j < 2; j++) {
                PB[i][j] = 0;
                for (int l = 0; l < 3; l++) PB[i][j] += P[i][l] * B[l][j];
            }
        }
        double Qxx[3][3], Qux[2][3], Quu[2][2], qx[3], qu[2];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                Qxx[i][j] = H[i][j];
                for (int l = 0; l < 3; l++) Qxx[i][j] += A[l][i] * PA[l][j];
            }
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 3; j++) {
                Qux[i][j] = H[3 + i][j];
                for (int l = 0; l < 3; l++) Qux[i][j] += B[l][i] * PA[l][j];
            }
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++) {
                Quu[i][j] = H[3 + i][3 + j];
                for (int l = 0; l < 3; l++) Quu[i][j] += B[l][i] * PB[l][j];
            }
        for (int i = 0; i < 3; i++) {
            qx[i] = q[i];
            for (int l = 0; l < 3; l++) qx[i] += A[l][i] * Pd[l];
        }
        for (int i = 0; i < 2; i++) {
            qu[i] = q[3 + i];
            for (int l = 0; l < 3; l++) qu[i] += B[l][i] * Pd[l];
        }
        double q01 = 0.5 * (Quu[0][1] + Quu[1][0]);
        double det = Quu[0][0] * Quu[1][1] - q01 * q01;
        if (!(Quu[0][0] > 0.0) || !(det > 0.0)) return 1;
        double i00 = Quu[1][1] / det, i01 = -q01 / det, i11 = Quu[0][0] / det;
        double *K = w->K + 6 * k, *kf = w->kf + 2 * k;
        for (int j = 0; j < 3; j++) {
            K[j] = -(i00 * Qux[0][j] + i01 * Qux[1][j]);
            K[3 + j] = -(i01 * Qux[0][j] + i11 * Qux[1][j]);
        }
        kf[0] = -(i00 * qu[0] + i01 * qu[1]);
        kf[1] = -(i01 * qu[0] + i11 * qu[1]);
        if (k >= 1) {
            double Pn[3][3];
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) Pn[i][j] = Qxx[i][j] + Qux[0][i] * K[j] + Qux[1][i] * K[3 + j];
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) P[i][j] = 0.5 * (Pn[i][j] + Pn[j][i]);
            for (int i = 0; i < 3; i++) pv[i] = qx[i] + Qux[0][i] * kf[0] + Qux[1][i] * kf[1];
            memcpy(w->P + 9 * k, P, sizeof(P));
            memcpy(w->pv + 3 * k, pv, sizeof(pv));
        }
    }
    /* forward */
    o->dX[0] = o->dX[1] = o->dX[2] = 0;
    for (int k = 0; k < N; k++) {
        const stage_t *s = &w->st[k];
        const double *K = w->K + 6 * k, *kf = w->kf + 2 * k;
        const double *dx = o->dX + 3 * k;
        double du0 = kf[0] + K[0] * dx[0] + K[1] * dx[1] + K[2] * dx[2];
        double du1 = kf[1] + K[3] * dx[0] + K[4] * dx[1] + K[5] * dx[2];
        o->dU[2 * k] = du0;
        o->dU[2 * k + 1] = du1;
        double *dn = o->dX + 3 * (k + 1);
        const double *c = rc + 3 * (k + 1);
        dn[0] = dx[0] + s->a13 * dx[2] + s->b11 * du0 + s->b12 * du1 - c[0];
        dn[1] = dx[1] + s->a23 * dx[2] + s->b21 * du0 + s->b22 * du1 - c[1];
        dn[2] = dx[2] + dt * du1 - c[2];
    }
    for (int k = 1; k <= N; k++) {
        const double *Pk = w->P + 9 * k, *pk = w->pv + 3 * k, *dx = o->dX + 3 * k;
        for (int i = 0; i < 3; i++)
            o->dlam[3 * k + i] = -(pk[i] + Pk[3 * i] * dx[0] + Pk[3 * i + 1] * dx[1] + Pk[3 * i + 2] * dx[2]);
    }
    for (int i = 0; i < 2 * N; i++) {
        o->dS[i] = o->dU[i] + rd[i];
        o->dyd[i] = w->Dsig[i] * o->dS[i] + w->rs[i];
    }
    return 0;
}

/* Dense backend: assemble the 12N x 12N augmented matrix literally and factor it (MUMPS stand-in). */
static int kkt_dense(ws_t *w, const orc_params *p, int useW, double dreg, const double *rc, const double *rd,
                     step_t *o) {
    int N = w->N, n = 12 * N;
    double dt = p->dt;
    double *A = w->Kd, *b = w->bd;
    memset(A, 0, sizeof(double) * (size_t)n * n);
#define IX(k, i) (3 * ((k)-1) + (i))         /* X_k, k=1..N */
#define IU(k, i) (3 * N + 2 * (k) + (i))     /* U_k, k=0..N-1 */
#define IS(k, i) (5 * N + 2 * (k) + (i))
#define IC(k, i) (7 * N + 3 * ((k)-1) + (i)) /* c_k, k=1..N */
#define ID(k, i) (10 * N + 2 * (k) + (i))
#define SET(r, c, v)                                 \
    do {                                             \
        int r_ = (r), c_ = (c);                      \
        if (r_ >= c_) A[(size_t)r_ * n + c_] += (v); \
    } while (0)
    for (int k = 0; k <= N; k++) {
        const stage_t *s = &w->st[k];
        int idx[5];
        for (int i = 0; i < 3; i++) idx[i] = (k >= 1) ? IX(k, i) : -1;
        for (int i = 0; i < 2; i++) idx[3 + i] = (k < N) ? IU(k, i) : -1;
        for (int i = 0; i < 5; i++)
            for (int j = 0; j < 5; j++) {
                if (idx[i] < 0 || idx[j] < 0) continue;
                double v = useW ? s->H[i][j] : 0.0;
                if (i == j) v += dreg;
                SET(idx[i], idx[j], v);
            }
    }
    for (int k = 0; k < N; k++) {
        const stage_t *s = &w->st[k];
        for (int i = 0; i < 2; i++) {
            SET(IS(k, i), IS(k, i), w->Dsig[2 * k + i]);
            SET(ID(k, i), IU(k, i), 1.0);
            SET(ID(k, i), IS(k, i), -1.0);
        }
        double Am[3][3] = {{1, 0, s->a13}, {0, 1, s->a23}, {0, 0, 1}};
        double Bm[3][2] = {{s->b11, s->b12}, {s->b21, s->b22}, {0, dt}};
        for (int i = 0; i < 3; i++) {
            SET(IC(k + 1, i), IX(k + 1, i), 1.0);
            if (k >= 1)
                for (int j = 0; j < 3; j++) SET(IC(k + 1, i), IX(k, j), -Am[i][j]);
            for (int j = 0; j < 2; j++) SET(IC(k + 1, i), IU(k, j), -Bm[i][j]);
        }
    }
    for (int k = 1; k <= N; k++)
        for (int i = 0; i < 3; i++) {
            b[IX(k, i)] = -w->rxX[3 * k + i];
            b[IC(k, i)] = -rc[3 * k + i];
        }
    for (int k = 0; k < N; k++)
        for (int i = 0; i < 2; i++) {
            b[IU(k, i)] = -w->rxU[2 * k + i];
            b[IS(k, i)] = -w->rs[2 * k + i];
            b[ID(k, i)] = -rd[2 * k + i];
        }
    int inertia[3];
    int info = bk_factor(n, A, w->ipiv);
    bk_inertia(n, A, w->ipiv, inertia);
    if (info || inertia[0] != 7 * N || inertia[1] != 5 * N) return 1;
    bk_solve(n, A, w->ipiv, b);
    o->dX[0] = o->dX[1] = o->dX[2] = 0;
    for (int k = 1; k <= N; k++)
        for (int i = 0; i < 3; i++) {
            o->dX[3 * k + i] = b[IX(k, i)];
            o->dlam[3 * k + i] = b[IC(k, i)];
        }
    for (int k = 0; k < N; k++)
        for (int i = 0; i < 2; i++) {
            o->dU[2 * k + i] = b[IU(k, i)];
            o->dS[2 * k + i] = b[IS(k, i)];
            o->dyd[2 * k + i] = b[ID(k, i)];
        }
#undef IX
#undef IU
#undef IS
#undef IC
#undef ID
#undef SET
    return 0;
}

static int kkt_solve(ws_t *w, const orc_params *p, int useW, double dreg, const double *rc, const double *rd,
                     step_t *o) {
    if (p->linear_solver == 1) return kkt_dense(w, p, useW, dreg, rc, rd, o);
    return kkt_riccati(w, p, useW, dreg, rc, rd, o);
}

void orc_default_options(orc_params *p) {
    p->tol = 1e-8;
    p->max_iter = 3000;
    p->acceptable_tol = 1e-6;
    p->acceptable_iter = 15;
    p->mu_init = 0.1;
    p->max_soc = 4;
    p->linear_solver = 0;
}

/* IPOPT constants (defaults) */
#define K_EPS 10.0 /* barrier_tol_factor */
#define K_MU 0.2   /* mu_linear_decrease_factor */
#define TH_MU 1.5  /* mu_superlinear_decrease_power */
#define TAU_MIN 0.99
#define S_MAX 100.0
#define GAMMA_THETA 1e-5
#define GAMMA_PHI 1e-8
#define ETA_PHI 1e-8
#define S_THETA 1.1
#define S_PHI 2.3
#define DELTA_LS 1.0
#define ALPHA_MIN_FRAC 0.05
#define KAPPA_SOC 0.99
#define KAPPA_SIGMA 1e10
#define BOUND_PUSH 0.01
#define BOUND_FRAC 0.01
#define BOUND_RELAX 1e-8
#define DW_INIT 1e-4
#define DW_MIN 1e-20
#define DW_MAX 1e20
#define DW_INC_FIRST 100.0
#define DW_INC 8.0
#define DW_DEC (1.0 / 3.0)
#define MAX_FILTER 32
#define OBJ_MAX_INC 5.0
#define TINY_STEP_TOL (10.0 * DBL_EPSILON)
#define TINY_STEP_Y_TOL 1e-2
#define KAPPA_RESTO 0.9
#define RESTO_T_MIN (1.0 / 1024.0)
#define MAX_RESTO 20

/* Filter: MAX_FILTER slots (phi, theta, valid).  A new entry evicts the entries it dominates and takes the
 * lowest free slot; when every slot is taken it overwrites slot (ring++ % MAX_FILTER). */
typedef struct {
    double phi[MAX_FILTER], theta[MAX_FILTER];
    int valid[MAX_FILTER];
    int ring;
} filter_t;

static int cmp_le(double lhs, double rhs, double bas) { return lhs - rhs <= 10.0 * DBL_EPSILON * fabs(bas); }

static void filter_reset(filter_t *f) {
    for (int i = 0; i < MAX_FILTER; i++) f->valid[i] = 0;
}

static int filter_acceptable(const filter_t *f, double phi, double theta) {
    for (int i = 0; i < MAX_FILTER; i++) {
        if (!f->valid[i]) continue;
        int ok = cmp_le(phi, f->phi[i], f->phi[i]) || cmp_le(theta, f->theta[i], f->theta[i]);
        if (!ok) return 0;
    }
    return 1;
}

static void filter_add(filter_t *f, double phi, double theta) {
    int slot = -1;
    for (int i = 0; i < MAX_FILTER; i++) {
        if (f->valid[i] && f->phi[i] >= phi && f->theta[i] >= theta) f->valid[i] = 0; /* dominated */
        if (!f->valid[i] && slot < 0) slot = i;
    }
    if (slot < 0) slot = (f->ring++) % MAX_FILTER;
    f->phi[slot] = phi;
    f->theta[slot] = theta;
    f->valid[slot] = 1;
}

/* line-search reference values of the current iterate */
typedef struct {
    double phi, theta, gbd, theta_min, theta_max;
    const filter_t *filt;
} ls_ref_t;

/* FilterLSAcceptor::CheckAcceptabilityOfTrialPoint; *ftype_armijo = 1 when the step is an f-type step that
 * satisfies the Armijo condition (then the filter is not augmented). */
static int ls_acceptable(const ls_ref_t *r, double alpha_test, double phi_t, double th_t, int *ftype_armijo) {
    *ftype_armijo = 0;
    if (!isfinite(th_t) || !isfinite(phi_t)) return 0;
    if (th_t > r->theta_max) return 0;
    int ftype = (r->gbd < 0) && (alpha_test * pow(-r->gbd, S_PHI) > DELTA_LS * pow(r->theta, S_THETA));
    int armijo = cmp_le(phi_t - r->phi, ETA_PHI * alpha_test * r->gbd, r->phi);
    *ftype_armijo = ftype && armijo;
    int ok;
    if (alpha_test > 0 && ftype && r->theta <= r->theta_min) {
        ok = armijo;
    } else {
        if (phi_t > r->phi) {
            double bas = 1.0;
            if (fabs(r->phi) > 10.0) bas = log10(fabs(r->phi));
            if (log10(phi_t - r->phi) > OBJ_MAX_INC + bas) return 0;
        }
        ok = cmp_le(th_t, (1 - GAMMA_THETA) * r->theta, r->theta) ||
             cmp_le(phi_t - r->phi, -GAMMA_PHI * r->theta, r->phi);
    }
    if (!ok) return 0;
    return filter_acceptable(r->filt, phi_t, th_t);
}

static double frac_to_bound(const ws_t *w, double tau, const double *dS) {
    double a = 1.0;
    for (int i = 0; i < 2 * w->N; i++) {
        double sl = w->S[i] - w->sL[i & 1], su = w->sU[i & 1] - w->S[i], ds = dS[i];
        if (ds < 0 && -tau * sl / ds < a) a = -tau * sl / ds;
        if (ds > 0 && tau * su / ds < a) a = tau * su / ds;
    }
    return a;
}

/* trial point curr + alpha*step; evaluates theta (1-norm) and the barrier objective there */
static void trial_eval(ws_t *w, const prob_t *q, const step_t *s, double alpha, double mu, double df, double *th_t,
                       double *phi_t) {
    int N = w->N, n2 = 2 * N;
    for (int i = 0; i < 3 * (N + 1); i++) w->Xt[i] = w->X[i] + alpha * s->dX[i];
    for (int i = 0; i < n2; i++) {
        w->Ut[i] = w->U[i] + alpha * s->dU[i];
        w->St[i] = w->S[i] + alpha * s->dS[i];
    }
    eval_c(q, w->Xt, w->Ut, w->ct);
    double th = 0, bar = 0;
    for (int i = 3; i < 3 * (N + 1); i++) th += fabs(w->ct[i]);
    for (int i = 0; i < n2; i++) {
        th += fabs(w->Ut[i] - w->St[i]);
        bar -= mu * (log(w->St[i] - w->sL[i & 1]) + log(w->sU[i & 1] - w->St[i]));
    }
    *th_t = th;
    *phi_t = df * eval_f(q, w->Xt, w->Ut) + bar;
}

static void rollout(const orc_params *p, const double *x0, const double *U, double *X) {
    X[0] = x0[0]; X[1] = x0[1]; X[2] = x0[2];
    for (int k = 0; k < p->N; k++) dyn_step(p, X + 3 * k, U + 2 * k, X + 3 * (k + 1));
}

int orc_solve(const orc_params *p, const double *x0, const double *xref, const double *uref,
              const double *obs_x, const double *obs_y, const double *u_init, const double *x_init,
              double *X_out, double *U_out, double *cost_out, orc_stats *stats) {
    const int N = p->N, n2 = 2 * N, n3 = 3 * (N + 1);
    prob_t q = {p, x0, xref, uref, obs_x, obs_y};
    ws_t W, *w = &W;
    ws_alloc(w, N, p->linear_solver == 1);
    orc_stats stt;
    memset(&stt, 0, sizeof(stt));
    int status = ORC_MAXITER_EXCEEDED;
    int iter = 0;
    const int trace = getenv("ORC_TRACE") != NULL; /* debugging aid: one line per iteration on stderr */

    /* relaxed slack bounds (bound_relax_factor) */
    for (int i = 0; i < 2; i++) {
        w->sL[i] = p->u_lo[i] - BOUND_RELAX * fmax(1.0, fabs(p->u_lo[i]));
        w->sU[i] = p->u_hi[i] + BOUND_RELAX * fmax(1.0, fabs(p->u_hi[i]));
    }
    /* starting point: Opti default X = 0 (X[:,0] is the parameter x0), U = the u0 argument
     * (local_planner_point_stabilization.py:78) */
    for (int k = 0; k <= N; k++)
        for (int i = 0; i < 3; i++) w->X[3 * k + i] = (k == 0) ? x0[i] : (x_init ? x_init[3 * k + i] : 0.0);
    for (int i = 0; i < n2; i++) w->U[i] = u_init ? u_init[i] : 0.0;
    /* slack initialisation: s = d(x) pushed into the interior (bound_push / bound_frac); bound multipliers 1 */
    for (int i = 0; i < n2; i++) {
        double lo = w->sL[i & 1], hi = w->sU[i & 1];
        double pl = fmin(BOUND_PUSH * fmax(1.0, fabs(lo)), BOUND_FRAC * (hi - lo));
        double pu = fmin(BOUND_PUSH * fmax(1.0, fabs(hi)), BOUND_FRAC * (hi - lo));
        double s = w->U[i];
        if (s < lo + pl) s = lo + pl;
        if (s > hi - pu) s = hi - pu;
        w->S[i] = s;
        w->vL[i] = 1.0;
        w->vU[i] = 1.0;
    }
    /* gradient-based objective scaling (nlp_scaling_max_gradient = 100) */
    double df = 1.0;
    {
        eval_stages(&q, w->X, w->U, w->lam, 1.0, w->st);
        double gmax = 0;
        int bad = 0;
        for (int k = 0; k <= N; k++) {
            for (int i = 0; i < 5; i++) {
                if ((k == 0 && i < 3) || (k == N && i >= 3)) continue;
                double g = w->st[k].g[i];
                if (!isfinite(g)) bad = 1;
                gmax = fmax(gmax, fabs(g));
            }
            if (k < N)
                for (int i = 0; i < 3; i++)
                    if (!isfinite(w->st[k].c[i])) bad = 1;
        }
        double f0 = eval_f(&q, w->X, w->U);
        if (bad || !isfinite(f0)) {
            status = ORC_INVALID_NUMBER_DETECTED;
            goto done;
        }
        if (gmax > 100.0) df = fmax(100.0 / gmax, 1e-8);
    }
    stt.obj_scale = df;

    /* least-squares multiplier estimate: min ||grad f + Jc' lam + Jd' yd||^2 + ||-yd - vL + vU||^2,
     * discarded when it exceeds constr_mult_init_max = 1e3 */
    {
        eval_stages(&q, w->X, w->U, w->lam, df, w->st);
        for (int k = 0; k <= N; k++)
            for (int i = 0; i < 3; i++) { w->rxX[3 * k + i] = w->st[k].g[i]; w->rc[3 * k + i] = 0; }
        for (int i = 0; i < n2; i++) {
            w->rxU[i] = w->st[i / 2].g[3 + (i & 1)];
            w->rs[i] = -w->vL[i] + w->vU[i];
            w->rd[i] = 0;
            w->Dsig[i] = 1.0;
        }
        int bad = kkt_solve(w, p, 0, 1.0, w->rc, w->rd, &w->main);
        double ymax = 0;
        for (int k = 1; k <= N; k++)
            for (int i = 0; i < 3; i++) ymax = fmax(ymax, fabs(w->main.dlam[3 * k + i]));
        for (int i = 0; i < n2; i++) ymax = fmax(ymax, fabs(w->main.dyd[i]));
        if (bad || !(ymax <= 1e3)) {
            memset(w->lam, 0, sizeof(double) * n3);
            memset(w->yd, 0, sizeof(double) * n2);
        } else {
            memcpy(w->lam, w->main.dlam, sizeof(double) * n3);
            memcpy(w->yd, w->main.dyd, sizeof(double) * n2);
        }
    }

    double mu = p->mu_init, tau = fmax(TAU_MIN, 1.0 - mu);
    const double mu_floor = fmin(p->tol, 1e-4) / (K_EPS + 1.0);
    filter_t filt;
    filt.ring = 0;
    filter_reset(&filt);
    double theta_max = -1, theta_min = -1;
    double dw_last = 0.0;
    int acceptable_count = 0;
    int tiny_last = 0, tiny_flag = 0; /* BacktrackingLineSearch::tiny_step_last_iteration_, IpoptData::tiny_step_flag */

    for (;;) {
        /* ---- evaluate the current point ---- */
        eval_stages(&q, w->X, w->U, w->lam, df, w->st);
        double theta = 0, prim_inf = 0;
        for (int k = 0; k < N; k++)
            for (int i = 0; i < 3; i++) {
                double c = w->st[k].c[i];
                w->rc[3 * (k + 1) + i] = c;
                theta += fabs(c);
                prim_inf = fmax(prim_inf, fabs(c));
            }
        for (int i = 0; i < n2; i++) {
            w->rd[i] = w->U[i] - w->S[i];
            theta += fabs(w->rd[i]);
            prim_inf = fmax(prim_inf, fabs(w->rd[i]));
        }
        if (theta_max < 0) {
            theta_max = 1e4 * fmax(1.0, theta);
            theta_min = 1e-4 * fmax(1.0, theta);
        }
        /* grad_lag_x, grad_lag_s */
        double dual_inf = 0, sum_y = 0, sum_z = 0;
        for (int k = 1; k <= N; k++) {
            const double *l = w->lam + 3 * k;
            double r0 = w->st[k].g[0] + l[0], r1 = w->st[k].g[1] + l[1], r2 = w->st[k].g[2] + l[2];
            if (k < N) {
                const double *ln = w->lam + 3 * (k + 1);
                r0 -= ln[0];
                r1 -= ln[1];
                r2 -= w->st[k].a13 * ln[0] + w->st[k].a23 * ln[1] + ln[2];
            }
            w->rxX[3 * k] = r0; w->rxX[3 * k + 1] = r1; w->rxX[3 * k + 2] = r2;
            dual_inf = fmax(dual_inf, fmax(fabs(r0), fmax(fabs(r1), fabs(r2))));
            sum_y += fabs(l[0]) + fabs(l[1]) + fabs(l[2]);
        }
        for (int k = 0; k < N; k++) {
            const stage_t *s = &w->st[k];
            const double *ln = w->lam + 3 * (k + 1);
            double r0 = s->g[3] - (s->b11 * ln[0] + s->b21 * ln[1]) + w->yd[2 * k];
            double r1 = s->g[4] - (s->b12 * ln[0] + s->b22 * ln[1] + p->dt * ln[2]) + w->yd[2 * k + 1];
            w->rxU[2 * k] = r0; w->rxU[2 * k + 1] = r1;
            dual_inf = fmax(dual_inf, fmax(fabs(r0), fabs(r1)));
        }
        double compl0 = 0;
        for (int i = 0; i < n2; i++) {
            double gs = -w->yd[i] - w->vL[i] + w->vU[i];
            dual_inf = fmax(dual_inf, fabs(gs));
            sum_y += fabs(w->yd[i]);
            sum_z += fabs(w->vL[i]) + fabs(w->vU[i]);
            double sl = w->S[i] - w->sL[i & 1], su = w->sU[i & 1] - w->S[i];
            compl0 = fmax(compl0, fmax(fabs(sl * w->vL[i]), fabs(su * w->vU[i])));
        }
        double sd = fmax(S_MAX, (sum_y + sum_z) / (double)(5 * N + 4 * N)) / S_MAX;
        double sc = fmax(S_MAX, sum_z / (double)(4 * N)) / S_MAX;
        double E0 = fmax(dual_inf / sd, fmax(prim_inf, compl0 / sc));
        stt.err = E0;
        stt.mu = mu;
        if (trace)
            fprintf(stderr, "it %4d f %.9e th %.3e du %.3e cm %.3e E0 %.3e mu %.1e dw_last %.1e resto %d lsx %d\n", iter,
                    eval_f(&q, w->X, w->U), theta, dual_inf, compl0, E0, mu, dw_last, stt.n_resto, stt.ls_extra),
            fprintf(stderr, "        df %.3e acc_count %d prim %.3e dual/df %.3e compl/df %.3e\n", df, acceptable_count, prim_inf, dual_inf / df, compl0 / df);
        if (!isfinite(E0)) { status = ORC_INVALID_NUMBER_DETECTED; break; }

        /* ---- convergence (OptimalityErrorConvergenceCheck) ---- */
        if (E0 <= p->tol && dual_inf / df <= 1.0 && prim_inf <= 1e-4 && compl0 / df <= 1e-4) {
            status = ORC_SOLVE_SUCCEEDED;
            break;
        }
        if (p->acceptable_iter > 0 && E0 <= p->acceptable_tol && dual_inf / df <= 1e10 && prim_inf <= 1e-2 &&
            compl0 / df <= 1e-2) {
            acceptable_count++;
            if (acceptable_count >= p->acceptable_iter) { status = ORC_SOLVED_TO_ACCEPTABLE_LEVEL; break; }
        } else {
            acceptable_count = 0;
        }
        if (iter >= p->max_iter) { status = ORC_MAXITER_EXCEEDED; break; }

        /* ---- barrier parameter update (MonotoneMuUpdate::UpdateBarrierParameter, fast decrease allowed).
         * A tiny step in two consecutive iterations forces a decrease of mu; when mu cannot decrease any more the
         * problem is "solved to best possible numerical accuracy": Search_Direction_Becomes_Too_Small. ---- */
        {
            int tflag = tiny_flag, stop = 0;
            tiny_flag = 0;
            for (;;) {
                double cm = 0;
                for (int i = 0; i < n2; i++) {
                    double sl = w->S[i] - w->sL[i & 1], su = w->sU[i & 1] - w->S[i];
                    cm = fmax(cm, fmax(fabs(sl * w->vL[i] - mu), fabs(su * w->vU[i] - mu)));
                }
                double Emu = fmax(dual_inf / sd, fmax(prim_inf, cm / sc));
                if (!(Emu <= K_EPS * mu) && !tflag) break;
                double nm = fmax(fmin(K_MU * mu, pow(mu, TH_MU)), mu_floor);
                if (nm == mu) { stop = tflag; break; }
                mu = nm;
                tau = fmax(TAU_MIN, 1.0 - mu);
                filter_reset(&filt);
                tflag = 0;
            }
            if (stop) { status = ORC_SEARCH_DIRECTION_TOO_SMALL; break; }
        }

        /* ---- search direction with inertia correction (PDPerturbationHandler, delta_x = delta_s) ---- */
        for (int i = 0; i < n2; i++) {
            double sl = w->S[i] - w->sL[i & 1], su = w->sU[i & 1] - w->S[i];
            w->rs[i] = -w->yd[i] - mu / sl + mu / su;
        }
        double dw = 0.0;
        int solved = 0;
        for (;;) {
            for (int i = 0; i < n2; i++) {
                double sl = w->S[i] - w->sL[i & 1], su = w->sU[i & 1] - w->S[i];
                w->Dsig[i] = w->vL[i] / sl + w->vU[i] / su + dw;
            }
            if (kkt_solve(w, p, 1, dw, w->rc, w->rd, &w->main) == 0) { solved = 1; break; }
            stt.n_reg++;
            if (dw == 0.0) dw = (dw_last == 0.0) ? DW_INIT : fmax(DW_MIN, dw_last * DW_DEC);
            else dw = (dw_last == 0.0 || 1e5 * dw_last < dw) ? dw * DW_INC_FIRST : dw * DW_INC;
            if (dw > DW_MAX) break;
        }
        if (!solved) { status = ORC_ERROR_IN_STEP_COMPUTATION; break; }
        if (dw > 0.0) dw_last = dw;
        {
            int bad = 0;
            for (int i = 0; i < n3; i++)
                if (!isfinite(w->main.dX[i]) || !isfinite(w->main.dlam[i])) bad = 1;
            for (int i = 0; i < n2; i++)
                if (!isfinite(w->main.dU[i]) || !isfinite(w->main.dS[i]) || !isfinite(w->main.dyd[i])) bad = 1;
            if (bad) { status = ORC_ERROR_IN_STEP_COMPUTATION; break; }
        }

        /* ---- filter line search (BacktrackingLineSearch + FilterLSAcceptor) ---- */
        double a_max = frac_to_bound(w, tau, w->main.dS);
        double f_cur = eval_f(&q, w->X, w->U);
        double bar = 0, gbd = 0;
        for (int i = 0; i < n2; i++) {
            double sl = w->S[i] - w->sL[i & 1], su = w->sU[i & 1] - w->S[i];
            bar -= mu * (log(sl) + log(su));
            gbd += (-mu / sl + mu / su) * w->main.dS[i];
        }
        for (int k = 1; k <= N; k++)
            for (int i = 0; i < 3; i++) gbd += w->st[k].g[i] * w->main.dX[3 * k + i];
        for (int k = 0; k < N; k++)
            for (int i = 0; i < 2; i++) gbd += w->st[k].g[3 + i] * w->main.dU[2 * k + i];
        ls_ref_t ref = {df * f_cur + bar, theta, gbd, theta_min, theta_max, &filt};

        double a_min = GAMMA_THETA;
        if (gbd < 0) {
            a_min = fmin(GAMMA_THETA, GAMMA_PHI * theta / (-gbd));
            if (theta <= theta_min) a_min = fmin(a_min, DELTA_LS * pow(theta, S_THETA) / pow(-gbd, S_PHI));
        }
        a_min *= ALPHA_MIN_FRAC;

        double alpha = a_max, alpha_acc = 0;
        const step_t *acc = NULL;
        int fa = 0, ntrial = 0;
        /* BacktrackingLineSearch::DetectTinyStep: every primal component moves by less than tiny_step_tol = 10 eps
         * (relative) and the point is nearly feasible -> the full step is taken without a line search. */
        int tiny = 0;
        {
            /* |d_i| / (1 + |x_i|) <= tol, written without the division (the CUDA kernels use the same form) */
            int big = 0;
            double c2 = 0;
            for (int i = 3; i < n3; i++)
                if (!(fabs(w->main.dX[i]) <= TINY_STEP_TOL * (1.0 + fabs(w->X[i])))) big = 1;
            for (int i = 0; i < n2; i++) {
                if (!(fabs(w->main.dU[i]) <= TINY_STEP_TOL * (1.0 + fabs(w->U[i])))) big = 1;
                if (!(fabs(w->main.dS[i]) <= TINY_STEP_TOL * (1.0 + fabs(w->S[i])))) big = 1;
                c2 += w->rd[i] * w->rd[i];
            }
            for (int i = 3; i < n3; i++) c2 += w->rc[i] * w->rc[i];
            tiny = !big && (sqrt(c2) <= 1e-4);
        }
        if (tiny) {
            double th_t, phi_t;
            trial_eval(w, &q, &w->main, alpha, mu, df, &th_t, &phi_t);
            if (isfinite(th_t) && isfinite(phi_t)) {
                (void)ls_acceptable(&ref, alpha, phi_t, th_t, &fa); /* only for the filter-augmentation rule */
                acc = &w->main;
                alpha_acc = alpha;
                if (tiny_last) tiny_flag = 1;
                double dy = 0;
                for (int i = 3; i < n3; i++) dy = fmax(dy, fabs(w->main.dlam[i]));
                for (int i = 0; i < n2; i++) dy = fmax(dy, fabs(w->main.dyd[i]));
                tiny_last = dy < TINY_STEP_Y_TOL;
            } else {
                tiny = 0;
            }
        }
        if (!tiny) { tiny_flag = 0; tiny_last = 0; }
        while (!acc) {
            double th_t, phi_t;
            trial_eval(w, &q, &w->main, alpha, mu, df, &th_t, &phi_t);
            if (ntrial++ > 0) stt.ls_extra++;
            if (ls_acceptable(&ref, alpha, phi_t, th_t, &fa)) { acc = &w->main; alpha_acc = alpha; break; }
            /* second-order correction: only after the first trial step, when the infeasibility did not drop */
            if (ntrial == 1 && p->max_soc > 0 && isfinite(th_t) && th_t >= theta) {
                memcpy(w->csoc, w->rc, sizeof(double) * n3);
                memcpy(w->dsoc, w->rd, sizeof(double) * n2);
                double alpha_soc = alpha, theta_soc_old = 0, th_trial = th_t;
                int count = 0;
                while (count < p->max_soc && !acc && (count == 0 || th_trial <= KAPPA_SOC * theta_soc_old)) {
                    theta_soc_old = th_trial;
                    for (int i = 3; i < n3; i++) w->csoc[i] = alpha_soc * w->csoc[i] + w->ct[i];
                    for (int i = 0; i < n2; i++) w->dsoc[i] = alpha_soc * w->dsoc[i] + (w->Ut[i] - w->St[i]);
                    stt.n_soc++;
                    if (kkt_solve(w, p, 1, dw, w->csoc, w->dsoc, &w->soc) != 0) break;
                    alpha_soc = frac_to_bound(w, tau, w->soc.dS);
                    double phi_s;
                    trial_eval(w, &q, &w->soc, alpha_soc, mu, df, &th_trial, &phi_s);
                    stt.ls_extra++;
                    if (ls_acceptable(&ref, alpha, phi_s, th_trial, &fa)) { acc = &w->soc; alpha_acc = alpha_soc; }
                    else count++;
                }
                if (acc) break;
            }
            alpha *= 0.5;
            if (alpha < a_min) break;
        }

        if (!acc) {
            /* Restoration stand-in.  IPOPT would switch to its feasibility-restoration NLP here.  For a
             * multiple-shooting transcription feasible points are available in closed form (roll the
             * controls out), so the stand-in augments the filter with the current point and moves to a
             * point that IPOPT's restoration acceptance test admits (theta <= kappa_resto * theta,
             * acceptable to the filter), then restarts the multipliers:
             *   stage 1  along the direction to "U := S, X rolled out" (S fixed), longest step
             *            t = 1, 1/2, ... 1/1024 that passes (incl. the obj_max_inc test);
             *   stage 2  (only if stage 1 found nothing: the roll-out runs into an obstacle point, where
             *            exp(c/s) overflows) the family "plan shrunk towards standing still":
             *            S_l = u_c + l (S - u_c), U := S_l, X rolled out, l = 1/2, 1/4, ... 1/1024, 0
             *            (u_c = the zero control pushed into the interior of its box); the member with the
             *            lowest barrier objective is taken.  l = 0 is the robot standing at x0, which is
             *            finite whenever the starting point was, so stage 2 always ends with a point. */
            if (theta <= 1e-10 || stt.n_resto >= MAX_RESTO) { status = ORC_RESTORATION_FAILED; break; }
            filter_add(&filt, ref.phi - GAMMA_PHI * theta, (1 - GAMMA_THETA) * theta);
            /* restoration direction: towards the closed-form feasible point (U = S, X rolled out) */
            for (int i = 0; i < n2; i++) { w->soc.dU[i] = w->S[i] - w->U[i]; w->soc.dS[i] = 0.0; }
            rollout(p, x0, w->S, w->Xt);
            for (int i = 0; i < n3; i++) w->soc.dX[i] = w->Xt[i] - w->X[i];
            double t_acc = 0.0;
            for (double t = 1.0; t >= RESTO_T_MIN; t *= 0.5) {
                double th_r, phi_r;
                trial_eval(w, &q, &w->soc, t, mu, df, &th_r, &phi_r);
                if (trace) fprintf(stderr, "   resto t %.4g th %.4e (ref %.4e) phi %.6e (ref %.6e) filt %d\n", t, th_r, theta, phi_r, ref.phi, filter_acceptable(&filt, phi_r, th_r));
                if (!isfinite(th_r) || !isfinite(phi_r)) continue;
                if (!(th_r <= KAPPA_RESTO * theta)) continue;
                if (phi_r > ref.phi) {
                    double bas = 1.0;
                    if (fabs(ref.phi) > 10.0) bas = log10(fabs(ref.phi));
                    if (log10(phi_r - ref.phi) > OBJ_MAX_INC + bas) continue;
                }
                if (!filter_acceptable(&filt, phi_r, th_r)) continue;
                t_acc = t;
                break;
            }
            if (t_acc == 0.0) {
                /* stage 2: the family of feasible points "the plan shrunk towards standing still": S_l = uc + l (S - uc),
                 * U = S_l, X rolled out, l = 1/2, 1/4, ..., 0; the member with the lowest barrier objective is taken */
                double uc[2];
                for (int i = 0; i < 2; i++) {
                    double lo = w->sL[i], hi = w->sU[i];
                    double pl = fmin(BOUND_PUSH * fmax(1.0, fabs(lo)), BOUND_FRAC * (hi - lo));
                    double pu = fmin(BOUND_PUSH * fmax(1.0, fabs(hi)), BOUND_FRAC * (hi - lo));
                    uc[i] = fmin(fmax(0.0, lo + pl), hi - pu);
                }
                double lam = 0.5, best = INFINITY, lam_best = -1.0;
                for (int pass = 0; pass < 2; pass++) {
                    for (;;) {
                        if (pass == 1) lam = lam_best;
                        for (int i = 0; i < n2; i++) w->St[i] = uc[i & 1] + lam * (w->S[i] - uc[i & 1]);
                        rollout(p, x0, w->St, w->Xt);
                        for (int i = 0; i < n2; i++) { w->soc.dU[i] = w->St[i] - w->U[i]; w->soc.dS[i] = w->St[i] - w->S[i]; }
                        for (int i = 0; i < n3; i++) w->soc.dX[i] = w->Xt[i] - w->X[i];
                        if (pass == 1) break;
                        double th_r, phi_r;
                        trial_eval(w, &q, &w->soc, 1.0, mu, df, &th_r, &phi_r);
                        if (trace) fprintf(stderr, "   resto2 lam %.4g th %.4e phi %.6e filt %d\n", lam, th_r, phi_r, filter_acceptable(&filt, phi_r, th_r));
                        if (isfinite(th_r) && isfinite(phi_r) && th_r <= KAPPA_RESTO * theta && filter_acceptable(&filt, phi_r, th_r) && phi_r < best) {
                            best = phi_r; lam_best = lam;
                        }
                        if (lam == 0.0) break;
                        lam = (lam * 0.5 >= RESTO_T_MIN) ? lam * 0.5 : 0.0;
                    }
                    if (lam_best < 0.0) break;
                    if (pass == 1) t_acc = 1.0;
                }
            }
            if (t_acc == 0.0) { status = ORC_RESTORATION_FAILED; break; }
            for (int i = 3; i < n3; i++) w->X[i] += t_acc * w->soc.dX[i];
            for (int i = 0; i < n2; i++) { w->U[i] += t_acc * w->soc.dU[i]; w->S[i] += t_acc * w->soc.dS[i]; }
            double zmax = 0;
            for (int i = 0; i < n2; i++) zmax = fmax(zmax, fmax(w->vL[i], w->vU[i]));
            if (zmax > 1e3)
                for (int i = 0; i < n2; i++) { w->vL[i] = 1.0; w->vU[i] = 1.0; }
            memset(w->lam, 0, sizeof(double) * n3);
            memset(w->yd, 0, sizeof(double) * n2);
            stt.n_resto++;
            iter++;
            continue;
        }

        /* ---- accept the trial point (IpoptAlgorithm::AcceptTrialPoint) ---- */
        if (!fa) filter_add(&filt, ref.phi - GAMMA_PHI * theta, (1 - GAMMA_THETA) * theta);
        double a_z = 1.0;
        for (int i = 0; i < n2; i++) {
            double sl = w->S[i] - w->sL[i & 1], su = w->sU[i & 1] - w->S[i];
            w->dvL[i] = mu / sl - w->vL[i] - w->vL[i] / sl * acc->dS[i];
            w->dvU[i] = mu / su - w->vU[i] + w->vU[i] / su * acc->dS[i];
            if (w->dvL[i] < 0 && -tau * w->vL[i] / w->dvL[i] < a_z) a_z = -tau * w->vL[i] / w->dvL[i];
            if (w->dvU[i] < 0 && -tau * w->vU[i] / w->dvU[i] < a_z) a_z = -tau * w->vU[i] / w->dvU[i];
        }
        for (int k = 1; k <= N; k++)
            for (int i = 0; i < 3; i++) {
                w->X[3 * k + i] += alpha_acc * acc->dX[3 * k + i];
                w->lam[3 * k + i] += alpha_acc * acc->dlam[3 * k + i];
            }
        for (int i = 0; i < n2; i++) {
            w->U[i] += alpha_acc * acc->dU[i];
            w->S[i] += alpha_acc * acc->dS[i];
            w->yd[i] += alpha_acc * acc->dyd[i];
            w->vL[i] += a_z * w->dvL[i];
            w->vU[i] += a_z * w->dvU[i];
            /* kappa_sigma safeguard on the bound multipliers */
            double sl = w->S[i] - w->sL[i & 1], su = w->sU[i & 1] - w->S[i];
            w->vL[i] = fmax(fmin(w->vL[i], KAPPA_SIGMA * mu / sl), mu / (KAPPA_SIGMA * sl));
            w->vU[i] = fmax(fmin(w->vU[i], KAPPA_SIGMA * mu / su), mu / (KAPPA_SIGMA * su));
        }
        iter++;
    }

done:
    stt.iters = iter;
    memcpy(X_out, w->X, sizeof(double) * n3);
    memcpy(U_out, w->U, sizeof(double) * n2);
    *cost_out = eval_f(&q, w->X, w->U);
    if (stats) *stats = stt;
    ws_free(w);
    return status;
}

/* ------------------------------------------------------------------------------------------ */
/* batch driver: one problem per task, nthreads host threads                                   */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    const orc_params *p;
    int B, obs_stride;
    const double *x0, *xref, *uref, *ox, *oy, *u_init;
    double *X, *U, *cost;
    int *status, *iters, *ls;
    int next;
    pthread_mutex_t mtx;
} batch_t;

static void *batch_worker(void *arg) {
    batch_t *b = (batch_t *)arg;
    const orc_params *p = b->p;
    int N = p->N;
    int nref = (p->ref_kind == ORC_REF_GOAL) ? 3 : 3 * N;
    for (;;) {
        pthread_mutex_lock(&b->mtx);
        int lo = b->next;
        b->next += 8;
        pthread_mutex_unlock(&b->mtx);
        if (lo >= b->B) break;
        int hi = lo + 8 < b->B ? lo + 8 : b->B;
        for (int i = lo; i < hi; i++) {
            orc_stats st;
            int s = orc_solve(p, b->x0 + 3 * (size_t)i, b->xref + (size_t)nref * i,
                              b->uref ? b->uref + 2 * (size_t)N * i : NULL,
                              b->ox ? b->ox + (size_t)b->obs_stride * i : NULL,
                              b->oy ? b->oy + (size_t)b->obs_stride * i : NULL,
                              b->u_init ? b->u_init + 2 * (size_t)N * i : NULL, NULL,
                              b->X + 3 * (size_t)(N + 1) * i, b->U + 2 * (size_t)N * i, b->cost + i, &st);
            b->status[i] = s;
            if (b->iters) b->iters[i] = st.iters;
            if (b->ls) b->ls[i] = st.ls_extra;
        }
    }
    return NULL;
}

int orc_solve_batch(const orc_params *p, int B, const double *x0, const double *xref, const double *uref,
                    const double *obs_x, const double *obs_y, int obs_stride, const double *u_init,
                    double *X_out, double *U_out, double *cost_out, int *status_out, int *iters_out,
                    int *ls_out, int nthreads) {
    batch_t b = {p, B, obs_stride, x0, xref, uref, obs_x, obs_y, u_init, X_out, U_out, cost_out,
                 status_out, iters_out, ls_out, 0, PTHREAD_MUTEX_INITIALIZER};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, batch_worker, &b);
    batch_worker(&b);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}
