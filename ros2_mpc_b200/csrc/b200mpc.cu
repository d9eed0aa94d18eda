// b200mpc.cu — batched nonlinear-MPC solve for B200 (sm_100a), FP64, one warp per problem.
//
// Replaces the CasADi/IPOPT call stack behind ros2_mpc's Mpc.perform_mpc
// (ros2_mpc/planner/local_planner_point_stabilization.py:69-87 and its two sibling variants):
//   K1 multiple-shooting transcription (rk4 :136-148 / euler_integration local_planner_tracking.py:132-137)
//   K2 unicycle RK4 step with analytic first and second derivatives (get_system_function :159-178)
//   K3 obstacle-cost value / gradient / Hessian over the obstacle list staged in shared memory
//      (define_obstacles_cost_function mpc_point_stabilization.py:46-53, local_planner_point_stabilization.py:60-67)
//   K4 primal-dual interior-point Newton iteration (what opti.solve() :84 runs inside IPOPT): stage-wise KKT
//      assembly, Riccati factorisation, fraction-to-the-boundary rule, filter line search with second-order
//      correction, inertia correction, monotone barrier update.
//
// Mapping: stage k of the horizon lives in lane k / J, slot k % J (J = ceil((N+1)/32) stages per lane), all
// per-stage quantities in registers.  Stage-parallel work (dynamics, cost, obstacle sums, residuals, trial
// points) runs on all lanes; norms and merit values are butterfly-shuffle reductions (bitwise identical on
// every lane, so all control flow is warp-uniform); the Riccati backward/forward recursions walk the lanes
// serially, broadcasting the 3x3 cost-to-go with shuffles.  Warps pull problems from a global counter, so a
// slow problem only delays its own warp.  No tensor cores: the blocks are 3x3 / 3x2.
//
// The algorithm and every constant are the same as in oracle/mpc_oracle.c (the CPU checker); the two share no code.
#include "b200mpc.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#define FULL 0xffffffffu
#ifndef B200MPC_MIN_CTAS
#define B200MPC_MIN_CTAS 2
#endif
#ifndef B200MPC_MIN_CTAS_OBS
#define B200MPC_MIN_CTAS_OBS 2 /* the obstacle instance of the warp kernel */
#endif
#ifndef B200MPC_TPP_MIN_CTAS
#define B200MPC_TPP_MIN_CTAS 1
#endif
#ifndef B200MPC_LANE_FUSED_DEFAULT
#define B200MPC_LANE_FUSED_DEFAULT 0
#endif

// ---- IPOPT defaults (Waechter & Biegler 2006; IPOPT option documentation) -------------------------------
#define K_EPS 10.0
#define K_MU 0.2
#define TH_MU 1.5
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
#define OBJ_MAX_INC 5.0
#define MAX_RESTO 20
#define KAPPA_RESTO 0.9
#define RESTO_T_MIN (1.0 / 1024.0)
#define TINY_STEP_TOL (10.0 * 2.220446049250313e-16)
#define TINY_STEP_Y_TOL 1e-2
#define DBL_EPS 2.220446049250313e-16

struct KParams {
    int N, M, integrator, ref_kind, obs_form, obs_k0, obs_k1, max_iter, acceptable_iter, max_soc;
    double dt, Q[3], R[2], kappa, obs_c, obs_r, u_lo[2], u_hi[2], tol, acceptable_tol, mu_init;
    double sL[2], sU[2], inv_r2, mu_floor;
    int kkt_scan; // warp kernel, N + 1 <= 32: parallel-in-time Riccati recursion (kkt_backward_scan)
};

struct BatchArgs {
    int B, obs_stride;
    const double *x0, *xref, *uref, *ox, *oy, *u_init;
    double *X, *U, *cost;
    int *status, *iters, *ls;
    unsigned int *counter;
    // Straggler hand-over (lane kernel -> warp kernel).  A lane-kernel launch lasts as long as its slowest problem, and a
    // trip of a nearly empty warp costs what a trip of a full one does, so the lane kernel EXPORTS the complete solver
    // state of a problem (record below) when the problem has taken hand_iter iterations, or when the work queue is empty
    // and its warp has thinned out; a second launch of the warp kernel (resume = the same records) picks those problems
    // up exactly where they stood — per-iteration latency 20 us instead of a 0.3-0.5 ms trip.
    double *hand_rec = nullptr;          // [hand_cap][HAND_REC(N)] records, or NULL (no hand-over)
    unsigned int *hand_count = nullptr;  // records written (lane kernel) / to read (warp kernel); may exceed hand_cap: clamp
    int hand_cap = 0, hand_iter = 0, hand_thin = 0;
    int hand_iter_tail = 0;              // threshold of the problems of the last wave (index >= B - lanes of the launch); 0: hand_iter
    // warp kernel, resume != 0: the batch IS the list of records resume_rec[min(*resume_count, resume_cap)] (B is ignored).
    // The warp kernel exports as well (hand_rec != NULL: problems reaching hand_iter iterations): a batch that does not fill
    // the machine several times over is solved as a cascade of iteration-bounded launches, so that its long problems run
    // side by side in the last launches instead of one behind the other at the end of a single one.
    int resume = 0, resume_cap = 0;
    const double *resume_rec = nullptr;
    const unsigned int *resume_count = nullptr;
};
// record of one exported problem (doubles): per stage the eight iterate rows of the lane kernel's workspace
// (X0 X1 | X2 lam0 | lam1 lam2 | U | S | yd | vL | vU), 16 scalars, the filter (32 x (phi, theta))
#define HAND_SCAL(N) (16 * ((N) + 1))
#define HAND_FILT(N) (HAND_SCAL(N) + 16)
#define HAND_REC(N) (HAND_FILT(N) + 64)
enum { HS_B = 0, HS_MU, HS_DF, HS_THETA0, HS_DWLAST, HS_ITER, HS_LS, HS_NRESTO, HS_ACCEPT, HS_TINYLAST, HS_TINYFLAG, HS_FMASK, HS_RING };

struct EvalArgs {
    int B, obs_stride;
    const double *x0, *xref, *uref, *ox, *oy, *X, *U, *lam;
    double obj_scale;
    double *f, *c, *grad, *stage;
};

// ---- warp reductions (butterfly: every lane ends with the same bits) ------------------------------------
#ifndef WRED_INLINE
#define WRED_INLINE __forceinline__
#endif
__device__ WRED_INLINE double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ WRED_INLINE double wmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ WRED_INLINE double wmin(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ int wor(int v) { return __any_sync(FULL, v); }

__device__ __forceinline__ bool cmp_le(double lhs, double rhs, double bas) {
    return lhs - rhs <= 10.0 * DBL_EPS * fabs(bas);
}

// sin and cos of one angle: three-constant Cody-Waite reduction by pi/2 (exact products through FMA) and the
// fdlibm kernel polynomials on [-pi/4, pi/4]; < 1.5 ulp for |x| <= 1e5 (headings are a few radians), the library
// routine beyond.  Half the instructions of sincos(), whose argument reduction for huge arguments the solver never needs.
__device__ __forceinline__ void tpp_sincos_core(double x, double &sn, double &cs) {
    const double kf = rint(x * 6.36619772367581382433e-01);
    double r = fma(-kf, 1.5707963267948966e+00, x);
    r = fma(-kf, 6.1232339957367574e-17, r);
    r = fma(-kf, 8.4784276603688985e-32, r);
    const double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double s = fma(z * r, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double c = fma(z * z, pc, fma(-0.5, z, 1.0));
    const int q = (int)kf;
    const double a = (q & 1) ? c : s, b = (q & 1) ? s : c;
    sn = (q & 2) ? -a : a;
    cs = ((q + 1) & 2) ? -b : b;
}

// sin / cos of th, th + hw, th + 2 hw (the RK4 stage angles) from two branch-free evaluations and the angle-addition
// formulas; the three library sincos() calls they replace each hide a slow-path branch, which keeps the compiler from
// interleaving their (independent) polynomial chains.  Library fallback for absurd arguments.
// Library routines the solver reaches only off its hot path (huge angles, the filter's power laws) live out of line, one
// copy each: inlined at every call site they made up a third of the kernel's code, which has to stream through a 32 KB
// instruction cache (profiles/r2_warp_a_*: 1.9 cycles of instruction-fetch stall per issued instruction).
__device__ __noinline__ void rk4_trig_far(double th, double hw, double *o) {
    sincos(th, o + 0, o + 1);
    sincos(th + hw, o + 2, o + 3);
    sincos(th + 2.0 * hw, o + 4, o + 5);
}
__device__ __noinline__ double pow_ool(double a, double b) { return pow(a, b); }
__device__ __noinline__ double log10_ool(double a) { return log10(a); }
__device__ __forceinline__ void rk4_trig(double th, double hw, double &s0, double &c0, double &sm, double &cm, double &se,
                                         double &ce) {
    // (per-lane test, no warp vote: the callers sit inside per-stage branches that not every lane takes)
    if (!(fabs(th) <= 1e5) || !(fabs(hw) <= 1e5)) {
        double o[6];
        rk4_trig_far(th, hw, o);
        s0 = o[0]; c0 = o[1]; sm = o[2]; cm = o[3]; se = o[4]; ce = o[5];
        return;
    }
    double sh, ch;
    tpp_sincos_core(th, s0, c0);
    tpp_sincos_core(hw, sh, ch);
    sm = s0 * ch + c0 * sh;
    cm = c0 * ch - s0 * sh;
    const double s2 = 2.0 * sh * ch, c2 = 1.0 - 2.0 * sh * sh;
    se = s0 * c2 + c0 * s2;
    ce = c0 * c2 - s0 * s2;
}

// Reciprocal to within 1 ulp: hardware seed (MUFU.RCP64H) + two Newton steps, five dependent instructions instead of
// the ~25 of an IEEE division — it sits on the serial critical path of the Riccati recursion.  det is finite, normal
// and > 0 whenever the result is used (otherwise the factorisation is rejected).
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
}

// ---- K3: obstacle sum at one predicted position ---------------------------------------------------------
// ox/oy are the problem's obstacle list staged in shared memory (all lanes read the same address: broadcast); lane = stage.
// The reference pads the list with copies of its first point (get_obstacles, scripts/point_follower_local_planner.py:
// 103-109; 160 x the sentinel (100, 100) when the scan hits nothing): the trailing run of copies of entry 0 is folded
// into a weight w0 on entry 0, so only the first n_eff entries are walked (ObsList, set up when the list is staged).
// Value, gradient and Hessian share one pass; the constant factors 2/r^2 of grad s and hess s are applied once after the
// loop:  with p1 = phi'(s), p2 = phi''(s), d = (x - ox, y - oy):
//     grad = (2/r^2) sum p1 d,     hess = (2/r^2)^2 sum p2 d d' + (2/r^2) (sum p1) I.
// One out-of-line copy (the solver calls it from the linearisation, the line search, the second-order correction and the
// restoration: inlined five times the loop alone was 20 KB of code for a 32 KB instruction cache).
struct ObsList {
    int n_eff;   // entries to walk (>= 1)
    double w0;   // weight of entry 0: 1 + number of folded trailing copies
};

// e^q for the obstacle terms (q = c/s, usually 0 < q < 100; +inf / NaN when s = 0).  Branch-free, so that the compiler
// interleaves the independent chains of the unrolled obstacle loop; Cody-Waite reduction by ln 2,
// degree-13 Taylor polynomial on [-ln2/2, ln2/2] (relative error < 1e-17 before rounding), scaling by 2^k in two factors
// (k reaches 1025 / -1075).  The coefficients sit in constant memory: as immediates each costs two instructions.
__constant__ double OBS_EXP_C[12] = {
    1.6059043836821613e-10, 2.0876756987868100e-09, 2.5052108385441720e-08, 2.7557319223985893e-07,   // 1/13! .. 1/10!
    2.7557319223985888e-06, 2.4801587301587302e-05, 1.9841269841269841e-04, 1.3888888888888889e-03,   // 1/9! .. 1/6!
    8.3333333333333332e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 1.4426950408889634};       // 1/5! 1/4! 1/3!, log2(e)
__constant__ double OBS_EXP_LN2[2] = {6.93147180369123816490e-01, 1.90821492927058770002e-10};
__device__ __forceinline__ double obs_exp(double q) {
    const double kf = rint(q * OBS_EXP_C[11]);
    double r = fma(-kf, OBS_EXP_LN2[0], q);
    r = fma(-kf, OBS_EXP_LN2[1], r);
    double p = fma(OBS_EXP_C[0], r, OBS_EXP_C[1]);
#pragma unroll
    for (int i = 2; i <= 10; i++) p = fma(p, r, OBS_EXP_C[i]);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    // 2^k in two factors; k is clamped so that the exponent arithmetic stays in range (the conversion saturates, NaN -> 0):
    // arguments beyond the overflow / underflow thresholds are replaced below, a NaN argument gives NaN through p
    const int k = max(min((int)kf, 1100), -1100), k1 = k >> 1, k2 = k - k1;
    double e = (p * __hiloint2double((k1 + 1023) << 20, 0)) * __hiloint2double((k2 + 1023) << 20, 0);
    e = (q > 709.782712893384) ? __longlong_as_double(0x7ff0000000000000ll) : e;
    e = (q < -745.2) ? 0.0 : e;
    return e;
}

// One term of the exp(c/s) sum.  ptxas lays independent FP64 chains out one after the other (it minimises registers; a
// four-term loop written side by side comes out term by term all the same), and a warp then issues one instruction per
// FP64 latency (profiles/r2_warp_a_v1, r2_lane_a_v1: 53-62 % of the obstacle loop's samples are fixed-latency waits).  So
// the parallelism is put INSIDE the term: the exponential's polynomial in Estrin form (4 levels instead of a 13-step Horner
// chain), the reciprocal with one second-order correction, and the derivative factors formed next to the reduction.
// c > 0 (the reference: reverse_factor 5.0 / cost_factor 0.5), so q = c/s >= 0 and 2^k is applied by an integer add to the
// exponent field (k <= 1024 whenever q is below the overflow threshold, and then p < 1: the result is a normal number).
struct ObsAcc {
    double v, sp1, ax, ay, bxx, bxy, byy;
};
template <bool USEW>
__device__ __forceinline__ void obs_term_explog(double c, double ir2, double x, double y, double ox, double oy, double w, ObsAcc &A) {
    const double dx = x - ox, dy = y - oy;
    const double s = (dx * dx + dy * dy) * ir2;
    double y0; // s = 0 (robot on an obstacle point): NaN instead of inf, invalid either way
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(s));
    const double e1 = fma(-s, y0, 1.0);   // |e1| <= 2^-23: 1/s = y0 (1 + e1 + e1^2) to 2^-69
    const double e2 = fma(e1, e1, e1);
    const double is = fma(y0, e2, y0);
    const double q = c * is;
    const double kf = rint(q * OBS_EXP_C[11]);
    const double t = q * is;              // c / s^2
    const double tt = fma(2.0, is, t);
    double r = fma(-kf, OBS_EXP_LN2[0], q);
    r = fma(-kf, OBS_EXP_LN2[1], r);
    // e^r, |r| <= ln2 / 2, degree 13 (relative error < 1e-17 before rounding)
    const double r2 = r * r;
    const double b0 = 1.0 + r, b1 = fma(OBS_EXP_C[10], r, 0.5), b2 = fma(OBS_EXP_C[8], r, OBS_EXP_C[9]),
                 b3 = fma(OBS_EXP_C[6], r, OBS_EXP_C[7]), b4 = fma(OBS_EXP_C[4], r, OBS_EXP_C[5]),
                 b5 = fma(OBS_EXP_C[2], r, OBS_EXP_C[3]), b6 = fma(OBS_EXP_C[0], r, OBS_EXP_C[1]);
    const double r4 = r2 * r2;
    const double c0 = fma(b1, r2, b0), c1 = fma(b3, r2, b2), c2 = fma(b5, r2, b4);
    const double r8 = r4 * r4;
    const double d0 = fma(c1, r4, c0), d1 = fma(b6, r4, c2);
    const double p = fma(d1, r8, d0);
    const int k = (int)kf; // (saturating conversion; NaN -> 0, and then p is NaN)
    double e = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    e = (q > 709.782712893384) ? __longlong_as_double(0x7ff0000000000000ll) : e;
    if (USEW) e *= w;
    const double te = t * e;              // -phi'
    const double p2 = te * tt;            // phi'' = (c/s^2)(2/s + c/s^2) phi
    const double p2x = p2 * dx;
    A.v += e;
    A.sp1 -= te;
    A.ax = fma(-te, dx, A.ax);
    A.ay = fma(-te, dy, A.ay);
    A.bxx = fma(p2x, dx, A.bxx);
    A.bxy = fma(p2x, dy, A.bxy);
    A.byy = fma(p2 * dy, dy, A.byy);
}

__device__ __noinline__ void obstacle_eval(int form, double c, double ir2, const double *__restrict__ sox,
                                           const double *__restrict__ soy, int n_eff, double w0, double x, double y,
                                           double *out6) {
    double v = 0, sp1 = 0, ax = 0, ay = 0, bxx = 0, bxy = 0, byy = 0;
    if (form == B200MPC_OBS_EXPLOG) {
        // phi(s) = exp(c/s):  phi' = -(c/s^2) phi,  phi'' = (c/s^2)(2/s + c/s^2) phi
#define OBS_TERM_EXPLOG(J_, W_, USEW_)                                                                  \
    do {                                                                                                \
        const double dx = x - sox[J_], dy = y - soy[J_];                                                \
        const double s_ = (dx * dx + dy * dy) * ir2;                                                    \
        const double is = fast_rcp(s_); /* s = 0 (robot on an obstacle point): NaN instead of inf, invalid either way */ \
        const double q = c * is;                                                                        \
        double e = obs_exp(q);                                                                          \
        if (USEW_) e *= (W_);                                                                           \
        const double t = q * is;               /* c / s^2 */                                            \
        const double te = t * e;               /* -phi' */                                              \
        const double p2 = te * fma(2.0, is, t);                                                         \
        const double p2x = p2 * dx;                                                                     \
        v += e;                                                                                         \
        sp1 -= te;                                                                                      \
        ax = fma(-te, dx, ax);                                                                          \
        ay = fma(-te, dy, ay);                                                                          \
        bxx = fma(p2x, dx, bxx);                                                                        \
        bxy = fma(p2x, dy, bxy);                                                                        \
        byy = fma(p2 * dy, dy, byy);                                                                    \
    } while (0)
        if (c > 0.0) {
            ObsAcc A;
            A.v = 0; A.sp1 = 0; A.ax = 0; A.ay = 0; A.bxx = 0; A.bxy = 0; A.byy = 0;
            obs_term_explog<true>(c, ir2, x, y, sox[0], soy[0], w0, A);
#pragma unroll 4
            for (int j = 1; j < n_eff; j++) obs_term_explog<false>(c, ir2, x, y, sox[j], soy[j], 1.0, A);
            v = A.v; sp1 = A.sp1; ax = A.ax; ay = A.ay; bxx = A.bxx; bxy = A.bxy; byy = A.byy;
        } else {
            OBS_TERM_EXPLOG(0, w0, true);
#pragma unroll 1
            for (int j = 1; j < n_eff; j++) OBS_TERM_EXPLOG(j, 1.0, false);
        }
#undef OBS_TERM_EXPLOG
    } else {
        // psi(s) = c exp(-s):  psi' = -psi,  psi'' = psi
#define OBS_TERM_GAUSS(J_, W_, USEW_)                                                                   \
    do {                                                                                                \
        const double dx = x - sox[J_], dy = y - soy[J_];                                                \
        const double s_ = (dx * dx + dy * dy) * ir2;                                                    \
        double e = c * exp(-s_);                                                                        \
        if (USEW_) e *= (W_);                                                                           \
        const double p2x = e * dx;                                                                      \
        v += e;                                                                                         \
        sp1 -= e;                                                                                       \
        ax = fma(-e, dx, ax);                                                                           \
        ay = fma(-e, dy, ay);                                                                           \
        bxx = fma(p2x, dx, bxx);                                                                        \
        bxy = fma(p2x, dy, bxy);                                                                        \
        byy = fma(e * dy, dy, byy);                                                                     \
    } while (0)
        OBS_TERM_GAUSS(0, w0, true);
#pragma unroll 4
        for (int j = 1; j < n_eff; j++) OBS_TERM_GAUSS(j, 1.0, false);
#undef OBS_TERM_GAUSS
    }
    const double g1 = 2.0 * ir2, g2 = g1 * g1;
    out6[0] = v;
    out6[1] = g1 * ax;
    out6[2] = g1 * ay;
    out6[3] = fma(g2, bxx, g1 * sp1);
    out6[4] = g2 * bxy;
    out6[5] = fma(g2, byy, g1 * sp1);
}

template <bool DERIV>
__device__ __forceinline__ void obstacle_sum(const KParams &P, const ObsList &L, const double *__restrict__ sox,
                                             const double *__restrict__ soy, double x, double y, double &val,
                                             double &gx, double &gy, double &hxx, double &hxy, double &hyy) {
    double o[6];
    obstacle_eval(P.obs_form, P.obs_c, P.inv_r2, sox, soy, L.n_eff, L.w0, x, y, o);
    val = o[0];
    if (DERIV) { gx = o[1]; gy = o[2]; hxx = o[3]; hxy = o[4]; hyy = o[5]; }
}

// The list has been staged in shared memory by the whole warp: fold the trailing copies of entry 0 (warp-uniform result).
__device__ __forceinline__ ObsList obstacle_list_setup(const double *sox, const double *soy, int M, int lane) {
    const double x0 = sox[0], y0 = soy[0];
    int last = 0; // highest index whose entry differs from entry 0
    for (int i = lane; i < M; i += 32)
        if (sox[i] != x0 || soy[i] != y0) last = i; // (a NaN entry differs from everything: it is walked)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(FULL, last, o));
    ObsList L;
    L.n_eff = last + 1;
    L.w0 = 1.0 + (double)(M - L.n_eff);
    return L;
}

// ---- per-stage register state ----------------------------------------------------------------------------
struct Step {
    double dX[3], dU[2], dS[2], dlam[3], dyd[2];
};

struct Stg {
    // iterate
    double X[3], U[2], S[2], lam[3], yd[2], vL[2], vU[2];
    // references
    double r[3], ub[2];
    // K2 stage derivatives: A = I + a13 e1e3' + a23 e2e3',  B = [[b11,b12],[b21,b22],[0,dt]]
    double a13, a23, b11, b12, b21, b22;
    // Lagrangian Hessian nonzeros on (x,y,th,v,w)
    double hxx, hxy, hyy, htt, htv, htw, hvv, hvw, hww;
    double g[5];  // scaled objective gradient
    double c[3];  // c_{k+1} = X_{k+1} - F(X_k,U_k)
    double rx[3], ru[2]; // grad_lag_x
    double Dsig[2], rs[2];
    // Riccati factors
    double K[6], kf[2], P[6], pv[3];
    // trial point
    double Xt[3], Ut[2], St[2], ct[3];
};

// K1+K2: value of the integration step F(x,u) (closed form of the staged RK4, identical to k1..k4 of rk4())
__device__ __forceinline__ void dyn_value(const KParams &P, const double X[3], const double U[2], double F[3]) {
    const double dt = P.dt, th = X[2], v = U[0], w = U[1];
    if (P.integrator == B200MPC_EULER) {
        double sn, cs;
        sincos(th, &sn, &cs);
        F[0] = X[0] + dt * v * cs;
        F[1] = X[1] + dt * v * sn;
        F[2] = th + dt * w;
    } else {
        double s0, c0, sm, cm, se, ce;
        rk4_trig(th, 0.5 * dt * w, s0, c0, sm, cm, se, ce);
        const double h = dt / 6.0;
        F[0] = X[0] + h * v * (c0 + 4.0 * cm + ce);
        F[1] = X[1] + h * v * (s0 + 4.0 * sm + se);
        F[2] = th + dt * w;
    }
}

// Full stage evaluation: defect, Jacobian entries, objective gradient, Lagrangian Hessian, stage cost.
// act: stage exists (k <= N); dyn: stage has controls and a successor (k < N); obs: obstacle sum on this stage.
__device__ __forceinline__ double stage_full(const KParams &P, const ObsList &OL, const double *sox, const double *soy, Stg &s,
                                             const double Xn[3], const double ln[3], double df, bool act,
                                             bool dyn, bool obs, const double *oc = nullptr) {
    double fval = 0;
    s.a13 = s.a23 = s.b11 = s.b12 = s.b21 = s.b22 = 0;
    s.hxx = s.hxy = s.hyy = s.htt = s.htv = s.htw = s.hvv = s.hvw = s.hww = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) s.g[i] = 0;
    s.c[0] = s.c[1] = s.c[2] = 0;
    if (dyn) {
        const double dt = P.dt, th = s.X[2], v = s.U[0], w = s.U[1];
        double F0, F1, d2x[6], d2y[6];
        if (P.integrator == B200MPC_EULER) {
            double sn, cs;
            sincos(th, &sn, &cs);
            F0 = s.X[0] + dt * v * cs;
            F1 = s.X[1] + dt * v * sn;
            s.a13 = -dt * v * sn; s.a23 = dt * v * cs;
            s.b11 = dt * cs; s.b21 = dt * sn;
            d2x[0] = -dt * v * cs; d2x[1] = -dt * sn; d2x[2] = d2x[3] = d2x[4] = d2x[5] = 0;
            d2y[0] = -dt * v * sn; d2y[1] = dt * cs; d2y[2] = d2y[3] = d2y[4] = d2y[5] = 0;
        } else {
            double s0, c0, sm, cm, se, ce;
            rk4_trig(th, 0.5 * dt * w, s0, c0, sm, cm, se, ce);
            const double h = dt / 6.0;
            const double C = c0 + 4.0 * cm + ce, S = s0 + 4.0 * sm + se;
            const double C1 = 2.0 * cm + ce, S1 = 2.0 * sm + se, C2 = cm + ce, S2 = sm + se;
            F0 = s.X[0] + h * v * C;
            F1 = s.X[1] + h * v * S;
            s.a13 = -h * v * S; s.a23 = h * v * C;
            s.b11 = h * C; s.b21 = h * S;
            s.b12 = -h * dt * v * S1; s.b22 = h * dt * v * C1;
            d2x[0] = -h * v * C;        d2y[0] = -h * v * S;
            d2x[1] = -h * S;            d2y[1] = h * C;
            d2x[2] = -h * dt * v * C1;  d2y[2] = -h * dt * v * S1;
            d2x[3] = 0;                 d2y[3] = 0;
            d2x[4] = -h * dt * S1;      d2y[4] = h * dt * C1;
            d2x[5] = -h * dt * dt * v * C2; d2y[5] = -h * dt * dt * v * S2;
        }
        s.c[0] = Xn[0] - F0;
        s.c[1] = Xn[1] - F1;
        s.c[2] = Xn[2] - (th + dt * w);
        // stage cost (define_cost_function): tracking + control + reverse penalty
        const double e0 = s.X[0] - s.r[0], e1 = s.X[1] - s.r[1], e2 = s.X[2] - s.r[2];
        const double m0 = v - s.ub[0], m1 = w - s.ub[1];
        const double er = exp(-P.kappa * v);
        fval = e0 * P.Q[0] * e0 + e1 * P.Q[1] * e1 + e2 * P.Q[2] * e2 + m0 * P.R[0] * m0 + m1 * P.R[1] * m1 + er;
        s.g[0] = df * 2.0 * P.Q[0] * e0;
        s.g[1] = df * 2.0 * P.Q[1] * e1;
        s.g[2] = df * 2.0 * P.Q[2] * e2;
        s.g[3] = df * (2.0 * P.R[0] * m0 - P.kappa * er);
        s.g[4] = df * 2.0 * P.R[1] * m1;
        s.hxx = df * 2.0 * P.Q[0];
        s.hyy = df * 2.0 * P.Q[1];
        s.htt = df * 2.0 * P.Q[2] - (ln[0] * d2x[0] + ln[1] * d2y[0]);
        s.htv = -(ln[0] * d2x[1] + ln[1] * d2y[1]);
        s.htw = -(ln[0] * d2x[2] + ln[1] * d2y[2]);
        s.hvv = df * (2.0 * P.R[0] + P.kappa * P.kappa * er) - (ln[0] * d2x[3] + ln[1] * d2y[3]);
        s.hvw = -(ln[0] * d2x[4] + ln[1] * d2y[4]);
        s.hww = df * 2.0 * P.R[1] - (ln[0] * d2x[5] + ln[1] * d2y[5]);
    }
    if (obs && act) {
        double ov, gx, gy, oxx, oxy, oyy;
        if (oc) { ov = oc[0]; gx = oc[1]; gy = oc[2]; oxx = oc[3]; oxy = oc[4]; oyy = oc[5]; }
        else obstacle_sum<true>(P, OL, sox, soy, s.X[0], s.X[1], ov, gx, gy, oxx, oxy, oyy);
        fval += ov;
        s.g[0] += df * gx;
        s.g[1] += df * gy;
        s.hxx += df * oxx;
        s.hxy += df * oxy;
        s.hyy += df * oyy;
    }
    return fval;
}

// Value-only evaluation of a trial stage: defect (3) and stage cost.
__device__ __forceinline__ double stage_value(const KParams &P, const ObsList &OL, const double *sox, const double *soy,
                                              const double X[3], const double U[2], const double r[3],
                                              const double ub[2], const double Xn[3], double ct[3], bool act,
                                              bool dyn, bool obs) {
    double fval = 0;
    ct[0] = ct[1] = ct[2] = 0;
    if (dyn) {
        double F[3];
        dyn_value(P, X, U, F);
        ct[0] = Xn[0] - F[0];
        ct[1] = Xn[1] - F[1];
        ct[2] = Xn[2] - F[2];
        const double e0 = X[0] - r[0], e1 = X[1] - r[1], e2 = X[2] - r[2];
        const double m0 = U[0] - ub[0], m1 = U[1] - ub[1];
        fval = e0 * P.Q[0] * e0 + e1 * P.Q[1] * e1 + e2 * P.Q[2] * e2 + m0 * P.R[0] * m0 + m1 * P.R[1] * m1 +
               exp(-P.kappa * U[0]);
    }
    if (obs && act) {
        double ov, d0, d1, d2, d3, d4;
        obstacle_sum<false>(P, OL, sox, soy, X[0], X[1], ov, d0, d1, d2, d3, d4);
        fval += ov;
    }
    return fval;
}

// value of the next stage (k+1) for slot j: same lane if j+1 < J, else slot 0 of lane+1
#define NEXT3(dst, field, j)                                                               \
    do {                                                                                   \
        if ((j) + 1 < J) {                                                                 \
            dst[0] = s[((j) + 1 < J) ? (j) + 1 : 0].field[0];                              \
            dst[1] = s[((j) + 1 < J) ? (j) + 1 : 0].field[1];                              \
            dst[2] = s[((j) + 1 < J) ? (j) + 1 : 0].field[2];                              \
        } else {                                                                           \
            dst[0] = __shfl_down_sync(FULL, s[0].field[0], 1);                             \
            dst[1] = __shfl_down_sync(FULL, s[0].field[1], 1);                             \
            dst[2] = __shfl_down_sync(FULL, s[0].field[2], 1);                             \
        }                                                                                  \
    } while (0)

// ---- K4: Riccati solve of the condensed stage-wise KKT system --------------------------------------------
// Solves (IPOPT augmented system with ds, dyd eliminated stage-locally):
//   (useW*W + dw) dx + Jc' dyc + Jd' dyd = -rx,  Dsig ds - dyd = -rs,  Jc dx = -rc,  Jd dx - ds = -rd
// rc[j] = defect rhs of c_{k+1} held with stage k, rd[j] = rhs of U_k - S_k.
// Returns false when a condensed Quu block is not positive definite (wrong inertia).
template <int J>
__device__ __forceinline__ bool kkt_solve(const KParams &P, Stg (&s)[J], const double (&rc)[J][3],
                                          const double (&rd)[J][2], bool useW, double dw, Step (&o)[J],
                                          int lane) {
    const int N = P.N;
    const double dt = P.dt;
    const int lN = N / J;
    double P00 = 0, P01 = 0, P02 = 0, P11 = 0, P12 = 0, P22 = 0, p0 = 0, p1 = 0, p2 = 0;
    int okall = 1;
    for (int l = lN; l >= 0; --l) {
        double q00 = P00, q01 = P01, q02 = P02, q11 = P11, q12 = P12, q22 = P22, v0 = p0, v1 = p1, v2 = p2;
        int ok = 1;
        const bool mine = (lane == l);
#pragma unroll
        for (int j = J - 1; j >= 0; --j) {
            const int ko = l * J + j; // stage of the owner lane (uniform)
            if (ko > N) continue;
            Stg &t = s[j];
            const double hxx = useW ? t.hxx : 0.0, hxy = useW ? t.hxy : 0.0, hyy = useW ? t.hyy : 0.0;
            const double htt = useW ? t.htt : 0.0, htv = useW ? t.htv : 0.0, htw = useW ? t.htw : 0.0;
            const double hvv = useW ? t.hvv : 0.0, hvw = useW ? t.hvw : 0.0, hww = useW ? t.hww : 0.0;
            if (ko == N) {
                q00 = hxx + dw; q01 = hxy; q02 = 0; q11 = hyy + dw; q12 = 0; q22 = htt + dw;
                v0 = t.rx[0]; v1 = t.rx[1]; v2 = t.rx[2];
            } else {
                const double a = t.a13, b = t.a23, b11 = t.b11, b12 = t.b12, b21 = t.b21, b22 = t.b22;
                const double d0 = -rc[j][0], d1 = -rc[j][1], d2 = -rc[j][2];
                // w = P d + p
                const double w0 = q00 * d0 + q01 * d1 + q02 * d2 + v0;
                const double w1 = q01 * d0 + q11 * d1 + q12 * d2 + v1;
                const double w2 = q02 * d0 + q12 * d1 + q22 * d2 + v2;
                // t = P[:,2] + a P[:,0] + b P[:,1]
                const double t0 = q02 + a * q00 + b * q01;
                const double t1 = q12 + a * q01 + b * q11;
                const double t2 = q22 + a * q02 + b * q12;
                // Qxx = Hxx + A'PA
                const double x00 = hxx + dw + q00, x01 = hxy + q01, x11 = hyy + dw + q11;
                const double x02 = t0, x12 = t1, x22 = htt + dw + t2 + a * t0 + b * t1;
                // PB columns
                const double e0 = b11 * q00 + b21 * q01, e1 = b11 * q01 + b21 * q11;
                const double f0 = b12 * q00 + b22 * q01 + dt * q02, f1 = b12 * q01 + b22 * q11 + dt * q12,
                             f2 = b12 * q02 + b22 * q12 + dt * q22;
                // Qux = Hux + B'PA
                const double u00 = e0, u01 = e1, u02 = htv + b11 * t0 + b21 * t1;
                const double u10 = f0, u11 = f1, u12 = htw + b12 * t0 + b22 * t1 + dt * t2;
                // Quu = Huu + Dsig + B'PB
                const double r00 = hvv + dw + t.Dsig[0] + b11 * e0 + b21 * e1;
                const double r01 = hvw + b11 * f0 + b21 * f1;
                const double r11 = hww + dw + t.Dsig[1] + b12 * f0 + b22 * f1 + dt * f2;
                // gradients
                const bool first = (ko == 0);
                const double gx0 = (first ? 0.0 : t.rx[0]) + w0;
                const double gx1 = (first ? 0.0 : t.rx[1]) + w1;
                const double gx2 = (first ? 0.0 : t.rx[2]) + a * w0 + b * w1 + w2;
                const double qu0 = t.ru[0] + t.Dsig[0] * rd[j][0] + t.rs[0];
                const double qu1 = t.ru[1] + t.Dsig[1] * rd[j][1] + t.rs[1];
                const double gu0 = qu0 + b11 * w0 + b21 * w1;
                const double gu1 = qu1 + b12 * w0 + b22 * w1 + dt * w2;
                const double det = r00 * r11 - r01 * r01;
                if (!(r00 > 0.0) || !(det > 0.0)) ok = 0;
                const double idet = fast_rcp(det);
                const double i00 = r11 * idet, i01 = -r01 * idet, i11 = r00 * idet;
                const double K00 = -(i00 * u00 + i01 * u10), K01 = -(i00 * u01 + i01 * u11),
                             K02 = -(i00 * u02 + i01 * u12);
                const double K10 = -(i01 * u00 + i11 * u10), K11 = -(i01 * u01 + i11 * u11),
                             K12 = -(i01 * u02 + i11 * u12);
                const double k0 = -(i00 * gu0 + i01 * gu1), k1 = -(i01 * gu0 + i11 * gu1);
                if (mine) {
                    t.K[0] = K00; t.K[1] = K01; t.K[2] = K02; t.K[3] = K10; t.K[4] = K11; t.K[5] = K12;
                    t.kf[0] = k0; t.kf[1] = k1;
                }
                // P' = Qxx + Qux'K (symmetrised), p' = gx + Qux' kf
                q00 = x00 + u00 * K00 + u10 * K10;
                q11 = x11 + u01 * K01 + u11 * K11;
                q22 = x22 + u02 * K02 + u12 * K12;
                q01 = x01 + 0.5 * ((u00 * K01 + u10 * K11) + (u01 * K00 + u11 * K10));
                q02 = x02 + 0.5 * ((u00 * K02 + u10 * K12) + (u02 * K00 + u12 * K10));
                q12 = x12 + 0.5 * ((u01 * K02 + u11 * K12) + (u02 * K01 + u12 * K11));
                v0 = gx0 + u00 * k0 + u10 * k1;
                v1 = gx1 + u01 * k0 + u11 * k1;
                v2 = gx2 + u02 * k0 + u12 * k1;
            }
            if (mine) {
                t.P[0] = q00; t.P[1] = q01; t.P[2] = q02; t.P[3] = q11; t.P[4] = q12; t.P[5] = q22;
                t.pv[0] = v0; t.pv[1] = v1; t.pv[2] = v2;
            }
        }
        P00 = __shfl_sync(FULL, q00, l); P01 = __shfl_sync(FULL, q01, l); P02 = __shfl_sync(FULL, q02, l);
        P11 = __shfl_sync(FULL, q11, l); P12 = __shfl_sync(FULL, q12, l); P22 = __shfl_sync(FULL, q22, l);
        p0 = __shfl_sync(FULL, v0, l); p1 = __shfl_sync(FULL, v1, l); p2 = __shfl_sync(FULL, v2, l);
        okall = __shfl_sync(FULL, ok, l);
        if (!okall) return false;
    }
    // forward sweep
    double x0 = 0, x1 = 0, x2 = 0;
    for (int l = 0; l <= lN; ++l) {
        double y0 = x0, y1 = x1, y2 = x2;
        const bool mine = (lane == l);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int ko = l * J + j;
            if (ko > N) continue;
            Stg &t = s[j];
            if (mine) { o[j].dX[0] = y0; o[j].dX[1] = y1; o[j].dX[2] = y2; }
            if (ko < N) {
                const double du0 = t.kf[0] + t.K[0] * y0 + t.K[1] * y1 + t.K[2] * y2;
                const double du1 = t.kf[1] + t.K[3] * y0 + t.K[4] * y1 + t.K[5] * y2;
                if (mine) { o[j].dU[0] = du0; o[j].dU[1] = du1; }
                const double n0 = y0 + t.a13 * y2 + t.b11 * du0 + t.b12 * du1 - rc[j][0];
                const double n1 = y1 + t.a23 * y2 + t.b21 * du0 + t.b22 * du1 - rc[j][1];
                const double n2 = y2 + dt * du1 - rc[j][2];
                y0 = n0; y1 = n1; y2 = n2;
            }
        }
        x0 = __shfl_sync(FULL, y0, l); x1 = __shfl_sync(FULL, y1, l); x2 = __shfl_sync(FULL, y2, l);
    }
    // multiplier and slack steps (stage-parallel)
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int k = lane * J + j;
        Stg &t = s[j];
        const bool act = (k <= N), dyn = (k < N);
        const double a0 = o[j].dX[0], a1 = o[j].dX[1], a2 = o[j].dX[2];
        const double l0 = -(t.pv[0] + t.P[0] * a0 + t.P[1] * a1 + t.P[2] * a2);
        const double l1 = -(t.pv[1] + t.P[1] * a0 + t.P[3] * a1 + t.P[4] * a2);
        const double l2 = -(t.pv[2] + t.P[2] * a0 + t.P[4] * a1 + t.P[5] * a2);
        const bool hasl = act && (k >= 1);
        o[j].dlam[0] = hasl ? l0 : 0.0; o[j].dlam[1] = hasl ? l1 : 0.0; o[j].dlam[2] = hasl ? l2 : 0.0;
        if (!act) { o[j].dX[0] = o[j].dX[1] = o[j].dX[2] = 0; }
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const double ds = dyn ? (o[j].dU[i] + rd[j][i]) : 0.0;
            o[j].dS[i] = ds;
            o[j].dyd[i] = dyn ? (t.Dsig[i] * ds + t.rs[i]) : 0.0;
            if (!dyn) o[j].dU[i] = 0;
        }
    }
    return true;
}

// ---- K4, lane-parallel form (default) ------------------------------------------------------------------------------
// kkt_solve above lets the lane that owns stage k do the whole 3x3 / 3x2 block algebra of the stage while the other
// 31 lanes execute the same instructions on garbage: ~380 warp instructions per stage on the serial critical path
// (65 % of all instructions of a solve, profiles/r1_warp_b1_*).  Here the stage data are first written to shared
// memory by their owners (stage-parallel), and the serial recursion then spreads the block algebra of ONE stage over
// the lanes of the warp, one matrix entry per lane, in the generic form
//     z = (x, u),  G = [A B] (3x5),  M = H + G' P G,  m = h + G'(P d + p),
//     P' = Mxx - Mxu Muu^-1 Mux,  p' = mx - Mxu Muu^-1 mu,  K = -Muu^-1 Mux,  kf = -Muu^-1 mu:
//   step A  lanes 0-14: T = P G (entry (i,c));              lanes 15-17: w = P d + p
//   step B  lanes 0-14: M (15 symmetric pairs (a,c));        lanes 15-19: m (entry a)
//   step C  all lanes:  2x2 inverse of Muu (inertia test: Muu positive definite)
//   step E  lanes 0-5: P' pairs; 6-8: p'; 9-14: K; 15-16: kf   -> shared memory (forward sweep, multiplier step)
// Values move between the steps with warp shuffles whose source lanes come from a per-lane role table; the structural
// entries of G (0, 1, dt) are read from three constant slots of the stage record, so every lane runs the same dense
// formulas.  ~90 warp instructions and ~300 cycles per stage.  The forward sweep needs no communication at all:
// every lane rolls the whole step out redundantly from the gains in shared memory and keeps the entries of its own
// stages.  The terminal stage is the same step with P = 0 coming in and Muu = I.
#define KKT_REC 50          /* doubles per stage: 32 record + 18 results */
#define KR_ZERO 0
#define KR_ONE 1
#define KR_DT 2
#define KR_G 3              /* a13 a23 b11 b12 b21 b22 */
#define KR_H 9              /* 15 symmetric pairs */
#define KR_h 24             /* 5 */
#define KR_D 29             /* 3: d = -rc */
#define KR_OUT 32           /* K[2][3] kf[2] P'[6] p'[3] */

struct KktRoles {
    unsigned a_off;  // record offsets of the three step-A operands (G[0..2][c] or d), one byte each
    unsigned a_src;  // source lanes (previous stage's step E) of P[i][0..2] and p[i]
    unsigned b_src;  // source lanes of T[0..2][c] (or w), byte 3: record offset of H[a][c] / h[a]
    unsigned b_off;  // record offsets of G[0..2][a], byte 3: result offset (0xff = none)
    unsigned e_src;  // source lanes of Mxu[a][0], Mxu[a][1], Mxu[b][0], Mxu[b][1]
    unsigned e_misc; // byte 0: source lane of m[a]; byte 1: role (0 P' diag, 1 P' off-diag, 2 p', 3 K row 0, 4 K row 1,
                     //          5 kf0, 6 kf1, 7 none); byte 2: 1 if the lane adds p[i] in step A
};

__device__ __forceinline__ unsigned kkt_g_off(int j, int c) {
    // record offset of G[j][c], G = [A B], A = I + a e1 e3' + b e2 e3', B = [[b11,b12],[b21,b22],[0,dt]]
    if (c < 2) return (j == c) ? KR_ONE : KR_ZERO;
    if (c == 2) return (j == 0) ? KR_G + 0 : (j == 1) ? KR_G + 1 : KR_ONE;
    if (c == 3) return (j == 0) ? KR_G + 2 : (j == 1) ? KR_G + 4 : KR_ZERO;
    return (j == 0) ? KR_G + 3 : (j == 1) ? KR_G + 5 : KR_DT;
}
__device__ __forceinline__ int kkt_sym3(int i, int j) { // lane holding P'[i][j]
    if (i > j) { const int t = i; i = j; j = t; }
    return (i == 0) ? j : (i == 1) ? 2 + j : 5;
}
__device__ __forceinline__ void kkt_pair(int e, int &a, int &c) { // the 15 symmetric pairs of M
    const int pa[15] = {0, 0, 0, 1, 1, 2, 0, 0, 1, 1, 2, 2, 3, 3, 4};
    const int pc[15] = {0, 1, 2, 1, 2, 2, 3, 4, 3, 4, 3, 4, 3, 4, 4};
    a = pa[e]; c = pc[e];
}

__device__ __forceinline__ KktRoles kkt_roles(int lane) {
    KktRoles r;
    // step A
    int i = 0;
    unsigned addp = 0;
    if (lane < 15) {
        i = lane / 5;
        const int c = lane % 5;
        r.a_off = kkt_g_off(0, c) | (kkt_g_off(1, c) << 8) | (kkt_g_off(2, c) << 16);
    } else {
        i = (lane < 18) ? lane - 15 : 0;
        addp = (lane < 18) ? 1u : 0u;
        r.a_off = (KR_D + 0) | ((KR_D + 1) << 8) | ((KR_D + 2) << 16);
    }
    r.a_src = (unsigned)kkt_sym3(i, 0) | ((unsigned)kkt_sym3(i, 1) << 8) | ((unsigned)kkt_sym3(i, 2) << 16) | ((unsigned)(6 + i) << 24);
    // step B
    unsigned out = 0xff;
    if (lane < 15) {
        int a, c;
        kkt_pair(lane, a, c);
        r.b_src = (unsigned)(0 + c) | ((unsigned)(5 + c) << 8) | ((unsigned)(10 + c) << 16) | ((unsigned)(KR_H + lane) << 24);
        r.b_off = kkt_g_off(0, a) | (kkt_g_off(1, a) << 8) | (kkt_g_off(2, a) << 16);
    } else {
        const int a = (lane < 20) ? lane - 15 : 0;
        r.b_src = 15u | (16u << 8) | (17u << 16) | ((unsigned)(KR_h + a) << 24);
        r.b_off = kkt_g_off(0, a) | (kkt_g_off(1, a) << 8) | (kkt_g_off(2, a) << 16);
    }
    // step E
    int ea = 0, eb = 0, role = 7;
    if (lane < 6) { kkt_pair(lane, ea, eb); role = (ea == eb) ? 0 : 1; out = 8 + lane; }
    else if (lane < 9) { ea = eb = lane - 6; role = 2; out = 14 + (lane - 6); }
    else if (lane < 15) { ea = eb = (lane - 9) % 3; role = 3 + (lane - 9) / 3; out = lane - 9; }
    else if (lane < 17) { role = 5 + (lane - 15); out = 6 + (lane - 15); }
    r.e_src = (unsigned)(6 + 2 * ea) | ((unsigned)(7 + 2 * ea) << 8) | ((unsigned)(6 + 2 * eb) << 16) | ((unsigned)(7 + 2 * eb) << 24);
    r.e_misc = (unsigned)(15 + ea) | ((unsigned)role << 8) | (addp << 16);
    r.b_off |= out << 24;
    return r;
}


// ---- K4, parallel-in-time form of the backward recursion (single-solve latency; N + 1 <= 32) --------------------------
// The serial recursion costs ~700 cycles per stage on its critical path (dependent FP64 operations at 11.5 cycles each,
// profiles/r1_warp_b1_*).  Here the cost-to-go of EVERY stage comes out of a five-level suffix scan over the lanes
// (Saerkkae & Garcia-Fernandez, "Temporal parallelization of dynamic programming and linear quadratic control", 2021):
// lane k holds the conditional value function of the stages k..j-1,
//     V(x_k, x_j) = max_lam  x_k'J x_k / 2 - x_k'eta - lam'C lam / 2 - lam'(x_j - A x_k - b),
// one stage is  A = A_k - B R^-1 S',  b = d_k - B R^-1 r_u,  C = B R^-1 B',  eta = -(r_x - S R^-1 r_u),
// J = Hxx - S R^-1 S'  (R = Huu + Sigma + delta_w, S = Hxu), and two adjacent ranges combine associatively as
//     M = (I + C1 J2)^-1,  A = A2 M A1,  b = A2 M (b1 + C1 eta2) + b2,  C = A2 M C1 A2' + C2,
//     eta = A1' M'(eta2 - J2 b1) + eta1,  J = A1' M' J2 A1 + J1.
// After the scan lane k holds (P_k, p_k) = (J, -eta) of the range k..N; the gains, the inertia test (Quu_k positive
// definite) and P'_k then follow from P_{k+1} with the formulas of the serial recursion, all stages at once.
// Algebraically this is the same block elimination in another order WHEN every local block R_k is nonsingular; it is
// only used when all R_k are safely positive definite (return value -1 otherwise: the caller then runs the serial
// recursion, which needs only the condensed Quu_k to be positive definite).  The values differ at rounding level.  Out of line, operands and results in shared memory: the scan needs ~150
// registers of its own, which the solver's per-stage state would otherwise be spilled for.
#define SCAN_NF 27 /* A[9] b[3] C[6] eta[3] J[6] */
__device__ __noinline__ int kkt_backward_scan(double *rec, double *el, int N, double dt, int lane) {
    const int k = lane;
    const bool act = k <= N, dyn = k < N;
    const double *q = rec + (size_t)(act ? k : 0) * KKT_REC;
    double A[9], b[3], C[6], e[3], Jm[6];
    int r_ok = 1;
    if (dyn) {
        const double a13 = q[KR_G + 0], a23 = q[KR_G + 1], b11 = q[KR_G + 2], b12 = q[KR_G + 3], b21 = q[KR_G + 4], b22 = q[KR_G + 5];
        const double R00 = q[KR_H + 12], R01 = q[KR_H + 13], R11 = q[KR_H + 14], htv = q[KR_H + 10], htw = q[KR_H + 11];
        const double Rdet = R00 * R11 - R01 * R01;
        // the scan eliminates the controls of every stage LOCALLY, so it needs R_k itself to be safely positive definite
        // (the serial recursion only needs the condensed Quu_k = R_k + B'PB): otherwise the caller takes the serial form
        if (!(R00 > 0.0) || !(R11 > 0.0) || !(Rdet > 1e-8 * R00 * R11)) r_ok = 0;
        const double idet = fast_rcp(Rdet);
        const double i00 = R11 * idet, i01 = -R01 * idet, i11 = R00 * idet;
        // B R^-1 (3x2)
        const double g00 = b11 * i00 + b12 * i01, g01 = b11 * i01 + b12 * i11;
        const double g10 = b21 * i00 + b22 * i01, g11 = b21 * i01 + b22 * i11;
        const double g20 = dt * i01, g21 = dt * i11;
        const double ru0 = q[KR_h + 3], ru1 = q[KR_h + 4];
        // A~ = A - B R^-1 S', S = e3 (htv, htw)
        A[0] = 1; A[1] = 0; A[2] = a13 - (g00 * htv + g01 * htw);
        A[3] = 0; A[4] = 1; A[5] = a23 - (g10 * htv + g11 * htw);
        A[6] = 0; A[7] = 0; A[8] = 1.0 - (g20 * htv + g21 * htw);
        b[0] = q[KR_D + 0] - (g00 * ru0 + g01 * ru1);
        b[1] = q[KR_D + 1] - (g10 * ru0 + g11 * ru1);
        b[2] = q[KR_D + 2] - (g20 * ru0 + g21 * ru1);
        C[0] = g00 * b11 + g01 * b12; C[1] = g00 * b21 + g01 * b22; C[2] = g01 * dt;
        C[3] = g10 * b21 + g11 * b22; C[4] = g11 * dt; C[5] = g21 * dt;
        const double s0 = htv * i00 + htw * i01, s1 = htv * i01 + htw * i11;
        Jm[0] = q[KR_H + 0]; Jm[1] = q[KR_H + 1]; Jm[2] = q[KR_H + 2]; Jm[3] = q[KR_H + 3]; Jm[4] = q[KR_H + 4];
        Jm[5] = q[KR_H + 5] - (s0 * htv + s1 * htw);
        e[0] = -q[KR_h + 0]; e[1] = -q[KR_h + 1]; e[2] = -(q[KR_h + 2] - (s0 * ru0 + s1 * ru1));
    } else {
        // terminal stage (and idle lanes): A = 0, b = 0, C = 0, eta = -r_x, J = Hxx
#pragma unroll
        for (int i = 0; i < 9; i++) A[i] = 0;
        b[0] = b[1] = b[2] = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) { C[i] = 0; Jm[i] = q[KR_H + i]; }
        e[0] = -q[KR_h + 0]; e[1] = -q[KR_h + 1]; e[2] = -q[KR_h + 2];
    }
    if (!__all_sync(FULL, r_ok)) return -1; // the scan does not apply to this system: the caller takes the serial recursion
#pragma unroll 1
    for (int lvl = 1; lvl <= N; lvl <<= 1) {
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 9; i++) el[i * 32 + lane] = A[i];
#pragma unroll
        for (int i = 0; i < 3; i++) { el[(9 + i) * 32 + lane] = b[i]; el[(18 + i) * 32 + lane] = e[i]; }
#pragma unroll
        for (int i = 0; i < 6; i++) { el[(12 + i) * 32 + lane] = C[i]; el[(21 + i) * 32 + lane] = Jm[i]; }
        __syncwarp();
        const int pj = lane + lvl;
        if (pj <= N) {
            double A2[9], b2[3], C2[6], e2[3], J2[6];
#pragma unroll
            for (int i = 0; i < 9; i++) A2[i] = el[i * 32 + pj];
#pragma unroll
            for (int i = 0; i < 3; i++) { b2[i] = el[(9 + i) * 32 + pj]; e2[i] = el[(18 + i) * 32 + pj]; }
#pragma unroll
            for (int i = 0; i < 6; i++) { C2[i] = el[(12 + i) * 32 + pj]; J2[i] = el[(21 + i) * 32 + pj]; }
            // symmetric 3x3 stored as (00,01,02,11,12,22)
#define SYM(M_, i_, j_) M_[(i_) <= (j_) ? ((i_) == 0 ? (j_) : (i_) == 1 ? 2 + (j_) : 5) : ((j_) == 0 ? (i_) : (j_) == 1 ? 2 + (i_) : 5)]
            double T[9];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++)
                    T[3 * i + j] = (i == j ? 1.0 : 0.0) + SYM(C, i, 0) * SYM(J2, 0, j) + SYM(C, i, 1) * SYM(J2, 1, j) + SYM(C, i, 2) * SYM(J2, 2, j);
            // M = T^-1 (adjugate)
            const double c00 = T[4] * T[8] - T[5] * T[7], c01 = T[5] * T[6] - T[3] * T[8], c02 = T[3] * T[7] - T[4] * T[6];
            const double idt = fast_rcp(T[0] * c00 + T[1] * c01 + T[2] * c02);
            double M[9];
            M[0] = c00 * idt; M[3] = c01 * idt; M[6] = c02 * idt;
            M[1] = (T[2] * T[7] - T[1] * T[8]) * idt; M[4] = (T[0] * T[8] - T[2] * T[6]) * idt; M[7] = (T[1] * T[6] - T[0] * T[7]) * idt;
            M[2] = (T[1] * T[5] - T[2] * T[4]) * idt; M[5] = (T[2] * T[3] - T[0] * T[5]) * idt; M[8] = (T[0] * T[4] - T[1] * T[3]) * idt;
            double MA[9], An[9], MC[9], X[9], MtJ[9], Y[9];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    MA[3 * i + j] = M[3 * i] * A[j] + M[3 * i + 1] * A[3 + j] + M[3 * i + 2] * A[6 + j];
                    MC[3 * i + j] = M[3 * i] * SYM(C, 0, j) + M[3 * i + 1] * SYM(C, 1, j) + M[3 * i + 2] * SYM(C, 2, j);
                    MtJ[3 * i + j] = M[i] * SYM(J2, 0, j) + M[3 + i] * SYM(J2, 1, j) + M[6 + i] * SYM(J2, 2, j);
                }
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    An[3 * i + j] = A2[3 * i] * MA[j] + A2[3 * i + 1] * MA[3 + j] + A2[3 * i + 2] * MA[6 + j];
                    X[3 * i + j] = A2[3 * i] * MC[j] + A2[3 * i + 1] * MC[3 + j] + A2[3 * i + 2] * MC[6 + j];
                    Y[3 * i + j] = MtJ[3 * i] * A[j] + MtJ[3 * i + 1] * A[3 + j] + MtJ[3 * i + 2] * A[6 + j];
                }
            // vectors
            double v[3], w[3], Mv[3], Mw[3];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                v[i] = b[i] + SYM(C, i, 0) * e2[0] + SYM(C, i, 1) * e2[1] + SYM(C, i, 2) * e2[2];
                w[i] = e2[i] - (SYM(J2, i, 0) * b[0] + SYM(J2, i, 1) * b[1] + SYM(J2, i, 2) * b[2]);
            }
#pragma unroll
            for (int i = 0; i < 3; i++) {
                Mv[i] = M[3 * i] * v[0] + M[3 * i + 1] * v[1] + M[3 * i + 2] * v[2];
                Mw[i] = M[i] * w[0] + M[3 + i] * w[1] + M[6 + i] * w[2];
            }
            double bn[3], en[3], Cn[6], Jn[6];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                bn[i] = A2[3 * i] * Mv[0] + A2[3 * i + 1] * Mv[1] + A2[3 * i + 2] * Mv[2] + b2[i];
                en[i] = A[i] * Mw[0] + A[3 + i] * Mw[1] + A[6 + i] * Mw[2] + e[i];
            }
            {
                int t = 0;
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int j = i; j < 3; j++, t++) {
                        Cn[t] = X[3 * i] * A2[3 * j] + X[3 * i + 1] * A2[3 * j + 1] + X[3 * i + 2] * A2[3 * j + 2] + C2[t];
                        Jn[t] = A[i] * Y[j] + A[3 + i] * Y[3 + j] + A[6 + i] * Y[6 + j] + Jm[t];
                    }
            }
#undef SYM
#pragma unroll
            for (int i = 0; i < 9; i++) A[i] = An[i];
#pragma unroll
            for (int i = 0; i < 3; i++) { b[i] = bn[i]; e[i] = en[i]; }
#pragma unroll
            for (int i = 0; i < 6; i++) { C[i] = Cn[i]; Jm[i] = Jn[i]; }
        }
    }
    // ---- (P_k, p_k) = (J, -eta) of the range k..N; every stage takes its successor's ----
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 6; i++) el[i * 32 + lane] = Jm[i];
#pragma unroll
    for (int i = 0; i < 3; i++) el[(6 + i) * 32 + lane] = -e[i];
    __syncwarp();
    int ok = 1;
    // closed-loop map of the stage, y_{k+1} = Phi y_k + phi (identity for the terminal stage and idle lanes), and its gains
    double Ph[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, ph[3] = {0, 0, 0}, Kr[6] = {0, 0, 0, 0, 0, 0}, kr[2] = {0, 0};
    if (act) {
        double *g = rec + (size_t)k * KKT_REC + KR_OUT;
        if (!dyn) {
#pragma unroll
            for (int i = 0; i < 8; i++) g[i] = 0.0;
#pragma unroll
            for (int i = 0; i < 6; i++) g[8 + i] = q[KR_H + i];
            g[14] = q[KR_h + 0]; g[15] = q[KR_h + 1]; g[16] = q[KR_h + 2];
        } else {
            const int n = lane + 1;
            const double q00 = el[0 * 32 + n], q01 = el[1 * 32 + n], q02 = el[2 * 32 + n], q11 = el[3 * 32 + n], q12 = el[4 * 32 + n],
                         q22 = el[5 * 32 + n], v0 = el[6 * 32 + n], v1 = el[7 * 32 + n], v2 = el[8 * 32 + n];
            const double a = q[KR_G + 0], bb = q[KR_G + 1], b11 = q[KR_G + 2], b12 = q[KR_G + 3], b21 = q[KR_G + 4], b22 = q[KR_G + 5];
            const double d0 = q[KR_D + 0], d1 = q[KR_D + 1], d2 = q[KR_D + 2];
            const double w0 = q00 * d0 + q01 * d1 + q02 * d2 + v0;
            const double w1 = q01 * d0 + q11 * d1 + q12 * d2 + v1;
            const double w2 = q02 * d0 + q12 * d1 + q22 * d2 + v2;
            const double t0 = q02 + a * q00 + bb * q01, t1 = q12 + a * q01 + bb * q11, t2 = q22 + a * q02 + bb * q12;
            const double x00 = q[KR_H + 0] + q00, x01 = q[KR_H + 1] + q01, x11 = q[KR_H + 3] + q11;
            const double x02 = t0, x12 = t1, x22 = q[KR_H + 5] + t2 + a * t0 + bb * t1;
            const double e0 = b11 * q00 + b21 * q01, e1 = b11 * q01 + b21 * q11;
            const double f0 = b12 * q00 + b22 * q01 + dt * q02, f1 = b12 * q01 + b22 * q11 + dt * q12, f2 = b12 * q02 + b22 * q12 + dt * q22;
            const double u00 = e0, u01 = e1, u02 = q[KR_H + 10] + b11 * t0 + b21 * t1;
            const double u10 = f0, u11 = f1, u12 = q[KR_H + 11] + b12 * t0 + b22 * t1 + dt * t2;
            const double r00 = q[KR_H + 12] + b11 * e0 + b21 * e1;
            const double r01 = q[KR_H + 13] + b11 * f0 + b21 * f1;
            const double r11 = q[KR_H + 14] + b12 * f0 + b22 * f1 + dt * f2;
            const double gx0 = q[KR_h + 0] + w0, gx1 = q[KR_h + 1] + w1, gx2 = q[KR_h + 2] + a * w0 + bb * w1 + w2;
            const double gu0 = q[KR_h + 3] + b11 * w0 + b21 * w1, gu1 = q[KR_h + 4] + b12 * w0 + b22 * w1 + dt * w2;
            const double det = r00 * r11 - r01 * r01;
            if (!(r00 > 0.0) || !(det > 0.0)) ok = 0;
            const double idet = fast_rcp(det);
            const double i00 = r11 * idet, i01 = -r01 * idet, i11 = r00 * idet;
            const double K00 = -(i00 * u00 + i01 * u10), K01 = -(i00 * u01 + i01 * u11), K02 = -(i00 * u02 + i01 * u12);
            const double K10 = -(i01 * u00 + i11 * u10), K11 = -(i01 * u01 + i11 * u11), K12 = -(i01 * u02 + i11 * u12);
            const double k0 = -(i00 * gu0 + i01 * gu1), k1 = -(i01 * gu0 + i11 * gu1);
            g[0] = K00; g[1] = K01; g[2] = K02; g[3] = K10; g[4] = K11; g[5] = K12; g[6] = k0; g[7] = k1;
            Kr[0] = K00; Kr[1] = K01; Kr[2] = K02; Kr[3] = K10; Kr[4] = K11; Kr[5] = K12; kr[0] = k0; kr[1] = k1;
            Ph[0] = 1.0 + b11 * K00 + b12 * K10; Ph[1] = b11 * K01 + b12 * K11; Ph[2] = a + b11 * K02 + b12 * K12;
            Ph[3] = b21 * K00 + b22 * K10; Ph[4] = 1.0 + b21 * K01 + b22 * K11; Ph[5] = bb + b21 * K02 + b22 * K12;
            Ph[6] = dt * K10; Ph[7] = dt * K11; Ph[8] = 1.0 + dt * K12;
            ph[0] = b11 * k0 + b12 * k1 + d0; ph[1] = b21 * k0 + b22 * k1 + d1; ph[2] = dt * k1 + d2;
            g[8] = x00 + u00 * K00 + u10 * K10;
            g[9] = x01 + 0.5 * ((u00 * K01 + u10 * K11) + (u01 * K00 + u11 * K10));
            g[10] = x02 + 0.5 * ((u00 * K02 + u10 * K12) + (u02 * K00 + u12 * K10));
            g[11] = x11 + u01 * K01 + u11 * K11;
            g[12] = x12 + 0.5 * ((u01 * K02 + u11 * K12) + (u02 * K01 + u12 * K11));
            g[13] = x22 + u02 * K02 + u12 * K12;
            g[14] = gx0 + u00 * k0 + u10 * k1;
            g[15] = gx1 + u01 * k0 + u11 * k1;
            g[16] = gx2 + u02 * k0 + u12 * k1;
        }
    }
    // ---- forward roll-out as an inclusive prefix scan of the affine maps: after it lane k maps y_0 to y_{k+1} ----
#pragma unroll 1
    for (int lvl = 1; lvl <= N; lvl <<= 1) {
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 9; i++) el[i * 32 + lane] = Ph[i];
#pragma unroll
        for (int i = 0; i < 3; i++) el[(9 + i) * 32 + lane] = ph[i];
        __syncwarp();
        const int pj = lane - lvl;
        if (pj >= 0) {
            double P1[9], p1[3], Pn[9], pn[3];
#pragma unroll
            for (int i = 0; i < 9; i++) P1[i] = el[i * 32 + pj];
#pragma unroll
            for (int i = 0; i < 3; i++) p1[i] = el[(9 + i) * 32 + pj];
#pragma unroll
            for (int i = 0; i < 3; i++) {
#pragma unroll
                for (int j = 0; j < 3; j++) Pn[3 * i + j] = Ph[3 * i] * P1[j] + Ph[3 * i + 1] * P1[3 + j] + Ph[3 * i + 2] * P1[6 + j];
                pn[i] = Ph[3 * i] * p1[0] + Ph[3 * i + 1] * p1[1] + Ph[3 * i + 2] * p1[2] + ph[i];
            }
#pragma unroll
            for (int i = 0; i < 9; i++) Ph[i] = Pn[i];
#pragma unroll
            for (int i = 0; i < 3; i++) ph[i] = pn[i];
        }
    }
    // y_0 = 0, so y_{k+1} = phi of lane k: every stage takes its predecessor's, then du_k = kf_k + K_k y_k
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 3; i++) el[i * 32 + lane] = ph[i];
    __syncwarp();
    if (act) {
        double y0 = 0, y1 = 0, y2 = 0;
        if (lane >= 1) { y0 = el[0 * 32 + lane - 1]; y1 = el[1 * 32 + lane - 1]; y2 = el[2 * 32 + lane - 1]; }
        // step entries of the stage -> rows 3-7 of the scan scratch (read back by kkt_solve_lp)
        el[3 * 32 + lane] = y0; el[4 * 32 + lane] = y1; el[5 * 32 + lane] = y2;
        el[6 * 32 + lane] = kr[0] + Kr[0] * y0 + Kr[1] * y1 + Kr[2] * y2;
        el[7 * 32 + lane] = kr[1] + Kr[3] * y0 + Kr[4] * y1 + Kr[5] * y2;
    }
    __syncwarp();
    return __all_sync(FULL, ok);
}

// Serial backward recursion of the lane-parallel form (steps A-E above): one stage per step, one matrix entry per lane.
// Returns false when a condensed Muu block is not positive definite (wrong inertia).
__device__ __forceinline__ bool kkt_backward_serial(double *rec, int N, const KktRoles &R) {
    // ---- backward recursion: one stage per step, one matrix entry per lane ----
    const int a0 = R.a_off & 0xff, a1 = (R.a_off >> 8) & 0xff, a2 = (R.a_off >> 16) & 0xff;
    const int sp0 = R.a_src & 0xff, sp1 = (R.a_src >> 8) & 0xff, sp2 = (R.a_src >> 16) & 0xff, spv = R.a_src >> 24;
    const int bs0 = R.b_src & 0xff, bs1 = (R.b_src >> 8) & 0xff, bs2 = (R.b_src >> 16) & 0xff, bh = R.b_src >> 24;
    const int b0 = R.b_off & 0xff, b1 = (R.b_off >> 8) & 0xff, b2 = (R.b_off >> 16) & 0xff, outo = R.b_off >> 24;
    const int eu0 = R.e_src & 0xff, eu1 = (R.e_src >> 8) & 0xff, eu2 = (R.e_src >> 16) & 0xff, eu3 = R.e_src >> 24;
    const int em = R.e_misc & 0xff, role = (R.e_misc >> 8) & 0xff;
    const bool addp = (R.e_misc >> 16) & 1;
    // role flags for the branch-free step E (a divergent if/else chain here would serialise seven paths)
    const bool isP = role <= 1, isp = role == 2, iskf = role >= 5, row0 = (role == 3 || role == 5);
    double res = 0.0; // this lane's step-E result of the previous stage (P = 0, p = 0 enter the terminal stage)
    // operands that do not depend on the recursion are fetched one stage ahead
    const double *q = rec + (size_t)N * KKT_REC;
    double ga0 = q[a0], ga1 = q[a1], ga2 = q[a2], gb0 = q[b0], gb1 = q[b1], gb2 = q[b2], hb = q[bh];
#pragma unroll 1
    for (int k = N; k >= 0; --k) {
        const double ca0 = ga0, ca1 = ga1, ca2 = ga2, cb0 = gb0, cb1 = gb1, cb2 = gb2, ch = hb;
        if (k > 0) {
            const double *qn = rec + (size_t)(k - 1) * KKT_REC;
            ga0 = qn[a0]; ga1 = qn[a1]; ga2 = qn[a2]; gb0 = qn[b0]; gb1 = qn[b1]; gb2 = qn[b2]; hb = qn[bh];
        }
        // step A
        const double p0 = __shfl_sync(FULL, res, sp0), p1 = __shfl_sync(FULL, res, sp1), p2 = __shfl_sync(FULL, res, sp2);
        const double pv = __shfl_sync(FULL, res, spv);
        const double tA = (p0 * ca0 + p1 * ca1) + (p2 * ca2 + (addp ? pv : 0.0)); // (tree form: shorter dependent chain)
        // step B
        const double t0 = __shfl_sync(FULL, tA, bs0), t1 = __shfl_sync(FULL, tA, bs1), t2 = __shfl_sync(FULL, tA, bs2);
        const double mB = (ch + cb0 * t0) + (cb1 * t1 + cb2 * t2);
        // step C: Muu and mu to every lane
        const double r00 = __shfl_sync(FULL, mB, 12), r01 = __shfl_sync(FULL, mB, 13), r11 = __shfl_sync(FULL, mB, 14);
        const double m3 = __shfl_sync(FULL, mB, 18), m4 = __shfl_sync(FULL, mB, 19);
        const double ua0 = __shfl_sync(FULL, mB, eu0), ua1 = __shfl_sync(FULL, mB, eu1);
        const double ub0 = __shfl_sync(FULL, mB, eu2), ub1 = __shfl_sync(FULL, mB, eu3);
        const double ma = __shfl_sync(FULL, mB, em);
        const double det = r00 * r11 - r01 * r01;
        if (!(r00 > 0.0) || !(det > 0.0)) return false; // warp-uniform: every lane holds the same Muu
        const double idet = fast_rcp(det);
        // step E, one formula for every role:  r = base + idet * ((x.y) + (z.w)) / 2  with the adjugate -adj(Muu) in
        // place of -Muu^-1, so that the products run while the reciprocal of det is still in flight:
        //   P'[a][b] = M[a][b] + (Mxu[a].K[.][b] + Mxu[b].K[.][a]) / 2      p'[a] = m[a] + Mxu[a].kf
        //   K[c][a]  = -Muu^-1[c].Mxu[a]                                     kf[c] = -Muu^-1[c].mu
        const double n00 = -r11, n01 = r01, n11 = -r00; // -adj(Muu)
        const double K0a = n00 * ua0 + n01 * ua1, K1a = n01 * ua0 + n11 * ua1;
        const double K0b = n00 * ub0 + n01 * ub1, K1b = n01 * ub0 + n11 * ub1;
        const double kf0 = n00 * m3 + n01 * m4, kf1 = n01 * m3 + n11 * m4;
        const double base = isP ? mB : (isp ? ma : 0.0);
        const double x0 = iskf ? m3 : ua0, x1 = iskf ? m4 : ua1;
        const double y0 = isP ? K0b : (isp ? kf0 : (row0 ? n00 : n01));
        const double y1 = isP ? K1b : (isp ? kf1 : (row0 ? n01 : n11));
        const double z0 = isP ? ub0 : x0, z1 = isP ? ub1 : x1;
        const double w0 = isP ? K0a : y0, w1 = isP ? K1a : y1;
        const double r = base + (0.5 * idet) * ((x0 * y0 + x1 * y1) + (z0 * w0 + z1 * w1));
        res = r;
        if (outo != 0xff) rec[(size_t)k * KKT_REC + KR_OUT + outo] = r;
    }
    return true;
}
// out-of-line copy for the scan instance (its fallback when the scan does not apply)
__device__ __noinline__ bool kkt_backward_serial_ool(double *rec, int N, const KktRoles &R) { return kkt_backward_serial(rec, N, R); }

template <int J, bool SCAN>
__device__ __forceinline__ bool kkt_solve_lp(const KParams &P, Stg (&s)[J], const double (&rc)[J][3],
                                             const double (&rd)[J][2], bool useW, double dw, Step (&o)[J], int lane,
                                             double *rec, const KktRoles &R) {
    const int N = P.N;
    const double dt = P.dt;
    // ---- stage records (stage-parallel) ----
    __syncwarp();
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int k = lane * J + j;
        if (k > N) continue;
        const Stg &t = s[j];
        double *q = rec + (size_t)k * KKT_REC;
        const bool dyn = k < N;
        q[KR_ZERO] = 0.0; q[KR_ONE] = 1.0; q[KR_DT] = dt;
        q[KR_G + 0] = dyn ? t.a13 : 0.0; q[KR_G + 1] = dyn ? t.a23 : 0.0;
        q[KR_G + 2] = dyn ? t.b11 : 0.0; q[KR_G + 3] = dyn ? t.b12 : 0.0;
        q[KR_G + 4] = dyn ? t.b21 : 0.0; q[KR_G + 5] = dyn ? t.b22 : 0.0;
                q[KR_H + 0] = (useW ? t.hxx : 0.0) + dw; q[KR_H + 1] = (useW ? t.hxy : 0.0); q[KR_H + 2] = 0.0;
        q[KR_H + 3] = (useW ? t.hyy : 0.0) + dw; q[KR_H + 4] = 0.0; q[KR_H + 5] = (useW ? t.htt : 0.0) + dw;
        q[KR_H + 6] = 0.0; q[KR_H + 7] = 0.0; q[KR_H + 8] = 0.0; q[KR_H + 9] = 0.0;
        q[KR_H + 10] = dyn ? (useW ? t.htv : 0.0) : 0.0; q[KR_H + 11] = dyn ? (useW ? t.htw : 0.0) : 0.0;
        q[KR_H + 12] = dyn ? ((useW ? t.hvv : 0.0) + dw + t.Dsig[0]) : 1.0;
        q[KR_H + 13] = dyn ? (useW ? t.hvw : 0.0) : 0.0;
        q[KR_H + 14] = dyn ? ((useW ? t.hww : 0.0) + dw + t.Dsig[1]) : 1.0;
        const bool hasx = k >= 1;
        q[KR_h + 0] = hasx ? t.rx[0] : 0.0; q[KR_h + 1] = hasx ? t.rx[1] : 0.0; q[KR_h + 2] = hasx ? t.rx[2] : 0.0;
        q[KR_h + 3] = dyn ? (t.ru[0] + t.Dsig[0] * rd[j][0] + t.rs[0]) : 0.0;
        q[KR_h + 4] = dyn ? (t.ru[1] + t.Dsig[1] * rd[j][1] + t.rs[1]) : 0.0;
        q[KR_D + 0] = dyn ? -rc[j][0] : 0.0; q[KR_D + 1] = dyn ? -rc[j][1] : 0.0; q[KR_D + 2] = dyn ? -rc[j][2] : 0.0;
    }
    __syncwarp();
    bool scanned = false;
    if (SCAN && J == 1) {
        // parallel-in-time form: all cost-to-go matrices from a suffix scan over the lanes (kkt_backward_scan);
        // 1 = done, 0 = wrong inertia, -1 = a local control block is not safely positive definite (serial form instead)
        const int r = kkt_backward_scan(rec, rec + (size_t)(N + 1) * (KKT_REC + 6), N, dt, lane);
        if (r == 0) return false;
        scanned = r > 0;
        if (!scanned && !kkt_backward_serial_ool(rec, N, R)) return false;
    } else {
        if (!kkt_backward_serial(rec, N, R)) return false;
    }
    __syncwarp();
    if (scanned) {
        // (the scan routine has rolled the step out as a prefix scan of the closed-loop maps)
        const int k = lane;
        if (k <= N) {
            const double *el = rec + (size_t)(N + 1) * (KKT_REC + 6);
            o[0].dX[0] = el[3 * 32 + lane]; o[0].dX[1] = el[4 * 32 + lane]; o[0].dX[2] = el[5 * 32 + lane];
            if (k < N) { o[0].dU[0] = el[6 * 32 + lane]; o[0].dU[1] = el[7 * 32 + lane]; }
        }
    } else
    // ---- forward roll-out: every lane computes the whole step, keeps the entries of its own stages ----
    {
        double y0 = 0, y1 = 0, y2 = 0;
        for (int k = 0; k <= N; ++k) {
            const double *q = rec + (size_t)k * KKT_REC;
#pragma unroll
            for (int j = 0; j < J; ++j)
                if (k == lane * J + j) { o[j].dX[0] = y0; o[j].dX[1] = y1; o[j].dX[2] = y2; }
            if (k < N) {
                const double *g = q + KR_OUT;
                const double du0 = g[6] + g[0] * y0 + g[1] * y1 + g[2] * y2;
                const double du1 = g[7] + g[3] * y0 + g[4] * y1 + g[5] * y2;
#pragma unroll
                for (int j = 0; j < J; ++j)
                    if (k == lane * J + j) { o[j].dU[0] = du0; o[j].dU[1] = du1; }
                const double n0 = y0 + q[KR_G + 0] * y2 + q[KR_G + 2] * du0 + q[KR_G + 3] * du1 + q[KR_D + 0];
                const double n1 = y1 + q[KR_G + 1] * y2 + q[KR_G + 4] * du0 + q[KR_G + 5] * du1 + q[KR_D + 1];
                const double n2 = y2 + dt * du1 + q[KR_D + 2];
                y0 = n0; y1 = n1; y2 = n2;
            }
        }
    }
    // ---- multiplier and slack steps (stage-parallel) ----
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int k = lane * J + j;
        Stg &t = s[j];
        const bool act = (k <= N), dyn = (k < N);
        const double *g = rec + (size_t)(act ? k : 0) * KKT_REC + KR_OUT;
        const double a0_ = o[j].dX[0], a1_ = o[j].dX[1], a2_ = o[j].dX[2];
        // P' pairs at g[8..13] = (0,0) (0,1) (0,2) (1,1) (1,2) (2,2), p' at g[14..16]
        const double l0 = -(g[14] + g[8] * a0_ + g[9] * a1_ + g[10] * a2_);
        const double l1 = -(g[15] + g[9] * a0_ + g[11] * a1_ + g[12] * a2_);
        const double l2 = -(g[16] + g[10] * a0_ + g[12] * a1_ + g[13] * a2_);
        const bool hasl = act && (k >= 1);
        o[j].dlam[0] = hasl ? l0 : 0.0; o[j].dlam[1] = hasl ? l1 : 0.0; o[j].dlam[2] = hasl ? l2 : 0.0;
        if (!act) { o[j].dX[0] = o[j].dX[1] = o[j].dX[2] = 0; }
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const double ds = dyn ? (o[j].dU[i] + rd[j][i]) : 0.0;
            o[j].dS[i] = ds;
            o[j].dyd[i] = dyn ? (t.Dsig[i] * ds + t.rs[i]) : 0.0;
            if (!dyn) o[j].dU[i] = 0;
        }
    }
    __syncwarp();
    return true;
}

// fraction-to-the-boundary step for the slacks
template <int J>
__device__ __forceinline__ double frac_to_bound(const KParams &P, const Stg (&s)[J], const Step (&o)[J], double tau,
                                                int lane) {
    double a = 1.0;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const bool dyn = (lane * J + j) < P.N;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const double sl = s[j].S[i] - P.sL[i], su = P.sU[i] - s[j].S[i], ds = o[j].dS[i];
            if (dyn && ds < 0) a = fmin(a, -tau * sl / ds);
            if (dyn && ds > 0) a = fmin(a, tau * su / ds);
        }
    }
    return wmin(a);
}

// trial point curr + alpha*step: theta (1-norm) and barrier objective, both warp-uniform
template <int J, bool OBS>
__device__ __forceinline__ void trial_eval(const KParams &P, const ObsList &OL, const double *sox, const double *soy, double *ocs, Stg (&s)[J],
                                           const Step (&o)[J], double alpha, double mu, double df, int lane,
                                           double &th_t, double &phi_t) {
#pragma unroll
    for (int j = 0; j < J; ++j) {
        Stg &t = s[j];
#pragma unroll
        for (int i = 0; i < 3; i++) t.Xt[i] = t.X[i] + alpha * o[j].dX[i];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            t.Ut[i] = t.U[i] + alpha * o[j].dU[i];
            t.St[i] = t.S[i] + alpha * o[j].dS[i];
        }
    }
    double th = 0, ph = 0;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int k = lane * J + j;
        const bool act = (k <= P.N), dyn = (k < P.N);
        const bool obs = OBS && P.obs_form != B200MPC_OBS_NONE && k >= P.obs_k0 && k <= P.obs_k1;
        Stg &t = s[j];
        double Xn[3];
        NEXT3(Xn, Xt, j);
        double fv = stage_value(P, OL, sox, soy, t.Xt, t.Ut, t.r, t.ub, Xn, t.ct, act, dyn, false);
        if (obs && act) {
            // value and derivatives: the accepted trial point is the next iterate, and its linearisation then takes
            // the sums from this per-stage cache in shared memory instead of walking the obstacle list again
            double ov, gx, gy, oxx, oxy, oyy;
            obstacle_sum<true>(P, OL, sox, soy, t.Xt[0], t.Xt[1], ov, gx, gy, oxx, oxy, oyy);
            double *oc = ocs + 6 * k;
            oc[0] = ov; oc[1] = gx; oc[2] = gy; oc[3] = oxx; oc[4] = oxy; oc[5] = oyy;
            fv += ov;
        }
        double bar = 0, thl = 0;
        if (dyn) {
            thl = fabs(t.ct[0]) + fabs(t.ct[1]) + fabs(t.ct[2]) + fabs(t.Ut[0] - t.St[0]) + fabs(t.Ut[1] - t.St[1]);
            // one logarithm per stage (of the product of the four slack distances; a slack outside its bounds gives
            // NaN like log(negative) would)
            const double l0 = t.St[0] - P.sL[0], l1 = P.sU[0] - t.St[0], l2 = t.St[1] - P.sL[1], l3 = P.sU[1] - t.St[1];
            const bool inside = (l0 > 0.0) && (l1 > 0.0) && (l2 > 0.0) && (l3 > 0.0);
            bar = -mu * log(inside ? (l0 * l1) * (l2 * l3) : -1.0);
        }
        th += thl;
        ph += df * fv + bar;
    }
    th_t = wsum(th);
    phi_t = wsum(ph);
}

struct LsRef {
    double phi, theta, gbd, theta_min, theta_max;
};

// FilterLSAcceptor::CheckAcceptabilityOfTrialPoint; the filter lives one entry per lane.
__device__ __noinline__ bool ls_acceptable(const LsRef &r, double alpha_test, double phi_t, double th_t,
                                           double fphi, double ftheta, bool fvalid, bool &ftype_armijo) {
    ftype_armijo = false;
    if (!isfinite(th_t) || !isfinite(phi_t)) return false;
    if (th_t > r.theta_max) return false;
    const bool ftype = (r.gbd < 0) && (alpha_test * pow_ool(-r.gbd, S_PHI) > DELTA_LS * pow_ool(r.theta, S_THETA));
    const bool armijo = cmp_le(phi_t - r.phi, ETA_PHI * alpha_test * r.gbd, r.phi);
    ftype_armijo = ftype && armijo;
    bool ok;
    if (alpha_test > 0 && ftype && r.theta <= r.theta_min) {
        ok = armijo;
    } else {
        if (phi_t > r.phi) {
            double bas = 1.0;
            if (fabs(r.phi) > 10.0) bas = log10_ool(fabs(r.phi));
            if (log10_ool(phi_t - r.phi) > OBJ_MAX_INC + bas) return false;
        }
        ok = cmp_le(th_t, (1 - GAMMA_THETA) * r.theta, r.theta) || cmp_le(phi_t - r.phi, -GAMMA_PHI * r.theta, r.phi);
    }
    if (!ok) return false;
    const bool rej = fvalid && !(cmp_le(phi_t, fphi, fphi) || cmp_le(th_t, ftheta, ftheta));
    return !__any_sync(FULL, rej);
}

__device__ __forceinline__ void filter_add(double phi, double theta, double &fphi, double &ftheta, bool &fvalid,
                                           int &ring, int lane) {
    if (fvalid && fphi >= phi && ftheta >= theta) fvalid = false; // dominated by the new entry
    const unsigned freem = __ballot_sync(FULL, !fvalid);
    int slot;
    if (freem) slot = __ffs(freem) - 1;
    else { slot = ring & 31; ring++; }
    if (lane == slot) { fphi = phi; ftheta = theta; fvalid = true; }
}

// ---- the per-problem solve ----------------------------------------------------------------------------------
#ifndef B200MPC_KKT_SERIAL
#define KKT_SOLVE(P, s, rc, rd, useW, dw, o, lane) kkt_solve_lp<J, SCAN>(P, s, rc, rd, useW, dw, o, lane, rec, roles)
#else
#define KKT_SOLVE(P, s, rc, rd, useW, dw, o, lane) kkt_solve<J>(P, s, rc, rd, useW, dw, o, lane)
#endif

// OBS: the kernel instance carries the obstacle cost (variant A; B with its gauss cost enabled).  The instance without
// it has no obstacle loops at all — they would cost the obstacle-free variants registers (spills) for nothing.
// SCAN: the instance uses the parallel-in-time form of the Riccati recursion (J == 1 only).  A separate instance, so
// that the out-of-line call does not cost the serial-form instance registers.
template <int J, bool OBS, bool SCAN>
__device__ void solve_one(const KParams &P, const BatchArgs &A, int b, int lane, double *sox, double *soy, double *rec,
                          const KktRoles &roles, const double *hrec = nullptr) {
    // hrec: record of a problem the lane kernel handed over (BatchArgs::hand_rec): the solve resumes at the top of the
    // interior-point loop with the exported iterate, barrier parameter, filter and counters instead of starting cold
    const int N = P.N;
    Stg s[J];
    Step st[J], soc[J];
    double rc[J][3], rd[J][2], csoc[J][3], dsoc[J][2];

    // ---- load the problem (coalesced: consecutive lanes read consecutive stages) ----
    const double x00 = A.x0[3 * (size_t)b], x01 = A.x0[3 * (size_t)b + 1], x02 = A.x0[3 * (size_t)b + 2];
    ObsList OL;
    OL.n_eff = 1; OL.w0 = 1.0;
    if (OBS && P.obs_form != B200MPC_OBS_NONE) {
        const double *gx = A.ox + (size_t)A.obs_stride * b, *gy = A.oy + (size_t)A.obs_stride * b;
        for (int i = lane; i < P.M; i += 32) { sox[i] = gx[i]; soy[i] = gy[i]; }
        __syncwarp();
        OL = obstacle_list_setup(sox, soy, P.M, lane);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int k = lane * J + j;
        Stg &t = s[j];
        const bool dyn = k < N;
        t.X[0] = (k == 0) ? x00 : 0.0; t.X[1] = (k == 0) ? x01 : 0.0; t.X[2] = (k == 0) ? x02 : 0.0;
        t.U[0] = t.U[1] = 0;
        t.r[0] = t.r[1] = t.r[2] = 0; t.ub[0] = t.ub[1] = 0;
        if (dyn) {
            if (A.u_init) {
                t.U[0] = A.u_init[(size_t)b * 2 * N + 2 * k];
                t.U[1] = A.u_init[(size_t)b * 2 * N + 2 * k + 1];
            }
            if (P.ref_kind == B200MPC_REF_GOAL) {
                t.r[0] = A.xref[3 * (size_t)b]; t.r[1] = A.xref[3 * (size_t)b + 1]; t.r[2] = A.xref[3 * (size_t)b + 2];
            } else {
                const double *xr = A.xref + (size_t)b * 3 * N + 3 * k;
                t.r[0] = xr[0]; t.r[1] = xr[1]; t.r[2] = xr[2];
                t.ub[0] = A.uref[(size_t)b * 2 * N + 2 * k];
                t.ub[1] = A.uref[(size_t)b * 2 * N + 2 * k + 1];
            }
        }
#pragma unroll
        for (int i = 0; i < 3; i++) t.lam[i] = 0;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            // slack initialisation: s = d(x) pushed into the interior; bound multipliers 1
            const double lo = P.sL[i], hi = P.sU[i];
            const double pl = fmin(BOUND_PUSH * fmax(1.0, fabs(lo)), BOUND_FRAC * (hi - lo));
            const double pu = fmin(BOUND_PUSH * fmax(1.0, fabs(hi)), BOUND_FRAC * (hi - lo));
            double sv = t.U[i];
            if (sv < lo + pl) sv = lo + pl;
            if (sv > hi - pu) sv = hi - pu;
            t.S[i] = dyn ? sv : 0.5 * (lo + hi);
            t.vL[i] = 1.0; t.vU[i] = 1.0; t.yd[i] = 0;
            t.Dsig[i] = 1.0; t.rs[i] = 0;
        }
        if (hrec && k <= N) {
            const double *q = hrec + 16 * k;
            if (k >= 1) { t.X[0] = q[0]; t.X[1] = q[1]; t.X[2] = q[2]; t.lam[0] = q[3]; t.lam[1] = q[4]; t.lam[2] = q[5]; }
            if (dyn) {
                t.U[0] = q[6]; t.U[1] = q[7]; t.S[0] = q[8]; t.S[1] = q[9]; t.yd[0] = q[10]; t.yd[1] = q[11];
                t.vL[0] = q[12]; t.vL[1] = q[13]; t.vU[0] = q[14]; t.vU[1] = q[15];
            }
        }
    }

    int status = B200MPC_MAXITER_EXCEEDED;
    int iter = 0, ls_extra = 0, n_resto = 0;
    double df = 1.0;
    double fcur = 0;
    if (hrec) {
        const double *q = hrec + HAND_SCAL(N);
        iter = (int)q[HS_ITER]; ls_extra = (int)q[HS_LS]; n_resto = (int)q[HS_NRESTO];
        df = q[HS_DF];
    }

    bool oc_valid = false; // ocs holds the obstacle sums of the current iterate (it was the accepted trial point)
    double *ocs = rec + (size_t)(N + 1) * KKT_REC; // per stage: value, gradient (2), Hessian (3) of the obstacle sum
    // full evaluation of the current point; returns the objective value (unscaled, warp-uniform)
    auto eval_point = [&](double dfv) -> double {
        double Xn0[3], ln0[3];
        double fl = 0;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int k = lane * J + j;
            const bool act = (k <= N), dyn = (k < N);
            const bool obs = OBS && P.obs_form != B200MPC_OBS_NONE && k >= P.obs_k0 && k <= P.obs_k1;
            NEXT3(Xn0, X, j);
            NEXT3(ln0, lam, j);
            fl += stage_full(P, OL, sox, soy, s[j], Xn0, ln0, dfv, act, dyn, obs, (oc_valid && k <= N) ? ocs + 6 * k : nullptr);
        }
        oc_valid = false;
        return wsum(fl);
    };

    // ---- objective scaling from the gradient at the starting point; invalid-number check ----
    if (!hrec) {
        fcur = eval_point(1.0);
        double gmax = 0;
        int bad = 0;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int k = lane * J + j;
#pragma unroll
            for (int i = 0; i < 5; i++) {
                const bool use = (i < 3) ? (k >= 1 && k <= N) : (k < N);
                if (use) {
                    if (!isfinite(s[j].g[i])) bad = 1;
                    gmax = fmax(gmax, fabs(s[j].g[i]));
                }
            }
#pragma unroll
            for (int i = 0; i < 3; i++)
                if (k < N && !isfinite(s[j].c[i])) bad = 1;
        }
        gmax = wmax(gmax);
        bad = wor(bad);
        if (bad || !isfinite(fcur)) {
            status = B200MPC_INVALID_NUMBER_DETECTED;
            goto finish;
        }
        if (gmax > 100.0) df = fmax(100.0 / gmax, 1e-8);
    }

    // ---- least-squares multiplier estimate ----
    if (!hrec) {
        fcur = eval_point(df);
        double ymax = 0;
#pragma unroll
        for (int j = 0; j < J; ++j) {
#pragma unroll
            for (int i = 0; i < 3; i++) { s[j].rx[i] = s[j].g[i]; rc[j][i] = 0; }
#pragma unroll
            for (int i = 0; i < 2; i++) {
                s[j].ru[i] = s[j].g[3 + i];
                s[j].rs[i] = -s[j].vL[i] + s[j].vU[i];
                s[j].Dsig[i] = 1.0;
                rd[j][i] = 0;
            }
        }
        const bool ok = KKT_SOLVE(P, s, rc, rd, false, 1.0, st, lane);
#pragma unroll
        for (int j = 0; j < J; ++j) {
#pragma unroll
            for (int i = 0; i < 3; i++) ymax = fmax(ymax, fabs(st[j].dlam[i]));
#pragma unroll
            for (int i = 0; i < 2; i++) ymax = fmax(ymax, fabs(st[j].dyd[i]));
        }
        ymax = wmax(ymax);
        const bool keep = ok && (ymax <= 1e3);
#pragma unroll
        for (int j = 0; j < J; ++j) {
#pragma unroll
            for (int i = 0; i < 3; i++) s[j].lam[i] = keep ? st[j].dlam[i] : 0.0;
#pragma unroll
            for (int i = 0; i < 2; i++) s[j].yd[i] = keep ? st[j].dyd[i] : 0.0;
        }
    }

    {
        double mu = P.mu_init, tau = fmax(TAU_MIN, 1.0 - mu);
        double theta_max = -1, theta_min = -1, dw_last = 0;
        int acceptable_count = 0;
        bool tiny_last = false, tiny_flag = false; // BacktrackingLineSearch::tiny_step_last_iteration_, IpoptData::tiny_step_flag
        // filter: one entry per lane
        double fphi = 0, ftheta = 0;
        bool fvalid = false;
        int ring = 0;
        double theta0 = -1; // max(1, theta at the starting point): theta_max / theta_min are 1e4 / 1e-4 times it
        if (hrec) {
            const double *q = hrec + HAND_SCAL(N);
            mu = q[HS_MU]; tau = fmax(TAU_MIN, 1.0 - mu);
            theta0 = q[HS_THETA0];
            theta_max = 1e4 * q[HS_THETA0]; theta_min = 1e-4 * q[HS_THETA0];
            dw_last = q[HS_DWLAST];
            acceptable_count = (int)q[HS_ACCEPT];
            tiny_last = q[HS_TINYLAST] != 0.0; tiny_flag = q[HS_TINYFLAG] != 0.0;
            const unsigned fm = (unsigned)q[HS_FMASK];
            ring = (int)q[HS_RING];
            fvalid = (fm >> lane) & 1u;
            fphi = hrec[HAND_FILT(N) + 2 * lane]; ftheta = hrec[HAND_FILT(N) + 2 * lane + 1];
        }

        for (;;) {
            // ---- iteration-bounded launch: a problem that has taken hand_iter iterations leaves with its complete state
            //      (the record the lane kernel writes, tpp_hand_trip) and is resumed by the next launch ----
            if (A.hand_rec && iter >= A.hand_iter) {
                unsigned slot = 0;
                if (lane == 0) slot = atomicAdd(A.hand_count, 1u);
                slot = __shfl_sync(FULL, slot, 0);
                if (slot < (unsigned)A.hand_cap) {
                    double *hr = A.hand_rec + (size_t)slot * HAND_REC(N);
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const int k = lane * J + j;
                        if (k <= N) {
                            double *q = hr + 16 * k;
                            q[0] = s[j].X[0]; q[1] = s[j].X[1]; q[2] = s[j].X[2];
                            q[3] = s[j].lam[0]; q[4] = s[j].lam[1]; q[5] = s[j].lam[2];
                            q[6] = s[j].U[0]; q[7] = s[j].U[1]; q[8] = s[j].S[0]; q[9] = s[j].S[1];
                            q[10] = s[j].yd[0]; q[11] = s[j].yd[1]; q[12] = s[j].vL[0]; q[13] = s[j].vL[1];
                            q[14] = s[j].vU[0]; q[15] = s[j].vU[1];
                        }
                    }
                    hr[HAND_FILT(N) + 2 * lane] = fphi; hr[HAND_FILT(N) + 2 * lane + 1] = ftheta;
                    const unsigned fm = __ballot_sync(FULL, fvalid);
                    if (lane == 0) {
                        double *q = hr + HAND_SCAL(N);
                        q[HS_B] = (double)b; q[HS_MU] = mu; q[HS_DF] = df; q[HS_THETA0] = theta0; q[HS_DWLAST] = dw_last;
                        q[HS_ITER] = (double)iter; q[HS_LS] = (double)ls_extra; q[HS_NRESTO] = (double)n_resto;
                        q[HS_ACCEPT] = (double)acceptable_count; q[HS_TINYLAST] = tiny_last ? 1.0 : 0.0;
                        q[HS_TINYFLAG] = tiny_flag ? 1.0 : 0.0; q[HS_FMASK] = (double)fm; q[HS_RING] = (double)ring;
                    }
                    __syncwarp();
                    return;
                }
            }
            // ---- evaluate the current point ----
            fcur = eval_point(df);
            double theta, prim_inf, dual_inf, sum_y, sum_z, compl0;
            {
                double th = 0, pi = 0, di = 0, sy = 0, sz = 0, c0 = 0;
                double ln0[3];
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int k = lane * J + j;
                    const bool act = (k <= N), dyn = (k < N);
                    Stg &t = s[j];
                    NEXT3(ln0, lam, j);
                    if (!dyn) { ln0[0] = ln0[1] = ln0[2] = 0; }
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        rc[j][i] = t.c[i];
                        th += fabs(t.c[i]);
                        pi = fmax(pi, fabs(t.c[i]));
                    }
                    // grad_lag_x
                    const bool hasx = act && k >= 1;
                    double r0 = t.g[0] + t.lam[0] - ln0[0];
                    double r1 = t.g[1] + t.lam[1] - ln0[1];
                    double r2 = t.g[2] + t.lam[2] - (t.a13 * ln0[0] + t.a23 * ln0[1] + ln0[2]);
                    t.rx[0] = hasx ? r0 : 0.0; t.rx[1] = hasx ? r1 : 0.0; t.rx[2] = hasx ? r2 : 0.0;
                    if (hasx) {
                        di = fmax(di, fmax(fabs(r0), fmax(fabs(r1), fabs(r2))));
                        sy += fabs(t.lam[0]) + fabs(t.lam[1]) + fabs(t.lam[2]);
                    }
                    double q0 = t.g[3] - (t.b11 * ln0[0] + t.b21 * ln0[1]) + t.yd[0];
                    double q1 = t.g[4] - (t.b12 * ln0[0] + t.b22 * ln0[1] + P.dt * ln0[2]) + t.yd[1];
                    t.ru[0] = dyn ? q0 : 0.0; t.ru[1] = dyn ? q1 : 0.0;
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        rd[j][i] = dyn ? (t.U[i] - t.S[i]) : 0.0;
                        if (dyn) {
                            th += fabs(rd[j][i]);
                            pi = fmax(pi, fabs(rd[j][i]));
                            di = fmax(di, fabs(t.ru[i]));
                            const double gs = -t.yd[i] - t.vL[i] + t.vU[i];
                            di = fmax(di, fabs(gs));
                            sy += fabs(t.yd[i]);
                            sz += fabs(t.vL[i]) + fabs(t.vU[i]);
                            const double sl = t.S[i] - P.sL[i], su = P.sU[i] - t.S[i];
                            c0 = fmax(c0, fmax(fabs(sl * t.vL[i]), fabs(su * t.vU[i])));
                        }
                    }
                }
                theta = wsum(th); prim_inf = wmax(pi); dual_inf = wmax(di);
                sum_y = wsum(sy); sum_z = wsum(sz); compl0 = wmax(c0);
            }
            if (theta_max < 0) {
                theta0 = fmax(1.0, theta);
                theta_max = 1e4 * theta0;
                theta_min = 1e-4 * theta0;
            }
            const double sd = fmax(S_MAX, (sum_y + sum_z) / (double)(9 * N)) / S_MAX;
            const double sc = fmax(S_MAX, sum_z / (double)(4 * N)) / S_MAX;
            const double E0 = fmax(dual_inf / sd, fmax(prim_inf, compl0 / sc));
            if (!isfinite(E0)) { status = B200MPC_INVALID_NUMBER_DETECTED; break; }

            // ---- convergence ----
            if (E0 <= P.tol && dual_inf / df <= 1.0 && prim_inf <= 1e-4 && compl0 / df <= 1e-4) {
                status = B200MPC_SOLVE_SUCCEEDED;
                break;
            }
            if (P.acceptable_iter > 0 && E0 <= P.acceptable_tol && dual_inf / df <= 1e10 && prim_inf <= 1e-2 &&
                compl0 / df <= 1e-2) {
                acceptable_count++;
                if (acceptable_count >= P.acceptable_iter) { status = B200MPC_SOLVED_TO_ACCEPTABLE_LEVEL; break; }
            } else {
                acceptable_count = 0;
            }
            if (iter >= P.max_iter) { status = B200MPC_MAXITER_EXCEEDED; break; }

            // ---- barrier parameter update (MonotoneMuUpdate).  A tiny step in two consecutive iterations forces a
            // decrease; when mu cannot decrease any more: Search_Direction_Becomes_Too_Small ----
            {
                bool tflag = tiny_flag, stop = false;
                tiny_flag = false;
                for (;;) {
                    double cm = 0;
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const bool dyn = (lane * J + j) < N;
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            const double sl = s[j].S[i] - P.sL[i], su = P.sU[i] - s[j].S[i];
                            if (dyn) cm = fmax(cm, fmax(fabs(sl * s[j].vL[i] - mu), fabs(su * s[j].vU[i] - mu)));
                        }
                    }
                    cm = wmax(cm);
                    const double Emu = fmax(dual_inf / sd, fmax(prim_inf, cm / sc));
                    if (!(Emu <= K_EPS * mu) && !tflag) break;
                    const double nm = fmax(fmin(K_MU * mu, pow_ool(mu, TH_MU)), P.mu_floor);
                    if (nm == mu) { stop = tflag; break; }
                    mu = nm;
                    tau = fmax(TAU_MIN, 1.0 - mu);
                    fvalid = false;
                    tflag = false;
                }
                if (stop) { status = B200MPC_SEARCH_DIRECTION_TOO_SMALL; break; }
            }

            // ---- search direction with inertia correction ----
            double dw = 0.0;
            bool solved = false;
            for (;;) {
#pragma unroll
                for (int j = 0; j < J; ++j) {
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        const double sl = s[j].S[i] - P.sL[i], su = P.sU[i] - s[j].S[i];
                        s[j].rs[i] = -s[j].yd[i] - mu / sl + mu / su;
                        s[j].Dsig[i] = s[j].vL[i] / sl + s[j].vU[i] / su + dw;
                    }
                }
                if (KKT_SOLVE(P, s, rc, rd, true, dw, st, lane)) { solved = true; break; }
                if (dw == 0.0) dw = (dw_last == 0.0) ? DW_INIT : fmax(DW_MIN, dw_last * DW_DEC);
                else dw = (dw_last == 0.0 || 1e5 * dw_last < dw) ? dw * DW_INC_FIRST : dw * DW_INC;
                if (dw > DW_MAX) break;
            }
            if (!solved) { status = B200MPC_ERROR_IN_STEP_COMPUTATION; break; }
            if (dw > 0.0) dw_last = dw;
            {
                int bad = 0;
#pragma unroll
                for (int j = 0; j < J; ++j) {
#pragma unroll
                    for (int i = 0; i < 3; i++)
                        if (!isfinite(st[j].dX[i]) || !isfinite(st[j].dlam[i])) bad = 1;
#pragma unroll
                    for (int i = 0; i < 2; i++)
                        if (!isfinite(st[j].dU[i]) || !isfinite(st[j].dS[i]) || !isfinite(st[j].dyd[i])) bad = 1;
                }
                if (wor(bad)) { status = B200MPC_ERROR_IN_STEP_COMPUTATION; break; }
            }

            // ---- filter line search ----
            const double a_max = frac_to_bound<J>(P, s, st, tau, lane);
            LsRef ref;
            {
                double bar = 0, gb = 0;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int k = lane * J + j;
                    const bool dyn = k < N;
                    if (dyn) {
                        double prod = 1.0; // (the current iterate's slacks are strictly inside their bounds)
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            const double sl = s[j].S[i] - P.sL[i], su = P.sU[i] - s[j].S[i];
                            prod *= sl * su;
                            gb += (-mu / sl + mu / su) * st[j].dS[i];
                            gb += s[j].g[3 + i] * st[j].dU[i];
                        }
                        bar -= mu * log(prod);
                    }
                    if (k >= 1 && k <= N) {
#pragma unroll
                        for (int i = 0; i < 3; i++) gb += s[j].g[i] * st[j].dX[i];
                    }
                }
                ref.phi = df * fcur + wsum(bar);
                ref.gbd = wsum(gb);
                ref.theta = theta; ref.theta_min = theta_min; ref.theta_max = theta_max;
            }
            double a_min = GAMMA_THETA;
            if (ref.gbd < 0) {
                a_min = fmin(GAMMA_THETA, GAMMA_PHI * theta / (-ref.gbd));
                if (theta <= theta_min) a_min = fmin(a_min, DELTA_LS * pow_ool(theta, S_THETA) / pow_ool(-ref.gbd, S_PHI));
            }
            a_min *= ALPHA_MIN_FRAC;

            double alpha = a_max, alpha_acc = 0;
            int acc = 0; // 0 none, 1 newton step, 2 corrected step
            bool fa = false;
            int ntrial = 0;
            // BacktrackingLineSearch::DetectTinyStep: every primal component moves by less than tiny_step_tol = 10 eps
            // (relative) and the point is nearly feasible -> the full step is taken without a line search
            bool tiny;
            {
                int big = 0;
                double c2 = 0;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int k = lane * J + j;
                    if (k >= 1 && k <= N) {
#pragma unroll
                        for (int i = 0; i < 3; i++) {
                            if (!(fabs(st[j].dX[i]) <= TINY_STEP_TOL * (1.0 + fabs(s[j].X[i])))) big = 1;
                        }
                    }
                    if (k < N) {
#pragma unroll
                        for (int i = 0; i < 3; i++) c2 += rc[j][i] * rc[j][i];
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            if (!(fabs(st[j].dU[i]) <= TINY_STEP_TOL * (1.0 + fabs(s[j].U[i])))) big = 1;
                            if (!(fabs(st[j].dS[i]) <= TINY_STEP_TOL * (1.0 + fabs(s[j].S[i])))) big = 1;
                            c2 += rd[j][i] * rd[j][i];
                        }
                    }
                }
                tiny = !wor(big);
                if (tiny) tiny = sqrt(wsum(c2)) <= 1e-4;
            }
            if (tiny) {
                double th_t, phi_t;
                trial_eval<J, OBS>(P, OL, sox, soy, ocs, s, st, alpha, mu, df, lane, th_t, phi_t);
                if (isfinite(th_t) && isfinite(phi_t)) {
                    (void)ls_acceptable(ref, alpha, phi_t, th_t, fphi, ftheta, fvalid, fa); // only for the filter-augmentation rule
                    acc = 1;
                    alpha_acc = alpha;
                    if (tiny_last) tiny_flag = true;
                    double dy = 0;
#pragma unroll
                    for (int j = 0; j < J; ++j) {
#pragma unroll
                        for (int i = 0; i < 3; i++) dy = fmax(dy, fabs(st[j].dlam[i]));
#pragma unroll
                        for (int i = 0; i < 2; i++) dy = fmax(dy, fabs(st[j].dyd[i]));
                    }
                    tiny_last = wmax(dy) < TINY_STEP_Y_TOL;
                } else {
                    tiny = false;
                }
            }
            if (!tiny) { tiny_flag = false; tiny_last = false; }
            while (!acc) {
                double th_t, phi_t;
                trial_eval<J, OBS>(P, OL, sox, soy, ocs, s, st, alpha, mu, df, lane, th_t, phi_t);
                if (ntrial++ > 0) ls_extra++;
                if (ls_acceptable(ref, alpha, phi_t, th_t, fphi, ftheta, fvalid, fa)) { acc = 1; alpha_acc = alpha; break; }
                if (ntrial == 1 && P.max_soc > 0 && isfinite(th_t) && th_t >= theta) {
                    // second-order correction
#pragma unroll
                    for (int j = 0; j < J; ++j) {
#pragma unroll
                        for (int i = 0; i < 3; i++) csoc[j][i] = rc[j][i];
#pragma unroll
                        for (int i = 0; i < 2; i++) dsoc[j][i] = rd[j][i];
                    }
                    double alpha_soc = alpha, theta_soc_old = 0, th_trial = th_t;
                    int count = 0;
                    while (count < P.max_soc && !acc && (count == 0 || th_trial <= KAPPA_SOC * theta_soc_old)) {
                        theta_soc_old = th_trial;
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            const bool dyn = (lane * J + j) < N;
#pragma unroll
                            for (int i = 0; i < 3; i++) csoc[j][i] = alpha_soc * csoc[j][i] + s[j].ct[i];
#pragma unroll
                            for (int i = 0; i < 2; i++)
                                dsoc[j][i] = dyn ? (alpha_soc * dsoc[j][i] + (s[j].Ut[i] - s[j].St[i])) : 0.0;
                        }
                        if (!KKT_SOLVE(P, s, csoc, dsoc, true, dw, soc, lane)) break;
                        alpha_soc = frac_to_bound<J>(P, s, soc, tau, lane);
                        double phi_s;
                        trial_eval<J, OBS>(P, OL, sox, soy, ocs, s, soc, alpha_soc, mu, df, lane, th_trial, phi_s);
                        ls_extra++;
                        if (ls_acceptable(ref, alpha, phi_s, th_trial, fphi, ftheta, fvalid, fa)) { acc = 2; alpha_acc = alpha_soc; }
                        else count++;
                    }
                    if (acc) break;
                }
                alpha *= 0.5;
                if (alpha < a_min) break;
            }

            if (!acc) {
                // Restoration stand-in (same rule as oracle/mpc_oracle.c).  The filter is augmented with the current point, then
                //   stage 1: along the direction to the closed-form feasible point (U = S, X rolled out; S fixed) the longest
                //            step t = 1, 1/2, ... is taken whose point passes IPOPT's restoration acceptance (finite,
                //            theta <= kappa_resto * theta, no excessive objective increase, acceptable to the filter);
                //   stage 2: if there is none (the roll-out runs into an obstacle point), the family "plan shrunk towards
                //            standing still": S_l = u_c + l (S - u_c), U = S_l, X rolled out, l = 1/2, 1/4, ..., 0; the member
                //            with the lowest barrier objective is taken.
                // The multipliers restart.
                if (theta <= 1e-10 || n_resto >= MAX_RESTO) { status = B200MPC_RESTORATION_FAILED; break; }
                filter_add(ref.phi - GAMMA_PHI * theta, (1 - GAMMA_THETA) * theta, fphi, ftheta, fvalid, ring, lane);
                double uc[2];
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const double lo = P.sL[i], hi = P.sU[i];
                    const double pl = fmin(BOUND_PUSH * fmax(1.0, fabs(lo)), BOUND_FRAC * (hi - lo));
                    const double pu = fmin(BOUND_PUSH * fmax(1.0, fabs(hi)), BOUND_FRAC * (hi - lo));
                    uc[i] = fmin(fmax(0.0, lo + pl), hi - pu);
                }
                // direction to "U = S_l, S = S_l, X rolled out" into the step record of the correction (l < 0: S_l = S exactly)
                auto resto_direction = [&](double lam) {
                    double tS[J][2];
#pragma unroll
                    for (int j = 0; j < J; ++j)
#pragma unroll
                        for (int i = 0; i < 2; i++) tS[j][i] = (lam < 0.0) ? s[j].S[i] : uc[i] + lam * (s[j].S[i] - uc[i]);
                    // serial rollout X_{k+1} = F(X_k, S_l,k)
                    double y[3] = {x00, x01, x02};
                    for (int l = 0; l <= N / J; ++l) {
                        double z[3] = {y[0], y[1], y[2]};
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            const int ko = l * J + j;
                            if (ko > N) continue;
                            if (lane == l) {
                                soc[j].dX[0] = z[0] - s[j].X[0]; soc[j].dX[1] = z[1] - s[j].X[1]; soc[j].dX[2] = z[2] - s[j].X[2];
                            }
                            if (ko < N) {
                                double F[3];
                                dyn_value(P, z, tS[j], F);
                                z[0] = F[0]; z[1] = F[1]; z[2] = F[2];
                            }
                        }
                        y[0] = __shfl_sync(FULL, z[0], l); y[1] = __shfl_sync(FULL, z[1], l); y[2] = __shfl_sync(FULL, z[2], l);
                    }
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const int k = lane * J + j;
                        if (k > N || k == 0) { soc[j].dX[0] = soc[j].dX[1] = soc[j].dX[2] = 0; }
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            soc[j].dU[i] = (k < N) ? (tS[j][i] - s[j].U[i]) : 0.0;
                            soc[j].dS[i] = (k < N) ? (tS[j][i] - s[j].S[i]) : 0.0;
                        }
                    }
                };
                resto_direction(-1.0);
                double t_acc = 0.0;
                for (double t = 1.0; t >= RESTO_T_MIN; t *= 0.5) {
                    double th_r, phi_r;
                    trial_eval<J, OBS>(P, OL, sox, soy, ocs, s, soc, t, mu, df, lane, th_r, phi_r);
                    if (!isfinite(th_r) || !isfinite(phi_r)) continue;
                    if (!(th_r <= KAPPA_RESTO * theta)) continue;
                    if (phi_r > ref.phi) {
                        double bas = 1.0;
                        if (fabs(ref.phi) > 10.0) bas = log10_ool(fabs(ref.phi));
                        if (log10_ool(phi_r - ref.phi) > OBJ_MAX_INC + bas) continue;
                    }
                    const bool rej = fvalid && !(cmp_le(phi_r, fphi, fphi) || cmp_le(th_r, ftheta, ftheta));
                    if (__any_sync(FULL, rej)) continue;
                    t_acc = t;
                    break;
                }
                if (t_acc == 0.0) {
                    double lam = 0.5, best = __longlong_as_double(0x7ff0000000000000ll), lam_best = -1.0;
                    for (;;) {
                        resto_direction(lam);
                        double th_r, phi_r;
                        trial_eval<J, OBS>(P, OL, sox, soy, ocs, s, soc, 1.0, mu, df, lane, th_r, phi_r);
                        bool okc = isfinite(th_r) && isfinite(phi_r) && (th_r <= KAPPA_RESTO * theta) && (phi_r < best);
                        if (okc) {
                            const bool rej = fvalid && !(cmp_le(phi_r, fphi, fphi) || cmp_le(th_r, ftheta, ftheta));
                            okc = !__any_sync(FULL, rej);
                        }
                        if (okc) { best = phi_r; lam_best = lam; }
                        if (lam == 0.0) break;
                        lam = (lam * 0.5 >= RESTO_T_MIN) ? lam * 0.5 : 0.0;
                    }
                    if (lam_best >= 0.0) {
                        // the trial record (and the obstacle cache) must belong to the point that is taken
                        resto_direction(lam_best);
                        double th_r, phi_r;
                        trial_eval<J, OBS>(P, OL, sox, soy, ocs, s, soc, 1.0, mu, df, lane, th_r, phi_r);
                        t_acc = 1.0;
                    }
                }
                if (t_acc == 0.0) { status = B200MPC_RESTORATION_FAILED; break; }
                double zm = 0;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int k = lane * J + j;
                    const bool dyn = k < N;
                    if (k >= 1 && k <= N) { s[j].X[0] = s[j].Xt[0]; s[j].X[1] = s[j].Xt[1]; s[j].X[2] = s[j].Xt[2]; }
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        if (dyn) { s[j].U[i] = s[j].Ut[i]; s[j].S[i] = s[j].St[i]; zm = fmax(zm, fmax(s[j].vL[i], s[j].vU[i])); }
                        s[j].yd[i] = 0;
                    }
#pragma unroll
                    for (int i = 0; i < 3; i++) s[j].lam[i] = 0;
                }
                zm = wmax(zm);
                if (zm > 1e3) {
#pragma unroll
                    for (int j = 0; j < J; ++j) { s[j].vL[0] = s[j].vL[1] = 1.0; s[j].vU[0] = s[j].vU[1] = 1.0; }
                }
                oc_valid = OBS && (P.obs_form != B200MPC_OBS_NONE); // ocs holds the sums of the accepted point
                n_resto++;
                iter++;
                continue;
            }

            // ---- accept the trial point ----
            if (!fa) filter_add(ref.phi - GAMMA_PHI * theta, (1 - GAMMA_THETA) * theta, fphi, ftheta, fvalid, ring, lane);
            double a_z = 1.0;
            double dvL[J][2], dvU[J][2];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const bool dyn = (lane * J + j) < N;
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const double ds = (acc == 2) ? soc[j].dS[i] : st[j].dS[i];
                    const double sl = s[j].S[i] - P.sL[i], su = P.sU[i] - s[j].S[i];
                    dvL[j][i] = mu / sl - s[j].vL[i] - s[j].vL[i] / sl * ds;
                    dvU[j][i] = mu / su - s[j].vU[i] + s[j].vU[i] / su * ds;
                    if (dyn && dvL[j][i] < 0) a_z = fmin(a_z, -tau * s[j].vL[i] / dvL[j][i]);
                    if (dyn && dvU[j][i] < 0) a_z = fmin(a_z, -tau * s[j].vU[i] / dvU[j][i]);
                }
            }
            a_z = wmin(a_z);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int k = lane * J + j;
                const bool dyn = k < N;
                Stg &t = s[j];
                const Step &q = (acc == 2) ? soc[j] : st[j];
                if (k >= 1 && k <= N) {
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        t.X[i] = t.Xt[i]; // = X + alpha_acc*dX, the point the last trial evaluation (and ocs) belongs to
                        t.lam[i] += alpha_acc * q.dlam[i];
                    }
                }
                if (dyn) {
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        t.U[i] += alpha_acc * q.dU[i];
                        t.S[i] += alpha_acc * q.dS[i];
                        t.yd[i] += alpha_acc * q.dyd[i];
                        t.vL[i] += a_z * dvL[j][i];
                        t.vU[i] += a_z * dvU[j][i];
                        const double sl = t.S[i] - P.sL[i], su = P.sU[i] - t.S[i];
                        t.vL[i] = fmax(fmin(t.vL[i], KAPPA_SIGMA * mu / sl), mu / (KAPPA_SIGMA * sl));
                        t.vU[i] = fmax(fmin(t.vU[i], KAPPA_SIGMA * mu / su), mu / (KAPPA_SIGMA * su));
                    }
                }
            }
            iter++;
            oc_valid = OBS && (P.obs_form != B200MPC_OBS_NONE);
        }
    }

finish:
    // ---- store results (coalesced) ----
    {
        // objective at the returned point (value-only evaluation of the current iterate)
        double fl = 0;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int k = lane * J + j;
            const bool act = (k <= N), dyn = (k < N);
            const bool obs = OBS && P.obs_form != B200MPC_OBS_NONE && k >= P.obs_k0 && k <= P.obs_k1;
            double Xn[3], ct[3];
            NEXT3(Xn, X, j);
            fl += stage_value(P, OL, sox, soy, s[j].X, s[j].U, s[j].r, s[j].ub, Xn, ct, act, dyn, obs);
        }
        const double fsum = wsum(fl);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int k = lane * J + j;
            if (k <= N) {
                double *xo = A.X + (size_t)b * 3 * (N + 1) + 3 * k;
                xo[0] = s[j].X[0]; xo[1] = s[j].X[1]; xo[2] = s[j].X[2];
            }
            if (k < N) {
                double *uo = A.U + (size_t)b * 2 * N + 2 * k;
                uo[0] = s[j].U[0]; uo[1] = s[j].U[1];
            }
        }
        if (lane == 0) {
            if (A.cost) A.cost[b] = fsum;
            A.status[b] = status;
            if (A.iters) A.iters[b] = iter;
            if (A.ls) A.ls[b] = ls_extra;
        }
    }
    __syncwarp();
}

template <int J, bool OBS, bool SCAN>
__global__ void __launch_bounds__(128, OBS ? B200MPC_MIN_CTAS_OBS : B200MPC_MIN_CTAS) mpc_solve_kernel(const KParams P, const BatchArgs A) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int Mpad = (P.M + 3) & ~3;
    // per warp: the problem's obstacle lists (2*Mpad doubles) and the stage records of the KKT solve
    const size_t per_warp = 2 * (size_t)Mpad + (size_t)(P.N + 1) * (KKT_REC + 6) + (SCAN ? SCAN_NF * 32 : 0);
    double *sox = smem + (size_t)wid * per_warp, *soy = sox + Mpad, *rec = soy + Mpad;
    const KktRoles roles = kkt_roles(lane);
    if (A.resume) {
        // a launch behind the lane kernel or behind an iteration-bounded launch of this kernel: the batch is the list of
        // problems that one handed over
        const int n = (int)min(*A.resume_count, (unsigned)A.resume_cap);
        for (;;) {
            int r = 0;
            if (lane == 0) r = (int)atomicAdd(A.counter, 1u);
            r = __shfl_sync(FULL, r, 0);
            if (r >= n) break;
            const double *hrec = A.resume_rec + (size_t)r * HAND_REC(P.N);
            solve_one<J, OBS, SCAN>(P, A, (int)hrec[HAND_SCAL(P.N) + HS_B], lane, sox, soy, rec, roles, hrec);
        }
        return;
    }
    for (;;) {
        int b = 0;
        if (lane == 0) b = (int)atomicAdd(A.counter, 1u);
        b = __shfl_sync(FULL, b, 0);
        if (b >= A.B) break;
        solve_one<J, OBS, SCAN>(P, A, b, lane, sox, soy, rec, roles);
    }
}

// ---- NLP function evaluation kernel (K1-K3 only), one warp per problem -----------------------------------
template <int J>
__global__ void __launch_bounds__(128) mpc_eval_kernel(const KParams P, const EvalArgs A) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int Mpad = (P.M + 3) & ~3;
    double *sox = smem + (size_t)wid * 2 * Mpad, *soy = sox + Mpad;
    const int N = P.N;
    const int b = blockIdx.x * (blockDim.x >> 5) + wid;
    if (b >= A.B) return;
    ObsList OL;
    OL.n_eff = 1; OL.w0 = 1.0;
    if (P.obs_form != B200MPC_OBS_NONE) {
        const double *gx = A.ox + (size_t)A.obs_stride * b, *gy = A.oy + (size_t)A.obs_stride * b;
        for (int i = lane; i < P.M; i += 32) { sox[i] = gx[i]; soy[i] = gy[i]; }
        __syncwarp();
        OL = obstacle_list_setup(sox, soy, P.M, lane);
    }
    __syncwarp();
    Stg s[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int k = lane * J + j;
        Stg &t = s[j];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            t.X[i] = (k <= N) ? A.X[(size_t)b * 3 * (N + 1) + 3 * k + i] : 0.0;
            t.lam[i] = (k >= 1 && k <= N && A.lam) ? A.lam[(size_t)b * 3 * N + 3 * (k - 1) + i] : 0.0;
            t.r[i] = 0;
        }
        t.U[0] = t.U[1] = t.ub[0] = t.ub[1] = 0;
        if (k < N) {
            t.U[0] = A.U[(size_t)b * 2 * N + 2 * k];
            t.U[1] = A.U[(size_t)b * 2 * N + 2 * k + 1];
            if (P.ref_kind == B200MPC_REF_GOAL) {
                for (int i = 0; i < 3; i++) t.r[i] = A.xref[3 * (size_t)b + i];
            } else {
                for (int i = 0; i < 3; i++) t.r[i] = A.xref[(size_t)b * 3 * N + 3 * k + i];
                t.ub[0] = A.uref[(size_t)b * 2 * N + 2 * k];
                t.ub[1] = A.uref[(size_t)b * 2 * N + 2 * k + 1];
            }
        }
    }
    double fl = 0;
    double Xn0[3], ln0[3];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int k = lane * J + j;
        const bool act = (k <= N), dyn = (k < N);
        const bool obs = P.obs_form != B200MPC_OBS_NONE && k >= P.obs_k0 && k <= P.obs_k1;
        NEXT3(Xn0, X, j);
        NEXT3(ln0, lam, j);
        fl += stage_full(P, OL, sox, soy, s[j], Xn0, ln0, A.obj_scale, act, dyn, obs);
    }
    const double f = wsum(fl);
    if (lane == 0 && A.f) A.f[b] = f;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int k = lane * J + j;
        const Stg &t = s[j];
        if (k < N && A.c)
            for (int i = 0; i < 3; i++) A.c[(size_t)b * 3 * N + 3 * k + i] = t.c[i];
        if (A.grad) {
            if (k >= 1 && k <= N)
                for (int i = 0; i < 3; i++) A.grad[(size_t)b * 5 * N + 3 * (k - 1) + i] = t.g[i];
            if (k < N)
                for (int i = 0; i < 2; i++) A.grad[(size_t)b * 5 * N + 3 * N + 2 * k + i] = t.g[3 + i];
        }
        if (k <= N && A.stage) {
            double *o = A.stage + ((size_t)b * (N + 1) + k) * 36;
            o[0] = t.a13; o[1] = t.a23; o[2] = t.b11; o[3] = t.b12; o[4] = t.b21; o[5] = t.b22;
            double H[25];
            for (int i = 0; i < 25; i++) H[i] = 0;
            H[0] = t.hxx; H[1] = t.hxy; H[5] = t.hxy; H[6] = t.hyy; H[12] = t.htt;
            H[13] = t.htv; H[17] = t.htv; H[14] = t.htw; H[22] = t.htw; H[18] = t.hvv;
            H[19] = t.hvw; H[23] = t.hvw; H[24] = t.hww;
            for (int i = 0; i < 25; i++) o[6 + i] = H[i];
            for (int i = 0; i < 3; i++) o[31 + i] = t.c[i];
            o[34] = o[35] = 0;
        }
    }
}

#include "tpp_kernel.cuh"
#include "tpp_fused.cuh"
#include "obstacles_kernel.cuh"
#include "refgen_kernel.cuh"
#include "control_kernel.cuh"
#include "costmap_kernel.cuh"
#include "sensor_kernel.cuh"

// =============================================================================================================
// Host side: C ABI
// =============================================================================================================

static std::mutex g_err_mtx;
static std::string g_create_err;

struct b200mpc_handle {
    b200mpc_params prm;
    KParams kp;
    int device;
    int J;
    int sm_count;
    int ctas;
    size_t smem_bytes, smem_scan_bytes; // dynamic shared memory of the warp kernel: serial-recursion / scan instance
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    unsigned int *d_counter;
    // staging buffers for the host-pointer entry points
    char *d_buf;
    size_t d_cap;
    long long launches;
    // lane-per-problem kernel (tpp_kernel.cuh): persistent grid and its HBM workspace
    int kernel_kind;   // B200MPC_KERNEL_AUTO / _WARP / _LANE
    int last_kind;     // kernel used by the most recent solve
    int tpp_ctas;
    size_t tpp_obs_smem; // bytes of per-warp obstacle-list buffers behind TPP_SMEM_BYTES (0: lists are read from global memory)
    int tpp_cta_sync, tpp_b_passes;
    // straggler hand-over from the lane kernel to the warp kernel (BatchArgs::hand_rec)
    int hand_iter, hand_thin;   // hand_iter <= 0: off
    int hand_iter_tail;
    double *d_hand[2];          // two record buffers: a launch reads one and writes the other
    size_t hand_cap[2];
    unsigned int *d_hand_count; // [2]
    int warp_caps[4], n_warp_caps; // iteration bounds of the warp kernel's cascade (launch p exports at warp_caps[p]; the last launch has none)
    int kkt_scan_mode; // warp kernel: -1 scan recursion for batches that leave warps idle, 0 never, 1 always
    int lane_spec;     // template instance of the lane kernels (TPP_SPEC_*)
    int lane_fused;    // 1: two-sweep lane kernel (tpp_fused.cuh), 0: three-sweep lane kernel (tpp_kernel.cuh)
    double *d_ws, *d_filt;
    unsigned long long *d_stats;
    // streamed host-buffer solves: copy-in / copy-out streams, device words (avail, done[chunks]) and host-mapped
    // words (flags[chunks], marks[chunks])
    cudaStream_t cstream, ostream;
    cudaEvent_t ev_sync;
    unsigned int *d_sync;
    unsigned int *h_sync, *h_sync_dev;
    int last_streamed;
    // small host-buffer solves (single-solve latency): page-locked, device-mapped staging the kernel reads and
    // writes directly over PCIe — no cudaMemcpy calls on the path
    char *h_small, *h_small_dev;
    size_t small_cap;
    std::string err;
};
#define B200MPC_MAX_CHUNKS 64
#ifndef B200MPC_STREAM_MIN_BATCH
#define B200MPC_STREAM_MIN_BATCH 131072 /* host-buffer solves from this size on are streamed in chunks */
#endif

static int set_err(b200mpc_handle *h, int code, const std::string &msg) {
    if (h) h->err = msg;
    else {
        std::lock_guard<std::mutex> g(g_err_mtx);
        g_create_err = msg;
    }
    return code;
}

#define CU_TRY(h, call)                                                                              \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return set_err(h, B200MPC_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));   \
    } while (0)

extern "C" int b200mpc_abi_version(void) { return B200MPC_ABI_VERSION; }

extern "C" void b200mpc_default_options(b200mpc_params *p) {
    p->tol = 1e-8;
    p->max_iter = 3000;
    p->acceptable_tol = 1e-6;
    p->acceptable_iter = 15;
    p->mu_init = 0.1;
    p->max_soc = 4;
}

extern "C" const char *b200mpc_last_error(const b200mpc_handle *h) {
    if (h) return h->err.c_str();
    std::lock_guard<std::mutex> g(g_err_mtx);
    return g_create_err.c_str();
}

template <int J>
static cudaError_t configure_kernels(size_t smem) {
    cudaError_t e = cudaFuncSetAttribute(mpc_solve_kernel<J, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(mpc_solve_kernel<J, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (J == 1) {
        e = cudaFuncSetAttribute(mpc_solve_kernel<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(mpc_solve_kernel<1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    return cudaFuncSetAttribute(mpc_eval_kernel<J>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

template <int J>
static cudaError_t occupancy(int *blocks, size_t smem, bool obs) {
    if (obs) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, mpc_solve_kernel<J, true, false>, 128, smem);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, mpc_solve_kernel<J, false, false>, 128, smem);
}

extern "C" b200mpc_handle *b200mpc_create(const b200mpc_params *p, int device) {
    if (!p) { set_err(nullptr, B200MPC_E_ARG, "params is NULL"); return nullptr; }
    if (p->N < 1 || p->N > B200MPC_MAX_N) { set_err(nullptr, B200MPC_E_ARG, "N out of range [1,127]"); return nullptr; }
    if (p->obs_form != B200MPC_OBS_NONE && (p->M < 1 || p->M > B200MPC_MAX_M)) {
        set_err(nullptr, B200MPC_E_ARG, "M out of range [1,1024]");
        return nullptr;
    }
    if (!(p->dt > 0) || !(p->u_lo[0] < p->u_hi[0]) || !(p->u_lo[1] < p->u_hi[1])) {
        set_err(nullptr, B200MPC_E_ARG, "dt must be > 0 and u_lo < u_hi");
        return nullptr;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_err(nullptr, B200MPC_E_NODEVICE,
                std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
        return nullptr;
    }
    if (device < 0 || device >= ndev) { set_err(nullptr, B200MPC_E_ARG, "device index out of range"); return nullptr; }
    b200mpc_handle *h = new b200mpc_handle();
    h->prm = *p;
    h->device = device;
    h->launches = 0;
    h->d_buf = nullptr;
    h->d_cap = 0;
    h->kernel_kind = B200MPC_KERNEL_AUTO;
    h->last_kind = B200MPC_KERNEL_WARP;
    h->tpp_ctas = 0;
    h->d_ws = nullptr;
    h->d_filt = nullptr;
    h->d_stats = nullptr;
    h->cstream = h->ostream = nullptr;
    h->ev_sync = nullptr;
    h->d_sync = nullptr;
    h->h_sync = h->h_sync_dev = nullptr;
    h->last_streamed = 0;
    h->h_small = h->h_small_dev = nullptr;
    h->small_cap = 0;
    h->stream = nullptr;
    h->ev0 = h->ev1 = nullptr;
    h->d_counter = nullptr;
    if (const char *ek = getenv("B200MPC_KERNEL")) {
        if (!strcmp(ek, "warp")) h->kernel_kind = B200MPC_KERNEL_WARP;
        else if (!strcmp(ek, "lane")) h->kernel_kind = B200MPC_KERNEL_LANE;
    }
    KParams &k = h->kp;
    memset(&k, 0, sizeof(k));
    k.N = p->N; k.M = (p->obs_form == B200MPC_OBS_NONE) ? 0 : p->M;
    k.integrator = p->integrator; k.ref_kind = p->ref_kind; k.obs_form = p->obs_form;
    k.obs_k0 = p->obs_k0; k.obs_k1 = p->obs_k1; k.max_iter = p->max_iter;
    k.acceptable_iter = p->acceptable_iter; k.max_soc = p->max_soc;
    k.dt = p->dt; k.kappa = p->kappa; k.obs_c = p->obs_c; k.obs_r = p->obs_r;
    k.tol = p->tol; k.acceptable_tol = p->acceptable_tol; k.mu_init = p->mu_init;
    for (int i = 0; i < 3; i++) k.Q[i] = p->Q[i];
    for (int i = 0; i < 2; i++) {
        k.R[i] = p->R[i]; k.u_lo[i] = p->u_lo[i]; k.u_hi[i] = p->u_hi[i];
        k.sL[i] = p->u_lo[i] - BOUND_RELAX * fmax(1.0, fabs(p->u_lo[i]));
        k.sU[i] = p->u_hi[i] + BOUND_RELAX * fmax(1.0, fabs(p->u_hi[i]));
    }
    k.inv_r2 = 1.0 / (p->obs_r * p->obs_r);
    k.mu_floor = fmin(p->tol, 1e-4) / (K_EPS + 1.0);
    // parallel-in-time Riccati recursion: -1 = by batch size (launch_solve), 0 / 1 = forced (env B200MPC_KKT_SCAN)
    h->kkt_scan_mode = -1;
    if (const char *es = getenv("B200MPC_KKT_SCAN")) h->kkt_scan_mode = (es[0] == '1') ? 1 : 0;
    k.kkt_scan = 0;
    h->J = (p->N + 1 + 31) / 32;
    const int Mpad = (k.M + 3) & ~3;
    // per warp: obstacle lists + the stage records of the lane-parallel KKT solve (KKT_REC doubles per stage)
    h->smem_bytes = (size_t)4 * (2 * (size_t)Mpad + (size_t)(p->N + 1) * (KKT_REC + 6)) * sizeof(double);
    h->smem_scan_bytes = h->smem_bytes + (h->J == 1 ? (size_t)4 * SCAN_NF * 32 * sizeof(double) : 0); // scan instance: + its scratch
    if (h->smem_scan_bytes > 227 * 1024) {
        set_err(nullptr, B200MPC_E_ARG, "N and M too large for the shared-memory staging of the warp kernel");
        b200mpc_destroy(h);
        return nullptr;
    }
    auto fail = [&](const std::string &m) -> b200mpc_handle * {
        set_err(nullptr, B200MPC_E_CUDA, m);
        b200mpc_destroy(h); // releases whatever has been created so far
        return nullptr;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(cudaGetErrorString(e));
    h->sm_count = prop.multiProcessorCount;
    int blocks = 0;
    // The dynamic shared-memory limit is an attribute of the kernel, shared by every handle of the process: it is
    // raised to the device's opt-in maximum once (a later handle with a smaller need must not lower it under an
    // earlier handle's launches); occupancy is computed for this handle's own size.
    const size_t smem_cap = prop.sharedMemPerBlockOptin;
    if (h->smem_scan_bytes > smem_cap) return fail("N and M too large for the shared memory of this device");
    switch (h->J) {
        case 1: e = configure_kernels<1>(smem_cap); if (e == cudaSuccess) e = occupancy<1>(&blocks, h->smem_bytes, p->obs_form != B200MPC_OBS_NONE); break;
        case 2: e = configure_kernels<2>(smem_cap); if (e == cudaSuccess) e = occupancy<2>(&blocks, h->smem_bytes, p->obs_form != B200MPC_OBS_NONE); break;
        case 3: e = configure_kernels<3>(smem_cap); if (e == cudaSuccess) e = occupancy<3>(&blocks, h->smem_bytes, p->obs_form != B200MPC_OBS_NONE); break;
        default: e = configure_kernels<4>(smem_cap); if (e == cudaSuccess) e = occupancy<4>(&blocks, h->smem_bytes, p->obs_form != B200MPC_OBS_NONE); break;
    }
    if (e != cudaSuccess) return fail(std::string("kernel configuration: ") + cudaGetErrorString(e));
    if (blocks < 1) blocks = 1;
    h->ctas = blocks * h->sm_count; // persistent grid: a multiple of the SM count (148 on B200)
    int tblocks = 0;
    const size_t tpp_smem = TPP_SMEM_BYTES;
    // lane kernels: one instance per problem family (tpp_kernel.cuh: TPP_SPEC_*)
    h->lane_spec = TPP_SPEC_GENERIC;
    if (p->integrator == B200MPC_RK4 && p->ref_kind == B200MPC_REF_GOAL) h->lane_spec = TPP_SPEC_RK4_GOAL;
    if (p->integrator == B200MPC_EULER && p->ref_kind == B200MPC_REF_TRAJ) h->lane_spec = TPP_SPEC_EULER_TRAJ;
    if (p->obs_form != B200MPC_OBS_NONE) // obstacle cost: the instance of variant A, or the generic one (run-time switch)
        h->lane_spec = (h->lane_spec == TPP_SPEC_RK4_GOAL) ? TPP_SPEC_RK4_GOAL_OBS : TPP_SPEC_GENERIC;
    if (getenv("B200MPC_LANE_GENERIC")) h->lane_spec = TPP_SPEC_GENERIC;
    {
        const void *fns[7] = {(const void *)mpc_solve_tpp_kernel<0>, (const void *)mpc_solve_tpp_kernel<1>,
                              (const void *)mpc_solve_tpp_kernel<2>, (const void *)mpc_solve_tppf_kernel<0>,
                              (const void *)mpc_solve_tppf_kernel<1>, (const void *)mpc_solve_tppf_kernel<2>,
                              (const void *)mpc_solve_tpp_kernel<3>};
        // obstacle cost: one obstacle list per warp behind the kernel's own shared memory, if it fits
        h->tpp_obs_smem = 0;
        if (p->obs_form != B200MPC_OBS_NONE) {
            const size_t lists = (size_t)(TPP_THREADS / 32) * 2 * (size_t)((p->M + 3) & ~3) * sizeof(double);
            if (tpp_smem + lists <= smem_cap) h->tpp_obs_smem = lists;
        }
        for (int i = 0; i < 7; i++) {
            const bool with_lists = (i == 0 || i == 6);
            if ((e = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(tpp_smem + (with_lists ? h->tpp_obs_smem : 0)))) != cudaSuccess)
                return fail(std::string("kernel configuration: ") + cudaGetErrorString(e));
        }
    }
    // hand-over threshold: about twice the mean iteration count of the family (variant A ~60, variants B / C ~22)
    // (measured, 1 M problems: variant B 198.2 -> 190.4 ms with 60 / 8; variant A: the launch no longer lasts as long as its
    // slowest problem, 1354 -> 417 ms for 65 536 problems with one 3000-iteration straggler)
    h->hand_iter = (p->obs_form != B200MPC_OBS_NONE) ? 128 : 60;
    // hand_thin: also hand over every problem of a warp that has thinned out to <= hand_thin active lanes once the work queue is
    // empty.  Off (-1) by default: which problems that rule catches depends on the timing of the launch, and results would
    // no longer be bit-identical from run to run (the two kernels agree to rounding, not to the bit).
    h->hand_thin = -1;
    // The last wave: the problems with the highest indices are pulled last, nothing follows them, and the lanes they leave
    // stay empty — the launch drains for as many trips as their longest takes.  They leave after hand_iter_tail iterations
    // (a rule on the problem index: deterministic).
    // Measured (variant B, cold starts, threshold 0 / 28): 1 M problems 169.4 / 163.9 ms, 131 072: 31.8 / 29.7 ms, 32 768: 19.4 /
    // 15.1 ms; thresholds of 12 and 20 overload the resuming warp kernel (42 / 33 ms at 131 072).  Variant A, 262 144: 838 / 822 ms
    // with 64.
    h->hand_iter_tail = (p->obs_form != B200MPC_OBS_NONE) ? 64 : 28;
    if (const char *eh = getenv("B200MPC_HAND_ITER_TAIL")) h->hand_iter_tail = atoi(eh);
    if (const char *eh = getenv("B200MPC_HAND_ITER")) h->hand_iter = atoi(eh);
    if (const char *eh = getenv("B200MPC_HAND_THIN")) h->hand_thin = atoi(eh);
    h->d_hand[0] = h->d_hand[1] = nullptr; h->hand_cap[0] = h->hand_cap[1] = 0; h->d_hand_count = nullptr;
    // warp kernel, batches of more than two waves of the persistent grid: optional cascade of iteration-bounded launches
    // (B200MPC_WARP_CAPS="64,160").  OFF by default — measured, it loses: config 3 variant A 29.7 ms in one launch, 33.2 ms
    // as 64 / 160 / rest; variant B 3.79 vs 4.25 ms with a bound of 32.  The persistent grid already starts a new problem the
    // moment a warp is free, a single launch lasts about as long as its longest problem alone (343 iterations sharing an SM
    // with seven other warps), and every bounded launch adds a tail of its own.
    h->n_warp_caps = 0;
    if (const char *ew = getenv("B200MPC_WARP_CAPS")) { // "64,160" | "0" (off)
        h->n_warp_caps = 0;
        for (const char *q = ew; *q && h->n_warp_caps < 4;) {
            const int v = atoi(q);
            if (v > 0) h->warp_caps[h->n_warp_caps++] = v;
            while (*q && *q != ',') q++;
            if (*q == ',') q++;
        }
    }
    h->tpp_cta_sync = (p->obs_form != B200MPC_OBS_NONE) ? 2 : 1; // (tpp_kernel.cuh, TppArgs::cta_sync; 2 = lock-step per scheduler group)
    // two executions of block B per trip (TppArgs::b_passes): 15 % (variants B / C) to 49 % (variant A) of the factorisations need
    // an inertia correction, which used to cost the lane a whole trip; measured on 1 M cold variant-B problems
    // 181.7 / 165.5 / 174.7 ms with 1 / 2 / 3 executions, variant C 181.7 -> 173.7 ms, variant A (262 144) 895 -> 863 ms
    h->tpp_b_passes = 2;
    if (const char *eb = getenv("B200MPC_LANE_BPASSES")) h->tpp_b_passes = std::max(1, atoi(eb));
    if (const char *es = getenv("B200MPC_LANE_SYNC")) h->tpp_cta_sync = (es[0] >= '0' && es[0] <= '3') ? es[0] - '0' : h->tpp_cta_sync;
    h->lane_fused = B200MPC_LANE_FUSED_DEFAULT;
    if (const char *ef = getenv("B200MPC_LANE_FUSED")) h->lane_fused = (ef[0] == '1');
    if (p->obs_form != B200MPC_OBS_NONE) h->lane_fused = 0; // the two-sweep kernel has no obstacle cost
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tblocks, mpc_solve_tpp_kernel<0>, TPP_THREADS, tpp_smem)) != cudaSuccess)
        return fail(std::string("kernel configuration: ") + cudaGetErrorString(e));
    if (tblocks < 1) tblocks = 1;
    if (const char *ec = getenv("B200MPC_LANE_CTAS_PER_SM")) tblocks = std::max(1, std::min(tblocks, atoi(ec))); // tuning experiments
    h->tpp_ctas = tblocks * h->sm_count;
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(cudaGetErrorString(e));
    if ((e = cudaEventCreate(&h->ev0)) != cudaSuccess) return fail(cudaGetErrorString(e));
    if ((e = cudaEventCreate(&h->ev1)) != cudaSuccess) return fail(cudaGetErrorString(e));
    if ((e = cudaMalloc(&h->d_counter, sizeof(unsigned int))) != cudaSuccess) return fail(cudaGetErrorString(e));
    if ((e = cudaStreamCreateWithFlags(&h->cstream, cudaStreamNonBlocking)) != cudaSuccess) return fail(cudaGetErrorString(e));
    if ((e = cudaStreamCreateWithFlags(&h->ostream, cudaStreamNonBlocking)) != cudaSuccess) return fail(cudaGetErrorString(e));
    if ((e = cudaEventCreateWithFlags(&h->ev_sync, cudaEventDisableTiming)) != cudaSuccess) return fail(cudaGetErrorString(e));
    if ((e = cudaMalloc(&h->d_sync, (1 + B200MPC_MAX_CHUNKS) * sizeof(unsigned int))) != cudaSuccess) return fail(cudaGetErrorString(e));
    if ((e = cudaHostAlloc(&h->h_sync, (2 * B200MPC_MAX_CHUNKS + 1) * sizeof(unsigned int), cudaHostAllocMapped)) != cudaSuccess)
        return fail(cudaGetErrorString(e));
    if ((e = cudaHostGetDevicePointer(&h->h_sync_dev, h->h_sync, 0)) != cudaSuccess) return fail(cudaGetErrorString(e));
    h->small_cap = (size_t)1 << 20;
    if ((e = cudaHostAlloc(&h->h_small, h->small_cap, cudaHostAllocMapped)) != cudaSuccess) return fail(cudaGetErrorString(e));
    if ((e = cudaHostGetDevicePointer(&h->h_small_dev, h->h_small, 0)) != cudaSuccess) return fail(cudaGetErrorString(e));
    return h;
}

extern "C" void b200mpc_destroy(b200mpc_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->d_buf) cudaFree(h->d_buf);
    if (h->d_hand[0]) cudaFree(h->d_hand[0]);
    if (h->d_hand[1]) cudaFree(h->d_hand[1]);
    if (h->d_hand_count) cudaFree(h->d_hand_count);
    if (h->d_ws) cudaFree(h->d_ws);
    if (h->d_filt) cudaFree(h->d_filt);
    if (h->d_stats) cudaFree(h->d_stats);
    if (h->d_counter) cudaFree(h->d_counter);
    if (h->cstream) { cudaStreamSynchronize(h->cstream); cudaStreamDestroy(h->cstream); }
    if (h->ostream) { cudaStreamSynchronize(h->ostream); cudaStreamDestroy(h->ostream); }
    if (h->ev_sync) cudaEventDestroy(h->ev_sync);
    if (h->d_sync) cudaFree(h->d_sync);
    if (h->h_sync) cudaFreeHost(h->h_sync);
    if (h->h_small) cudaFreeHost(h->h_small);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

// ---- FP64 DFMA peak micro-benchmark (roofline denominator; MEASURED_PEAKS.json has no FP64 entry) ----------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
    double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double b = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

extern "C" int b200mpc_measure_fp64_peak(b200mpc_handle *h, double *tflops_out) {
    if (!h || !tflops_out) return B200MPC_E_ARG;
    CU_TRY(h, cudaSetDevice(h->device));
    const int blocks = h->sm_count * 8, threads = 256, iters = 1 << 15;
    double *d_out = nullptr;
    CU_TRY(h, cudaMalloc(&d_out, sizeof(double) * (size_t)blocks * threads));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CU_TRY(h, cudaEventRecord(h->ev0, h->stream));
        fp64_peak_kernel<<<blocks, threads, 0, h->stream>>>(d_out, iters);
        CU_TRY(h, cudaGetLastError());
        CU_TRY(h, cudaEventRecord(h->ev1, h->stream));
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        float ms = 0;
        CU_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        if (rep > 0 && ms < best) best = ms;
        h->launches++;
    }
    cudaFree(d_out);
    const double flops = (double)blocks * threads * (double)iters * 8.0 * 2.0;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    return 0;
}

extern "C" int b200mpc_sizeof_params(void) { return (int)sizeof(b200mpc_params); }

extern "C" long long b200mpc_launch_count(const b200mpc_handle *h) { return h ? h->launches : 0; }

extern "C" float b200mpc_last_kernel_ms(b200mpc_handle *h) {
    if (!h) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.f;
    return ms;
}

static int check_args(b200mpc_handle *h, int B, const double *x0, const double *xref, const double *uref,
                      const double *obs_x, const double *obs_y, int obs_stride) {
    if (!h) return B200MPC_E_ARG;
    if (B < 0) return set_err(h, B200MPC_E_ARG, "B < 0");
    if (B > 0 && (!x0 || !xref)) return set_err(h, B200MPC_E_ARG, "x0 / xref is NULL");
    if (h->prm.ref_kind == B200MPC_REF_TRAJ && B > 0 && !uref) return set_err(h, B200MPC_E_ARG, "uref is NULL for a trajectory reference");
    if (h->prm.obs_form != B200MPC_OBS_NONE && B > 0) {
        if (!obs_x || !obs_y) return set_err(h, B200MPC_E_ARG, "obstacle lists are required: the obstacle cost is active");
        if (obs_stride != 0 && obs_stride < h->prm.M) return set_err(h, B200MPC_E_ARG, "obs_stride must be 0 or >= M");
    }
    return 0;
}

extern "C" int b200mpc_set_kernel(b200mpc_handle *h, int kind) {
    if (!h) return B200MPC_E_ARG;
    if (kind != B200MPC_KERNEL_AUTO && kind != B200MPC_KERNEL_WARP && kind != B200MPC_KERNEL_LANE)
        return set_err(h, B200MPC_E_ARG, "unknown kernel kind");
    h->kernel_kind = kind;
    return 0;
}

extern "C" int b200mpc_last_kernel_kind(const b200mpc_handle *h) { return h ? h->last_kind : B200MPC_E_ARG; }

extern "C" int b200mpc_last_solve_chunks(const b200mpc_handle *h) { return h ? h->last_streamed : B200MPC_E_ARG; }

// Diagnostics of the lane-per-problem kernel (only counted in builds with -DTPP_STATS=1): cumulative
// {executions, active lanes} of the sweeps B, F, T and of the trips; zeros otherwise.
extern "C" int b200mpc_lane_kernel_stats(b200mpc_handle *h, unsigned long long out[8]) {
    if (!h || !out) return B200MPC_E_ARG;
    for (int i = 0; i < 8; i++) out[i] = 0;
    if (!h->d_stats) return 0;
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaMemcpy(out, h->d_stats, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return 0;
}

// Lane-per-problem kernel: persistent grid, one workspace stripe per warp (allocated on first use).
struct StreamWords {
    const unsigned *avail;
    unsigned *done, *flags;
    int chunk;
    const unsigned *abort = nullptr; // host-mapped: set by the host when the streamed call fails
};

static int launch_warp_kernel(b200mpc_handle *h, const BatchArgs &a, int grid, bool scan, cudaStream_t stream);

// record buffer `which` (0 / 1) with room for `cap` problems, its counter zeroed on `stream`
static int ensure_hand_buffer(b200mpc_handle *h, int which, size_t cap, cudaStream_t stream) {
    const int N = h->prm.N;
    if (cap > h->hand_cap[which]) {
        if (h->d_hand[which]) { cudaFree(h->d_hand[which]); h->d_hand[which] = nullptr; h->hand_cap[which] = 0; }
        cudaError_t e = cudaMalloc(&h->d_hand[which], cap * HAND_REC(N) * sizeof(double));
        if (e != cudaSuccess) return set_err(h, B200MPC_E_NOMEM, std::string("cudaMalloc(hand-over records): ") + cudaGetErrorString(e));
        h->hand_cap[which] = cap;
    }
    if (!h->d_hand_count) {
        cudaError_t e = cudaMalloc(&h->d_hand_count, 2 * sizeof(unsigned int));
        if (e != cudaSuccess) return set_err(h, B200MPC_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
    CU_TRY(h, cudaMemsetAsync(h->d_hand_count + which, 0, sizeof(unsigned int), stream));
    return 0;
}

static int launch_solve_tpp(b200mpc_handle *h, const BatchArgs &a_in, cudaStream_t stream, const StreamWords *sw = nullptr) {
    BatchArgs a = a_in;
    // straggler hand-over (not for streamed host-buffer solves: their per-chunk completion flags count lane-kernel results)
    const bool hand = h->hand_iter > 0 && !sw;
    if (hand) {
        // room for a quarter of the batch: a lane that finds the buffer full stays in the lane kernel for good, and a
        // 3000-iteration problem left there holds the launch for seconds (measured with room for an eighth and a threshold of 96)
        const int rc = ensure_hand_buffer(h, 0, (size_t)a.B / 4 + 4096, stream);
        if (rc) return rc;
        a.hand_rec = h->d_hand[0]; a.hand_count = h->d_hand_count; a.hand_cap = (int)h->hand_cap[0];
        if (const char *ec = getenv("B200MPC_HAND_CAP")) a.hand_cap = std::min(a.hand_cap, std::max(0, atoi(ec))); // tests: the buffer-full path
        a.hand_iter = h->hand_iter; a.hand_thin = h->hand_thin;
        a.hand_iter_tail = h->hand_iter_tail;
    }
    const size_t nwarps = (size_t)h->tpp_ctas * (TPP_THREADS / 32);
    if (!h->d_ws) {
        const size_t ws_bytes = nwarps * (size_t)(h->prm.N + 1) * TPP_STAGE_B_MAX;
        cudaError_t e = cudaMalloc(&h->d_ws, ws_bytes);
        if (e != cudaSuccess) return set_err(h, B200MPC_E_NOMEM, std::string("cudaMalloc(workspace): ") + cudaGetErrorString(e));
        e = cudaMalloc(&h->d_filt, nwarps * 64 * 32 * sizeof(double));
        if (e != cudaSuccess) return set_err(h, B200MPC_E_NOMEM, std::string("cudaMalloc(filter): ") + cudaGetErrorString(e));
        e = cudaMalloc(&h->d_stats, 8 * sizeof(unsigned long long));
        if (e != cudaSuccess) return set_err(h, B200MPC_E_NOMEM, std::string("cudaMalloc(stats): ") + cudaGetErrorString(e));
        cudaMemset(h->d_stats, 0, 8 * sizeof(unsigned long long));
    }
    CU_TRY(h, cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned int), stream));
    int grid = h->tpp_ctas;
    const int need = (a.B + TPP_THREADS - 1) / TPP_THREADS;
    if (need < grid) grid = need;
    if (grid < 1) grid = 1;
    TppArgs t;
    t.a = a; t.ws = h->d_ws; t.filt = h->d_filt; t.stats = h->d_stats;
    t.avail = sw ? sw->avail : nullptr; t.done = sw ? sw->done : nullptr; t.flags = sw ? sw->flags : nullptr;
    t.chunk = sw ? sw->chunk : 1;
    t.abort = sw ? sw->abort : nullptr;
    t.cta_sync = h->lane_fused ? 1 : h->tpp_cta_sync;
    t.obs_smem = h->tpp_obs_smem ? 1 : 0;
    t.b_passes = h->tpp_b_passes;
    t.stage_b = (h->lane_spec == TPP_SPEC_GENERIC || h->lane_spec == TPP_SPEC_RK4_GOAL_OBS) ? TPP_STAGE_B_OF(TPP_SPEC_GENERIC) : TPP_STAGE_B_OF(TPP_SPEC_RK4_GOAL);
    const size_t smem_obs = TPP_SMEM_BYTES + h->tpp_obs_smem;
    CU_TRY(h, cudaEventRecord(h->ev0, stream));
    const int spec = h->lane_spec;
    if (h->lane_fused) {
        if (spec == TPP_SPEC_RK4_GOAL) mpc_solve_tppf_kernel<TPP_SPEC_RK4_GOAL><<<grid, TPP_THREADS, TPP_SMEM_BYTES, stream>>>(h->kp, t);
        else if (spec == TPP_SPEC_EULER_TRAJ) mpc_solve_tppf_kernel<TPP_SPEC_EULER_TRAJ><<<grid, TPP_THREADS, TPP_SMEM_BYTES, stream>>>(h->kp, t);
        else mpc_solve_tppf_kernel<TPP_SPEC_GENERIC><<<grid, TPP_THREADS, TPP_SMEM_BYTES, stream>>>(h->kp, t);
    } else {
        if (spec == TPP_SPEC_RK4_GOAL) mpc_solve_tpp_kernel<TPP_SPEC_RK4_GOAL><<<grid, TPP_THREADS, TPP_SMEM_BYTES, stream>>>(h->kp, t);
        else if (spec == TPP_SPEC_RK4_GOAL_OBS) mpc_solve_tpp_kernel<TPP_SPEC_RK4_GOAL_OBS><<<grid, TPP_THREADS, smem_obs, stream>>>(h->kp, t);
        else if (spec == TPP_SPEC_EULER_TRAJ) mpc_solve_tpp_kernel<TPP_SPEC_EULER_TRAJ><<<grid, TPP_THREADS, TPP_SMEM_BYTES, stream>>>(h->kp, t);
        else mpc_solve_tpp_kernel<TPP_SPEC_GENERIC><<<grid, TPP_THREADS, smem_obs, stream>>>(h->kp, t);
    }
    CU_TRY(h, cudaGetLastError());
    if (hand) {
        // the problems the lane kernel handed over: warp kernel, resuming from the records (their number stays on the device)
        BatchArgs r = a;
        r.resume = 1; r.resume_rec = a.hand_rec; r.resume_count = a.hand_count; r.resume_cap = a.hand_cap;
        r.hand_rec = nullptr; r.hand_count = nullptr; r.hand_cap = 0; r.hand_iter = 0;
        CU_TRY(h, cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned int), stream));
        const int rc = launch_warp_kernel(h, r, h->ctas, false, stream);
        if (rc) return rc;
        h->launches++;
    }
    CU_TRY(h, cudaEventRecord(h->ev1, stream));
    h->launches++;
    h->last_kind = B200MPC_KERNEL_LANE;
    return 0;
}

// Kernel choice: the lane-per-problem kernel needs enough problems to fill the machine with lanes; below that the
// warp-per-problem kernel is used.  With the obstacle cost a lane-kernel launch lasts as long as its slowest problems
// (hundreds of trips), so the crossover lies higher.
#ifndef B200MPC_LANE_KERNEL_MIN_BATCH_OBS
#define B200MPC_LANE_KERNEL_MIN_BATCH_OBS 131072
#endif
static int choose_kernel(const b200mpc_handle *h, int B) {
    int kind = h->kernel_kind;
    const bool obs = h->prm.obs_form != B200MPC_OBS_NONE;
    if (kind == B200MPC_KERNEL_AUTO)
        kind = (B >= (obs ? B200MPC_LANE_KERNEL_MIN_BATCH_OBS : B200MPC_LANE_KERNEL_MIN_BATCH)) ? B200MPC_KERNEL_LANE : B200MPC_KERNEL_WARP;
    return kind;
}

static int launch_solve(b200mpc_handle *h, const BatchArgs &a, cudaStream_t stream) {
    const int kind = choose_kernel(h, a.B);
    if (kind == B200MPC_KERNEL_LANE) return launch_solve_tpp(h, a, stream);
    h->last_kind = B200MPC_KERNEL_WARP;
    CU_TRY(h, cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned int), stream));
    int grid = h->ctas;
    const int need = (a.B + 3) / 4;
    if (need < grid) grid = need;
    if (grid < 1) grid = 1;
    // The scan form of the Riccati recursion shortens the critical path of ONE solve (-27 % latency) but does more
    // total work: it is used while the batch leaves warps of the persistent grid idle, i.e. when latency is what
    // the caller sees; a batch that fills the machine keeps the lane-parallel serial form (+17 % solves/s there).
    const bool scan = (h->kkt_scan_mode >= 0) ? (h->kkt_scan_mode == 1) : (a.B <= h->ctas * 4);
    CU_TRY(h, cudaEventRecord(h->ev0, stream));
    // A batch of more than two waves of the persistent grid goes as a cascade of iteration-bounded launches: launch p hands
    // the problems that reach warp_caps[p] iterations to launch p + 1 (records in two alternating buffers, counts on the
    // device), so the long problems run side by side in the last launches instead of one behind the other at the end.
    const int ncaps = (a.B > 2 * h->ctas * 4) ? h->n_warp_caps : 0;
    BatchArgs cur = a;
    for (int pass = 0; pass <= ncaps; pass++) {
        if (pass < ncaps) {
            const int w = pass & 1;
            const int rc = ensure_hand_buffer(h, w, (size_t)a.B, stream);
            if (rc) return rc;
            cur.hand_rec = h->d_hand[w]; cur.hand_count = h->d_hand_count + w; cur.hand_cap = (int)h->hand_cap[w];
            cur.hand_iter = h->warp_caps[pass];
        } else {
            cur.hand_rec = nullptr; cur.hand_count = nullptr; cur.hand_cap = 0; cur.hand_iter = 0;
        }
        if (pass > 0) CU_TRY(h, cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned int), stream));
        const int rc = launch_warp_kernel(h, cur, pass == 0 ? grid : h->ctas, scan, stream);
        if (rc) return rc;
        h->launches++;
        // the next launch resumes what this one exported
        cur.resume = 1; cur.resume_rec = cur.hand_rec; cur.resume_count = cur.hand_count; cur.resume_cap = cur.hand_cap;
    }
    CU_TRY(h, cudaEventRecord(h->ev1, stream));
    return 0;
}

static int launch_warp_kernel(b200mpc_handle *h, const BatchArgs &a, int grid, bool scan, cudaStream_t stream) {
    h->kp.kkt_scan = scan ? 1 : 0;
    switch (h->J) {
#define LAUNCH_WARP(JJ, SC)                                                                                \
    do {                                                                                                    \
        const size_t sm_ = SC ? h->smem_scan_bytes : h->smem_bytes;                                         \
        if (h->prm.obs_form != B200MPC_OBS_NONE) mpc_solve_kernel<JJ, true, SC><<<grid, 128, sm_, stream>>>(h->kp, a); \
        else mpc_solve_kernel<JJ, false, SC><<<grid, 128, sm_, stream>>>(h->kp, a);                        \
    } while (0)
        case 1: if (scan) LAUNCH_WARP(1, true); else LAUNCH_WARP(1, false); break;
        case 2: LAUNCH_WARP(2, false); break;
        case 3: LAUNCH_WARP(3, false); break;
        default: LAUNCH_WARP(4, false); break;
#undef LAUNCH_WARP
    }
    CU_TRY(h, cudaGetLastError());
    return 0;
}

extern "C" int b200mpc_solve_batch_device(b200mpc_handle *h, int B, const double *x0, const double *xref,
                                          const double *uref, const double *obs_x, const double *obs_y,
                                          int obs_stride, const double *u_init, double *X_out, double *U_out,
                                          double *cost_out, int32_t *status_out, int32_t *iters_out,
                                          int32_t *ls_out, void *stream) {
    int rc = check_args(h, B, x0, xref, uref, obs_x, obs_y, obs_stride);
    if (rc) return rc;
    if (B == 0) return 0;
    if (!X_out || !U_out || !status_out) return set_err(h, B200MPC_E_ARG, "X_out / U_out / status_out is NULL");
    CU_TRY(h, cudaSetDevice(h->device));
    BatchArgs a;
    a.B = B; a.obs_stride = obs_stride;
    a.x0 = x0; a.xref = xref; a.uref = uref; a.ox = obs_x; a.oy = obs_y; a.u_init = u_init;
    a.X = X_out; a.U = U_out; a.cost = cost_out; a.status = status_out; a.iters = iters_out; a.ls = ls_out;
    a.counter = h->d_counter;
    return launch_solve(h, a, (cudaStream_t)stream);
}

static int ensure_buf(b200mpc_handle *h, size_t bytes) {
    if (bytes <= h->d_cap) return 0;
    if (h->d_buf) { cudaFree(h->d_buf); h->d_buf = nullptr; h->d_cap = 0; }
    size_t cap = bytes + bytes / 4;
    cudaError_t e = cudaMalloc(&h->d_buf, cap);
    if (e != cudaSuccess) return set_err(h, B200MPC_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    h->d_cap = cap;
    return 0;
}

static size_t al256(size_t n) { return (n + 255) & ~(size_t)255; }

// page-locked (cudaHostAlloc / cudaHostRegister) host memory?  Only then do asynchronous copies overlap the kernel.
static bool is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

extern "C" int b200mpc_solve_batch(b200mpc_handle *h, int B, const double *x0, const double *xref, const double *uref,
                                   const double *obs_x, const double *obs_y, int obs_stride, const double *u_init,
                                   double *X_out, double *U_out, double *cost_out, int32_t *status_out,
                                   int32_t *iters_out, int32_t *ls_out) {
    int rc = check_args(h, B, x0, xref, uref, obs_x, obs_y, obs_stride);
    if (rc) return rc;
    if (B == 0) return 0;
    if (!X_out || !U_out || !status_out) return set_err(h, B200MPC_E_ARG, "X_out / U_out / status_out is NULL");
    CU_TRY(h, cudaSetDevice(h->device));
    const int N = h->prm.N, M = h->prm.M;
    const bool traj = h->prm.ref_kind == B200MPC_REF_TRAJ, obs = h->prm.obs_form != B200MPC_OBS_NONE;
    const size_t nb = (size_t)B;
    const size_t sz_x0 = al256(nb * 3 * 8), sz_xref = al256(nb * (traj ? 3 * N : 3) * 8);
    const size_t sz_uref = traj ? al256(nb * 2 * N * 8) : 0;
    const size_t n_obs = obs ? (obs_stride ? nb * obs_stride : (size_t)M) : 0;
    const size_t sz_obs = al256(n_obs * 8);
    const size_t sz_ui = u_init ? al256(nb * 2 * N * 8) : 0;
    const size_t sz_X = al256(nb * 3 * (N + 1) * 8), sz_U = al256(nb * 2 * N * 8), sz_c = al256(nb * 8), sz_i = al256(nb * 4);
    const size_t total = sz_x0 + sz_xref + sz_uref + 2 * sz_obs + sz_ui + sz_X + sz_U + sz_c + 3 * sz_i;
    h->last_streamed = 0;
    if (B <= 64 && total <= h->small_cap && !getenv("B200MPC_NO_ZEROCOPY")) {
        // ---- small batch (the single solve of a control step): zero-copy staging ----
        // The inputs are placed in page-locked, device-mapped memory and the kernel reads them and writes its results
        // over PCIe itself: one kernel launch and one synchronisation, no copy-engine round trips.
        char *hp = h->h_small;
        const ptrdiff_t dev_off = h->h_small_dev - h->h_small;
        auto put = [&](const void *src, size_t bytes, size_t slot) -> char * {
            char *r = hp;
            if (src) memcpy(r, src, bytes);
            hp += slot;
            return r + dev_off;
        };
        BatchArgs a;
        a.B = B; a.obs_stride = obs_stride;
        a.x0 = (const double *)put(x0, nb * 3 * 8, sz_x0);
        a.xref = (const double *)put(xref, nb * (traj ? 3 * N : 3) * 8, sz_xref);
        a.uref = traj ? (const double *)put(uref, nb * 2 * N * 8, sz_uref) : nullptr;
        a.ox = obs ? (const double *)put(obs_x, n_obs * 8, sz_obs) : nullptr;
        a.oy = obs ? (const double *)put(obs_y, n_obs * 8, sz_obs) : nullptr;
        a.u_init = u_init ? (const double *)put(u_init, nb * 2 * N * 8, sz_ui) : nullptr;
        char *o_X = hp; a.X = (double *)put(nullptr, 0, sz_X);
        char *o_U = hp; a.U = (double *)put(nullptr, 0, sz_U);
        char *o_c = hp; a.cost = (double *)put(nullptr, 0, sz_c);
        char *o_st = hp; a.status = (int *)put(nullptr, 0, sz_i);
        char *o_it = hp; a.iters = (int *)put(nullptr, 0, sz_i);
        char *o_ls = hp; a.ls = (int *)put(nullptr, 0, sz_i);
        a.counter = h->d_counter;
        rc = launch_solve(h, a, h->stream);
        if (rc) return rc;
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        memcpy(X_out, o_X, nb * 3 * (N + 1) * 8);
        memcpy(U_out, o_U, nb * 2 * N * 8);
        if (cost_out) memcpy(cost_out, o_c, nb * 8);
        memcpy(status_out, o_st, nb * 4);
        if (iters_out) memcpy(iters_out, o_it, nb * 4);
        if (ls_out) memcpy(ls_out, o_ls, nb * 4);
        return 0;
    }
    rc = ensure_buf(h, total);
    if (rc) return rc;
    char *p = h->d_buf;
    auto take = [&](size_t n) { char *r = p; p += n; return r; };
    double *d_x0 = (double *)take(sz_x0), *d_xref = (double *)take(sz_xref);
    double *d_uref = traj ? (double *)take(sz_uref) : nullptr;
    double *d_ox = obs ? (double *)take(sz_obs) : nullptr, *d_oy = obs ? (double *)take(sz_obs) : nullptr;
    double *d_ui = u_init ? (double *)take(sz_ui) : nullptr;
    double *d_X = (double *)take(sz_X), *d_U = (double *)take(sz_U), *d_c = (double *)take(sz_c);
    int *d_st = (int *)take(sz_i), *d_it = (int *)take(sz_i), *d_ls = (int *)take(sz_i);
    cudaStream_t s = h->stream;
    // every buffer the streamed path touches must be page-locked: a pageable one turns its copies into staged, effectively
    // synchronous ones, and the overlap (and the ordering the watermark relies on) is gone
    const bool all_pinned = is_pinned(x0) && is_pinned(xref) && (!traj || is_pinned(uref)) && (!u_init || is_pinned(u_init)) &&
                            is_pinned(X_out) && is_pinned(U_out) && is_pinned(status_out) && (!cost_out || is_pinned(cost_out)) &&
                            (!iters_out || is_pinned(iters_out)) && (!ls_out || is_pinned(ls_out));
    if (choose_kernel(h, B) == B200MPC_KERNEL_LANE && B >= B200MPC_STREAM_MIN_BATCH && !getenv("B200MPC_NO_STREAMING") && !obs &&
        all_pinned) {
        // ---- streamed solve: inputs arrive and results leave in chunks while the persistent kernel runs ----
        // copy stream:   [chunk c inputs H2D][avail := end of chunk c] ...          (the kernel waits for `avail`)
        // kernel:        finishes problems in roughly ascending order; the lane completing chunk c raises flags[c]
        // this thread:   polls flags[c] in host memory and enqueues the chunk's D2H copies on the output stream
        int nchunks = B / 32768;
        if (nchunks > B200MPC_MAX_CHUNKS) nchunks = B200MPC_MAX_CHUNKS;
        if (nchunks < 2) nchunks = 2;
        int chunk = (B + nchunks - 1) / nchunks;
        chunk = (chunk + 255) & ~255;
        nchunks = (B + chunk - 1) / chunk;
        volatile unsigned int *flags = h->h_sync;
        unsigned int *marks = h->h_sync + B200MPC_MAX_CHUNKS;
        volatile unsigned int *abort_word = h->h_sync + 2 * B200MPC_MAX_CHUNKS; // the kernel's wait on the watermark gives up when set
        *abort_word = 0;
        // any failure from here on: release the persistent kernel (it may be waiting for inputs that will never arrive), drain
        // the three streams so that nothing of this call is in flight when the staging buffer is reused, then report
        auto bail = [&](int code, const std::string &msg) {
            *abort_word = 1;
            cudaStreamSynchronize(s); cudaStreamSynchronize(h->cstream); cudaStreamSynchronize(h->ostream);
            cudaGetLastError();
            return set_err(h, code, msg);
        };
#define CU_TRY_S(call)                                                                                  \
    do {                                                                                                \
        const cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) return bail(B200MPC_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
        for (int c = 0; c < nchunks; c++) {
            flags[c] = 0;
            const long long end = (long long)(c + 1) * chunk;
            marks[c] = (unsigned)(end < B ? end : B);
        }
        CU_TRY_S(cudaMemsetAsync(h->d_sync, 0, (1 + B200MPC_MAX_CHUNKS) * sizeof(unsigned int), s));
        CU_TRY_S(cudaEventRecord(h->ev_sync, s));
        CU_TRY_S(cudaStreamWaitEvent(h->cstream, h->ev_sync, 0));
        CU_TRY_S(cudaStreamWaitEvent(h->ostream, h->ev_sync, 0)); // earlier work on the staging buffer is complete
        const size_t w_ref = traj ? 3 * N : 3;
        for (int c = 0; c < nchunks; c++) {
            const size_t b0 = (size_t)c * chunk, n = marks[c] - b0;
            cudaStream_t cs = h->cstream;
            CU_TRY_S(cudaMemcpyAsync(d_x0 + b0 * 3, x0 + b0 * 3, n * 3 * 8, cudaMemcpyHostToDevice, cs));
            CU_TRY_S(cudaMemcpyAsync(d_xref + b0 * w_ref, xref + b0 * w_ref, n * w_ref * 8, cudaMemcpyHostToDevice, cs));
            if (traj) CU_TRY_S(cudaMemcpyAsync(d_uref + b0 * 2 * N, uref + b0 * 2 * N, n * 2 * N * 8, cudaMemcpyHostToDevice, cs));
            if (u_init) CU_TRY_S(cudaMemcpyAsync(d_ui + b0 * 2 * N, u_init + b0 * 2 * N, n * 2 * N * 8, cudaMemcpyHostToDevice, cs));
            CU_TRY_S(cudaMemcpyAsync(h->d_sync, marks + c, sizeof(unsigned int), cudaMemcpyHostToDevice, cs));
        }
        BatchArgs a;
        a.B = B; a.obs_stride = obs_stride;
        a.x0 = d_x0; a.xref = d_xref; a.uref = d_uref; a.ox = d_ox; a.oy = d_oy; a.u_init = d_ui;
        a.X = d_X; a.U = d_U; a.cost = d_c; a.status = d_st; a.iters = d_it; a.ls = d_ls;
        a.counter = h->d_counter;
        StreamWords sw;
        sw.avail = h->d_sync; sw.done = h->d_sync + 1; sw.flags = h->h_sync_dev; sw.chunk = chunk;
        sw.abort = h->h_sync_dev + 2 * B200MPC_MAX_CHUNKS;
        rc = launch_solve_tpp(h, a, s, &sw);
        if (rc) return bail(rc, h->err);
        for (int c = 0; c < nchunks; c++) {
            unsigned spins = 0;
            while (!flags[c]) {
                if ((++spins & 0x3ff) == 0) {
                    const cudaError_t q = cudaStreamQuery(s);
                    if (q == cudaSuccess) break; // kernel finished: every chunk is complete
                    if (q != cudaErrorNotReady) return bail(B200MPC_E_CUDA, std::string("solve kernel: ") + cudaGetErrorString(q));
                    const cudaError_t qc = cudaStreamQuery(h->cstream); // a failed input copy would leave the kernel waiting
                    if (qc != cudaSuccess && qc != cudaErrorNotReady)
                        return bail(B200MPC_E_CUDA, std::string("input copy: ") + cudaGetErrorString(qc));
                    std::this_thread::yield();
                }
            }
            const size_t b0 = (size_t)c * chunk, n = marks[c] - b0;
            cudaStream_t os = h->ostream;
            CU_TRY_S(cudaMemcpyAsync(X_out + b0 * 3 * (N + 1), d_X + b0 * 3 * (N + 1), n * 3 * (N + 1) * 8, cudaMemcpyDeviceToHost, os));
            CU_TRY_S(cudaMemcpyAsync(U_out + b0 * 2 * N, d_U + b0 * 2 * N, n * 2 * N * 8, cudaMemcpyDeviceToHost, os));
            if (cost_out) CU_TRY_S(cudaMemcpyAsync(cost_out + b0, d_c + b0, n * 8, cudaMemcpyDeviceToHost, os));
            CU_TRY_S(cudaMemcpyAsync(status_out + b0, d_st + b0, n * 4, cudaMemcpyDeviceToHost, os));
            if (iters_out) CU_TRY_S(cudaMemcpyAsync(iters_out + b0, d_it + b0, n * 4, cudaMemcpyDeviceToHost, os));
            if (ls_out) CU_TRY_S(cudaMemcpyAsync(ls_out + b0, d_ls + b0, n * 4, cudaMemcpyDeviceToHost, os));
        }
        CU_TRY_S(cudaStreamSynchronize(s));
        CU_TRY_S(cudaStreamSynchronize(h->cstream));
        CU_TRY_S(cudaStreamSynchronize(h->ostream));
        h->last_streamed = nchunks;
        return 0;
#undef CU_TRY_S
    }
    CU_TRY(h, cudaMemcpyAsync(d_x0, x0, nb * 3 * 8, cudaMemcpyHostToDevice, s));
    CU_TRY(h, cudaMemcpyAsync(d_xref, xref, nb * (traj ? 3 * N : 3) * 8, cudaMemcpyHostToDevice, s));
    if (traj) CU_TRY(h, cudaMemcpyAsync(d_uref, uref, nb * 2 * N * 8, cudaMemcpyHostToDevice, s));
    if (obs) {
        CU_TRY(h, cudaMemcpyAsync(d_ox, obs_x, n_obs * 8, cudaMemcpyHostToDevice, s));
        CU_TRY(h, cudaMemcpyAsync(d_oy, obs_y, n_obs * 8, cudaMemcpyHostToDevice, s));
    }
    if (u_init) CU_TRY(h, cudaMemcpyAsync(d_ui, u_init, nb * 2 * N * 8, cudaMemcpyHostToDevice, s));
    BatchArgs a;
    a.B = B; a.obs_stride = obs_stride;
    a.x0 = d_x0; a.xref = d_xref; a.uref = d_uref; a.ox = d_ox; a.oy = d_oy; a.u_init = d_ui;
    a.X = d_X; a.U = d_U; a.cost = d_c; a.status = d_st; a.iters = d_it; a.ls = d_ls;
    a.counter = h->d_counter;
    rc = launch_solve(h, a, s);
    if (rc) return rc;
    CU_TRY(h, cudaMemcpyAsync(X_out, d_X, nb * 3 * (N + 1) * 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(h, cudaMemcpyAsync(U_out, d_U, nb * 2 * N * 8, cudaMemcpyDeviceToHost, s));
    if (cost_out) CU_TRY(h, cudaMemcpyAsync(cost_out, d_c, nb * 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(h, cudaMemcpyAsync(status_out, d_st, nb * 4, cudaMemcpyDeviceToHost, s));
    if (iters_out) CU_TRY(h, cudaMemcpyAsync(iters_out, d_it, nb * 4, cudaMemcpyDeviceToHost, s));
    if (ls_out) CU_TRY(h, cudaMemcpyAsync(ls_out, d_ls, nb * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(h, cudaStreamSynchronize(s));
    return 0;
}

extern "C" int b200mpc_eval_batch(b200mpc_handle *h, int B, const double *x0, const double *xref, const double *uref,
                                  const double *obs_x, const double *obs_y, int obs_stride, const double *X,
                                  const double *U, const double *lam, double obj_scale, double *f_out, double *c_out,
                                  double *grad_out, double *stage_out) {
    int rc = check_args(h, B, x0, xref, uref, obs_x, obs_y, obs_stride);
    if (rc) return rc;
    if (B == 0) return 0;
    if (!X || !U) return set_err(h, B200MPC_E_ARG, "X / U is NULL");
    CU_TRY(h, cudaSetDevice(h->device));
    const int N = h->prm.N, M = h->prm.M;
    const bool traj = h->prm.ref_kind == B200MPC_REF_TRAJ, obs = h->prm.obs_form != B200MPC_OBS_NONE;
    const size_t nb = (size_t)B;
    const size_t sz_x0 = al256(nb * 3 * 8), sz_xref = al256(nb * (traj ? 3 * N : 3) * 8);
    const size_t sz_uref = traj ? al256(nb * 2 * N * 8) : 0;
    const size_t n_obs = obs ? (obs_stride ? nb * obs_stride : (size_t)M) : 0;
    const size_t sz_obs = al256(n_obs * 8);
    const size_t sz_X = al256(nb * 3 * (N + 1) * 8), sz_U = al256(nb * 2 * N * 8), sz_l = al256(nb * 3 * N * 8);
    const size_t sz_f = al256(nb * 8), sz_g = al256(nb * 5 * N * 8), sz_s = al256(nb * (N + 1) * 36 * 8);
    const size_t total = sz_x0 + sz_xref + sz_uref + 2 * sz_obs + sz_X + sz_U + sz_l + sz_f + sz_l + sz_g + sz_s;
    rc = ensure_buf(h, total);
    if (rc) return rc;
    char *p = h->d_buf;
    auto take = [&](size_t n) { char *r = p; p += n; return r; };
    double *d_x0 = (double *)take(sz_x0), *d_xref = (double *)take(sz_xref);
    double *d_uref = traj ? (double *)take(sz_uref) : nullptr;
    double *d_ox = obs ? (double *)take(sz_obs) : nullptr, *d_oy = obs ? (double *)take(sz_obs) : nullptr;
    double *d_X = (double *)take(sz_X), *d_U = (double *)take(sz_U), *d_lam = (double *)take(sz_l);
    double *d_f = (double *)take(sz_f), *d_c = (double *)take(sz_l), *d_g = (double *)take(sz_g), *d_s = (double *)take(sz_s);
    cudaStream_t s = h->stream;
    CU_TRY(h, cudaMemcpyAsync(d_x0, x0, nb * 3 * 8, cudaMemcpyHostToDevice, s));
    CU_TRY(h, cudaMemcpyAsync(d_xref, xref, nb * (traj ? 3 * N : 3) * 8, cudaMemcpyHostToDevice, s));
    if (traj) CU_TRY(h, cudaMemcpyAsync(d_uref, uref, nb * 2 * N * 8, cudaMemcpyHostToDevice, s));
    if (obs) {
        CU_TRY(h, cudaMemcpyAsync(d_ox, obs_x, n_obs * 8, cudaMemcpyHostToDevice, s));
        CU_TRY(h, cudaMemcpyAsync(d_oy, obs_y, n_obs * 8, cudaMemcpyHostToDevice, s));
    }
    CU_TRY(h, cudaMemcpyAsync(d_X, X, nb * 3 * (N + 1) * 8, cudaMemcpyHostToDevice, s));
    CU_TRY(h, cudaMemcpyAsync(d_U, U, nb * 2 * N * 8, cudaMemcpyHostToDevice, s));
    if (lam) CU_TRY(h, cudaMemcpyAsync(d_lam, lam, nb * 3 * N * 8, cudaMemcpyHostToDevice, s));
    EvalArgs a;
    a.B = B; a.obs_stride = obs_stride;
    a.x0 = d_x0; a.xref = d_xref; a.uref = d_uref; a.ox = d_ox; a.oy = d_oy;
    a.X = d_X; a.U = d_U; a.lam = lam ? d_lam : nullptr; a.obj_scale = obj_scale;
    a.f = d_f; a.c = d_c; a.grad = d_g; a.stage = d_s;
    const int grid = (B + 3) / 4;
    switch (h->J) {
        case 1: mpc_eval_kernel<1><<<grid, 128, h->smem_bytes, s>>>(h->kp, a); break;
        case 2: mpc_eval_kernel<2><<<grid, 128, h->smem_bytes, s>>>(h->kp, a); break;
        case 3: mpc_eval_kernel<3><<<grid, 128, h->smem_bytes, s>>>(h->kp, a); break;
        default: mpc_eval_kernel<4><<<grid, 128, h->smem_bytes, s>>>(h->kp, a); break;
    }
    CU_TRY(h, cudaGetLastError());
    h->launches++;
    if (f_out) CU_TRY(h, cudaMemcpyAsync(f_out, d_f, nb * 8, cudaMemcpyDeviceToHost, s));
    if (c_out) CU_TRY(h, cudaMemcpyAsync(c_out, d_c, nb * 3 * N * 8, cudaMemcpyDeviceToHost, s));
    if (grad_out) CU_TRY(h, cudaMemcpyAsync(grad_out, d_g, nb * 5 * N * 8, cudaMemcpyDeviceToHost, s));
    if (stage_out) CU_TRY(h, cudaMemcpyAsync(stage_out, d_s, nb * (N + 1) * 36 * 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(h, cudaStreamSynchronize(s));
    return 0;
}

// ---- obstacle-list construction (get_obstacles, scripts/point_follower_local_planner.py:88-118) -------------------
static int launch_obstacles(b200mpc_handle *h, int B, int n_beams, const double *scan, const double *bcos,
                            const double *bsin, const double *pos, const double *yaw, double size, double resolution,
                            int slots, double *ox, double *oy, int32_t *count, cudaStream_t stream) {
    ObsBuildArgs a;
    const double map_size = size * 2.0; // the caller passes costmap_size; the grid spans 2*size (reference :89)
    a.B = B; a.n = n_beams; a.slots = slots;
    a.nc = (int)(map_size / resolution); // int(map_size / cell_size), utils.py:13
    if (a.nc < 1 || a.nc > 256) return set_err(h, B200MPC_E_ARG, "grid side int(2*size/resolution) out of range [1,256]");
    a.nwords = (a.nc * a.nc + 31) / 32;
    a.scan = scan; a.bcos = bcos; a.bsin = bsin; a.pos = pos; a.yaw = yaw;
    a.half = map_size / 2; a.res = resolution; a.origin = (double)(a.nc / 2) * resolution;
    a.ox = ox; a.oy = oy; a.count = count;
    const size_t per_warp = (((size_t)a.nwords * 8 + (size_t)slots * 4) + 15) & ~(size_t)15;
    const size_t smem = per_warp * OBS_WARPS + 16;
    if (smem > 200 * 1024) return set_err(h, B200MPC_E_ARG, "grid / slots too large for the shared-memory staging");
    if (smem > 48 * 1024) CU_TRY(h, cudaFuncSetAttribute(obstacles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int grid = (B + OBS_WARPS - 1) / OBS_WARPS;
    const int cap = h->sm_count * 8; // grid-stride above 8 CTAs per SM
    if (grid > cap) grid = cap;
    CU_TRY(h, cudaEventRecord(h->ev0, stream));
    obstacles_kernel<<<grid, OBS_WARPS * 32, smem, stream>>>(a);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaEventRecord(h->ev1, stream));
    h->launches++;
    return 0;
}

static int check_obstacle_args(b200mpc_handle *h, int B, int n_beams, const void *scan, const void *bcos, const void *bsin,
                               const void *pos, const void *yaw, double size, double resolution, int slots,
                               const void *ox, const void *oy) {
    if (!h) return B200MPC_E_ARG;
    if (B < 0 || n_beams < 1) return set_err(h, B200MPC_E_ARG, "B < 0 or n_beams < 1");
    if (slots < 1 || slots > B200MPC_MAX_M) return set_err(h, B200MPC_E_ARG, "slots out of range [1,1024]");
    if (!(size > 0) || !(resolution > 0)) return set_err(h, B200MPC_E_ARG, "size and resolution must be > 0");
    if (B > 0 && (!scan || !bcos || !bsin || !pos || !yaw || !ox || !oy)) return set_err(h, B200MPC_E_ARG, "NULL argument");
    return 0;
}

extern "C" int b200mpc_obstacles_batch_device(b200mpc_handle *h, int B, int n_beams, const double *scan,
                                              const double *beam_cos, const double *beam_sin, const double *pos,
                                              const double *yaw, double size, double resolution, int slots,
                                              double *obs_x, double *obs_y, int32_t *count, void *stream) {
    int rc = check_obstacle_args(h, B, n_beams, scan, beam_cos, beam_sin, pos, yaw, size, resolution, slots, obs_x, obs_y);
    if (rc) return rc;
    if (B == 0) return 0;
    CU_TRY(h, cudaSetDevice(h->device));
    return launch_obstacles(h, B, n_beams, scan, beam_cos, beam_sin, pos, yaw, size, resolution, slots, obs_x, obs_y, count,
                            (cudaStream_t)stream);
}

extern "C" int b200mpc_obstacles_batch(b200mpc_handle *h, int B, int n_beams, const double *scan, const double *beam_cos,
                                       const double *beam_sin, const double *pos, const double *yaw, double size,
                                       double resolution, int slots, double *obs_x, double *obs_y, int32_t *count) {
    int rc = check_obstacle_args(h, B, n_beams, scan, beam_cos, beam_sin, pos, yaw, size, resolution, slots, obs_x, obs_y);
    if (rc) return rc;
    if (B == 0) return 0;
    CU_TRY(h, cudaSetDevice(h->device));
    const size_t nb = (size_t)B;
    const size_t sz_scan = al256(nb * n_beams * 8), sz_tab = al256((size_t)n_beams * 8), sz_pos = al256(nb * 2 * 8);
    const size_t sz_yaw = al256(nb * 8), sz_o = al256(nb * slots * 8), sz_c = al256(nb * 4);
    rc = ensure_buf(h, sz_scan + 2 * sz_tab + sz_pos + sz_yaw + 2 * sz_o + sz_c);
    if (rc) return rc;
    char *p = h->d_buf;
    auto take = [&](size_t n) { char *r = p; p += n; return r; };
    double *d_scan = (double *)take(sz_scan), *d_c = (double *)take(sz_tab), *d_s = (double *)take(sz_tab);
    double *d_pos = (double *)take(sz_pos), *d_yaw = (double *)take(sz_yaw);
    double *d_ox = (double *)take(sz_o), *d_oy = (double *)take(sz_o);
    int *d_cnt = (int *)take(sz_c);
    cudaStream_t s = h->stream;
    CU_TRY(h, cudaMemcpyAsync(d_scan, scan, nb * n_beams * 8, cudaMemcpyHostToDevice, s));
    CU_TRY(h, cudaMemcpyAsync(d_c, beam_cos, (size_t)n_beams * 8, cudaMemcpyHostToDevice, s));
    CU_TRY(h, cudaMemcpyAsync(d_s, beam_sin, (size_t)n_beams * 8, cudaMemcpyHostToDevice, s));
    CU_TRY(h, cudaMemcpyAsync(d_pos, pos, nb * 2 * 8, cudaMemcpyHostToDevice, s));
    CU_TRY(h, cudaMemcpyAsync(d_yaw, yaw, nb * 8, cudaMemcpyHostToDevice, s));
    rc = launch_obstacles(h, B, n_beams, d_scan, d_c, d_s, d_pos, d_yaw, size, resolution, slots, d_ox, d_oy, d_cnt, s);
    if (rc) return rc;
    CU_TRY(h, cudaMemcpyAsync(obs_x, d_ox, nb * slots * 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(h, cudaMemcpyAsync(obs_y, d_oy, nb * slots * 8, cudaMemcpyDeviceToHost, s));
    if (count) CU_TRY(h, cudaMemcpyAsync(count, d_cnt, nb * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(h, cudaStreamSynchronize(s));
    return 0;
}

// ---- reference producers (get_goal_for_mpc / get_reference_trajectory) ---------------------------------------------
static int refgen_grid(const b200mpc_handle *h, int B) {
    int grid = (B + REFGEN_WARPS - 1) / REFGEN_WARPS;
    const int cap = h->sm_count * 8;
    return grid > cap ? cap : (grid < 1 ? 1 : grid);
}

extern "C" int b200mpc_goals_batch_device(b200mpc_handle *h, int B, int K, const double *path_xy, const double *path_heading,
                                          int per_robot_paths, const double *goal, const double *pos, int pos_stride,
                                          double lookahead, double *goal_out, int32_t *index_out, void *stream) {
    if (!h) return B200MPC_E_ARG;
    if (B < 0 || K < 1 || pos_stride < 2) return set_err(h, B200MPC_E_ARG, "B < 0, K < 1 or pos_stride < 2");
    if (B == 0) return 0;
    if (!path_xy || !path_heading || !goal || !pos || !goal_out) return set_err(h, B200MPC_E_ARG, "NULL argument");
    CU_TRY(h, cudaSetDevice(h->device));
    RefGenArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.K = K; a.N = h->prm.N; a.Ko = K;
    a.path_stride = per_robot_paths ? K : 0;
    a.path_xy = path_xy; a.heading = path_heading;
    a.goal = goal; a.goal_stride = 5; a.pos = pos; a.pos_stride = pos_stride; a.lookahead = lookahead;
    a.out_goal = goal_out; a.nearest = index_out;
    goals_kernel<<<refgen_grid(h, B), REFGEN_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    CU_TRY(h, cudaGetLastError());
    h->launches++;
    return 0;
}

extern "C" int b200mpc_reftraj_batch_device(b200mpc_handle *h, int B, int K, const double *path_xy,
                                            const double *path_heading, const double *path_velocity,
                                            const double *path_omega, int n_omega, int per_robot_paths, const double *x0,
                                            const double *goal, double *pxf_out, double *puf_out, int32_t *index_out,
                                            void *stream) {
    if (!h) return B200MPC_E_ARG;
    if (B < 0 || K < 1 || n_omega < 1 || n_omega > K) return set_err(h, B200MPC_E_ARG, "B < 0, K < 1 or n_omega outside [1,K]");
    if (B == 0) return 0;
    if (!path_xy || !path_heading || !path_velocity || !path_omega || !x0 || !goal || !pxf_out || !puf_out)
        return set_err(h, B200MPC_E_ARG, "NULL argument");
    CU_TRY(h, cudaSetDevice(h->device));
    RefGenArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.K = K; a.N = h->prm.N; a.Ko = n_omega;
    a.path_stride = per_robot_paths ? K : 0;
    a.path_xy = path_xy; a.heading = path_heading; a.velocity = path_velocity; a.omega = path_omega;
    a.goal = goal; a.goal_stride = 3; a.pos = x0; a.pos_stride = 3;
    a.pxf = pxf_out; a.puf = puf_out; a.nearest = index_out;
    reftraj_kernel<<<refgen_grid(h, B), REFGEN_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    CU_TRY(h, cudaGetLastError());
    h->launches++;
    return 0;
}

// host-buffer variants: stage through the handle's device buffer, blocking
struct HostStage {
    b200mpc_handle *h;
    char *p;
    size_t used;
};
static size_t stage_size(size_t n) { return al256(n); }
template <class T>
static T *stage_in(HostStage &st, const T *src, size_t count, cudaError_t &e) {
    T *d = (T *)(st.p + st.used);
    st.used += stage_size(count * sizeof(T));
    if (src && e == cudaSuccess) e = cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyHostToDevice, st.h->stream);
    return d;
}

extern "C" int b200mpc_goals_batch(b200mpc_handle *h, int B, int K, const double *path_xy, const double *path_heading,
                                   int per_robot_paths, const double *goal, const double *pos, double lookahead,
                                   double *goal_out, int32_t *index_out) {
    if (!h) return B200MPC_E_ARG;
    if (B < 0 || K < 1) return set_err(h, B200MPC_E_ARG, "B < 0 or K < 1");
    if (B == 0) return 0;
    if (!path_xy || !path_heading || !goal || !pos || !goal_out) return set_err(h, B200MPC_E_ARG, "NULL argument");
    CU_TRY(h, cudaSetDevice(h->device));
    const size_t nb = (size_t)B, np = per_robot_paths ? nb * K : (size_t)K;
    int rc = ensure_buf(h, stage_size(np * 16) + stage_size(np * 8) + stage_size(nb * 40) + stage_size(nb * 16) +
                               stage_size(nb * 24) + stage_size(nb * 4));
    if (rc) return rc;
    HostStage st{h, h->d_buf, 0};
    cudaError_t e = cudaSuccess;
    double *d_xy = stage_in(st, path_xy, np * 2, e), *d_h = stage_in(st, path_heading, np, e);
    double *d_g = stage_in(st, goal, nb * 5, e), *d_p = stage_in(st, pos, nb * 2, e);
    double *d_o = stage_in<double>(st, nullptr, nb * 3, e);
    int32_t *d_i = stage_in<int32_t>(st, nullptr, nb, e);
    CU_TRY(h, e);
    rc = b200mpc_goals_batch_device(h, B, K, d_xy, d_h, per_robot_paths, d_g, d_p, 2, lookahead, d_o, d_i, h->stream);
    if (rc) return rc;
    CU_TRY(h, cudaMemcpyAsync(goal_out, d_o, nb * 24, cudaMemcpyDeviceToHost, h->stream));
    if (index_out) CU_TRY(h, cudaMemcpyAsync(index_out, d_i, nb * 4, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int b200mpc_reftraj_batch(b200mpc_handle *h, int B, int K, const double *path_xy, const double *path_heading,
                                     const double *path_velocity, const double *path_omega, int n_omega,
                                     int per_robot_paths, const double *x0, const double *goal, double *pxf_out,
                                     double *puf_out, int32_t *index_out) {
    if (!h) return B200MPC_E_ARG;
    if (B < 0 || K < 1 || n_omega < 1 || n_omega > K) return set_err(h, B200MPC_E_ARG, "B < 0, K < 1 or n_omega outside [1,K]");
    if (B == 0) return 0;
    if (!path_xy || !path_heading || !path_velocity || !path_omega || !x0 || !goal || !pxf_out || !puf_out)
        return set_err(h, B200MPC_E_ARG, "NULL argument");
    CU_TRY(h, cudaSetDevice(h->device));
    const int N = h->prm.N;
    const size_t nb = (size_t)B, np = per_robot_paths ? nb * K : (size_t)K, nw = per_robot_paths ? nb * n_omega : (size_t)n_omega;
    int rc = ensure_buf(h, stage_size(np * 16) + 2 * stage_size(np * 8) + stage_size(nw * 8) + 2 * stage_size(nb * 24) +
                               stage_size(nb * 3 * N * 8) + stage_size(nb * 2 * N * 8) + stage_size(nb * 4));
    if (rc) return rc;
    HostStage st{h, h->d_buf, 0};
    cudaError_t e = cudaSuccess;
    double *d_xy = stage_in(st, path_xy, np * 2, e), *d_h = stage_in(st, path_heading, np, e);
    double *d_v = stage_in(st, path_velocity, np, e), *d_w = stage_in(st, path_omega, nw, e);
    double *d_x0 = stage_in(st, x0, nb * 3, e), *d_g = stage_in(st, goal, nb * 3, e);
    double *d_pxf = stage_in<double>(st, nullptr, nb * 3 * N, e), *d_puf = stage_in<double>(st, nullptr, nb * 2 * N, e);
    int32_t *d_i = stage_in<int32_t>(st, nullptr, nb, e);
    CU_TRY(h, e);
    rc = b200mpc_reftraj_batch_device(h, B, K, d_xy, d_h, d_v, d_w, n_omega, per_robot_paths, d_x0, d_g, d_pxf, d_puf, d_i,
                                      h->stream);
    if (rc) return rc;
    CU_TRY(h, cudaMemcpyAsync(pxf_out, d_pxf, nb * 3 * N * 8, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaMemcpyAsync(puf_out, d_puf, nb * 2 * N * 8, cudaMemcpyDeviceToHost, h->stream));
    if (index_out) CU_TRY(h, cudaMemcpyAsync(index_out, d_i, nb * 4, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// ---- after the solve: limiter, goal logic, plant step, next measurement (scripts/point_follower_local_planner.py:196-231)
extern "C" int b200mpc_control_step_device(b200mpc_handle *h, int B, const double *U_sol, const int32_t *status,
                                           double *state, double *x0, double *u_last, const double *goal, int goal_stride,
                                           int32_t *goal_flag, double goal_threshold, double accel_limit, int quantise,
                                           double *cmd_out, double *u_next, void *stream) {
    if (!h) return B200MPC_E_ARG;
    if (B < 0 || goal_stride < 2) return set_err(h, B200MPC_E_ARG, "B < 0 or goal_stride < 2");
    if (B == 0) return 0;
    if (!U_sol || !state || !x0 || !u_last || !goal || !goal_flag || !cmd_out) return set_err(h, B200MPC_E_ARG, "NULL argument");
    CU_TRY(h, cudaSetDevice(h->device));
    ControlArgs a;
    a.B = B; a.N = h->prm.N; a.U = U_sol; a.status = status; a.state = state; a.x0 = x0; a.u_last = u_last;
    a.goal = goal; a.goal_stride = goal_stride; a.goal_flag = goal_flag; a.goal_threshold = goal_threshold;
    a.accel_limit = accel_limit; a.dt = h->prm.dt; a.quantise = quantise; a.cmd = cmd_out; a.u_next = u_next;
    control_step_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
    CU_TRY(h, cudaGetLastError());
    h->launches++;
    return 0;
}

// ---- costmap inflation / dilation ---------------------------------------------------------------
static void costmap_tile(int H, int W, int halo_h, int halo_w, size_t per_cell_extra, int &TH, int &TW, size_t &smem,
                         bool with_tmp) {
    // tile = as much of the grid as fits ~64 KB of shared memory (three CTAs per SM), rows first
    TW = W < 128 ? W : 128;
    TH = H < 64 ? H : 64;
    for (;;) {
        const size_t SH = (size_t)TH + halo_h, SW = (size_t)TW + halo_w;
        smem = SH * SW * (8 + per_cell_extra) + (with_tmp ? SH * (size_t)TW * 8 : 0);
        if (smem <= 64 * 1024 || (TH <= 8 && TW <= 16)) break;
        if (TH > 8) TH = (TH + 1) / 2; else TW = (TW + 1) / 2;
    }
}

extern "C" int b200mpc_dilate_batch_device(b200mpc_handle *h, int B, int H, int W, const double *grid, int kh, int kw,
                                           uint8_t *out, void *stream) {
    if (!h) return B200MPC_E_ARG;
    if (B < 0 || H < 1 || W < 1 || kh < 1 || kw < 1 || kh > 32 || kw > 32) return set_err(h, B200MPC_E_ARG, "B < 0, empty grid or kernel size outside [1,32]");
    if (B == 0) return 0;
    if (!grid || !out) return set_err(h, B200MPC_E_ARG, "NULL argument");
    CU_TRY(h, cudaSetDevice(h->device));
    DilateArgs a;
    a.B = B; a.H = H; a.W = W; a.kh = kh; a.kw = kw; a.in = grid; a.out = out;
    size_t smem;
    if (kh == 10 && kw == 10) {
        // the structuring element of both costmap publishers: strip kernel (integer maxima by doubling, registers)
        DilateStripArgs s;
        s.B = B; s.H = H; s.W = W; s.in = grid; s.out = out;
        auto layout = [&](int TH, int TW) {
            s.TH = TH; s.TW = TW;
            const int nsx = (TW + DIL_SEG - 1) / DIL_SEG, nsy = (TH + DIL_SEG - 1) / DIL_SEG;
            s.PS = (nsx * DIL_SEG + kw - 1) | 1;
            s.PT = TW | 1;
            s.src_rows = (TH == H && TW == W) ? H : TH + kh - 1;
            const size_t ints = (size_t)s.src_rows * s.PS + (size_t)(nsy * DIL_SEG + kh - 1) * s.PT;
            s.obuf_off = (int)((ints * 4 + 15) & ~(size_t)15);
            return (size_t)s.obuf_off + (((size_t)TH * TW + 15) & ~(size_t)15);
        };
        smem = layout(H, W);
        s.raw_bytes = 0; s.bar_off = 0;
        {
            // whole grid fetched by the TMA engine (bulk copy of H*W*8 contiguous bytes), two CTAs per SM
            const size_t raw = (((size_t)H * W * 8) + 127) & ~(size_t)127;
            const size_t total = raw + smem + 16;
            const char *tma_env = getenv("B200MPC_DILATE_TMA");
            if ((W & 1) == 0 && (reinterpret_cast<size_t>(grid) & 15) == 0 && total <= 113 * 1024 && !(tma_env && tma_env[0] == '0')) {
                s.raw_bytes = (int)raw;
                s.bar_off = (int)(raw + smem);
                s.obuf_off += (int)raw;
                CU_TRY(h, cudaFuncSetAttribute(dilate_strip_tma_kernel<10, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
                const int cap = h->sm_count * 2;
                CU_TRY(h, cudaEventRecord(h->ev0, (cudaStream_t)stream));
                dilate_strip_tma_kernel<10, 10><<<B < cap ? B : cap, DIL_THREADS, total, (cudaStream_t)stream>>>(s);
                CU_TRY(h, cudaGetLastError());
                CU_TRY(h, cudaEventRecord(h->ev1, (cudaStream_t)stream));
                h->launches++;
                return 0;
            }
        }
        if (smem > 72 * 1024) smem = layout(H < 80 ? H : 80, W < 80 ? W : 80);
        if (smem > 48 * 1024)
            CU_TRY(h, cudaFuncSetAttribute(dilate_strip_kernel<10, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        const long long tiles = (long long)B * ((H + s.TH - 1) / s.TH) * ((W + s.TW - 1) / s.TW);
        if (tiles >= (1ll << 31)) return set_err(h, B200MPC_E_ARG, "too many tiles for one launch (B x tiles per grid >= 2^31)");
        const long long cap = (long long)h->sm_count * 3;
        CU_TRY(h, cudaEventRecord(h->ev0, (cudaStream_t)stream));
        dilate_strip_kernel<10, 10><<<(int)(tiles < cap ? tiles : cap), DIL_THREADS, smem, (cudaStream_t)stream>>>(s);
        CU_TRY(h, cudaGetLastError());
        CU_TRY(h, cudaEventRecord(h->ev1, (cudaStream_t)stream));
        h->launches++;
        return 0;
    }
    const size_t whole = 2 * (size_t)H * W * sizeof(double);
    if (whole <= 110 * 1024 && (kh - 1) / 2 < H && (kw - 1) / 2 < W) {
        // the whole grid and its row maxima fit: one CTA per grid, the grid is read once (two CTAs per SM at 80 x 80)
        a.TH = H; a.TW = W;
        CU_TRY(h, cudaFuncSetAttribute(dilate_whole_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        const long long cap = (long long)h->sm_count * 2;
        CU_TRY(h, cudaEventRecord(h->ev0, (cudaStream_t)stream));
        dilate_whole_kernel<<<(int)(B < cap ? B : cap), COSTMAP_THREADS, whole, (cudaStream_t)stream>>>(a);
        CU_TRY(h, cudaGetLastError());
        CU_TRY(h, cudaEventRecord(h->ev1, (cudaStream_t)stream));
        h->launches++;
        return 0;
    }
    costmap_tile(H, W, kh - 1, kw - 1, 0, a.TH, a.TW, smem, true);
    if (smem > 48 * 1024) CU_TRY(h, cudaFuncSetAttribute(dilate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const long long tiles = (long long)B * ((H + a.TH - 1) / a.TH) * ((W + a.TW - 1) / a.TW);
    const long long cap = (long long)h->sm_count * 6;
    CU_TRY(h, cudaEventRecord(h->ev0, (cudaStream_t)stream));
    dilate_kernel<<<(int)(tiles < cap ? tiles : cap), COSTMAP_THREADS, smem, (cudaStream_t)stream>>>(a);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaEventRecord(h->ev1, (cudaStream_t)stream));
    h->launches++;
    return 0;
}

extern "C" int b200mpc_inflate_batch_device(b200mpc_handle *h, int B, int H, int W, const double *grid,
                                            const double *inflation_matrix, int cells_inflation, double *out, void *stream) {
    if (!h) return B200MPC_E_ARG;
    if (B < 0 || H < 1 || W < 1 || cells_inflation < 0 || cells_inflation > 32)
        return set_err(h, B200MPC_E_ARG, "B < 0, empty grid or cells_inflation outside [0,32]");
    if (B == 0) return 0;
    if (!grid || !inflation_matrix || !out) return set_err(h, B200MPC_E_ARG, "NULL argument");
    CU_TRY(h, cudaSetDevice(h->device));
    InflateArgs a;
    a.B = B; a.H = H; a.W = W; a.c = cells_inflation; a.in = grid; a.M = inflation_matrix; a.out = out;
    size_t smem;
    costmap_tile(H, W, 2 * cells_inflation, 2 * cells_inflation, 1, a.TH, a.TW, smem, false);
    const int n = 2 * cells_inflation + 1;
    smem += (size_t)n * n * 8 + 16;
    const long long tiles = (long long)B * ((H + a.TH - 1) / a.TH) * ((W + a.TW - 1) / a.TW);
    const long long cap = (long long)h->sm_count * 6;
    const char *ib_env = getenv("B200MPC_INFLATE_BITS");
    if (cells_inflation <= 15 && !(ib_env && ib_env[0] == '0')) {
        // sources as bit rows (a window row of <= 31 bits is one funnel shift)
        const size_t SH = (size_t)a.TH + 2 * cells_inflation, SW = (size_t)a.TW + 2 * cells_inflation;
        const size_t sb = SH * SW * 8 + (size_t)n * n * 8 + SH * ((SW + 31) / 32 + 1) * 4;
        if (sb > 48 * 1024) CU_TRY(h, cudaFuncSetAttribute(inflate_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CU_TRY(h, cudaEventRecord(h->ev0, (cudaStream_t)stream));
        inflate_bits_kernel<<<(int)(tiles < cap ? tiles : cap), COSTMAP_THREADS, sb, (cudaStream_t)stream>>>(a);
        CU_TRY(h, cudaGetLastError());
        CU_TRY(h, cudaEventRecord(h->ev1, (cudaStream_t)stream));
        h->launches++;
        return 0;
    }
    if (smem > 48 * 1024) CU_TRY(h, cudaFuncSetAttribute(inflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU_TRY(h, cudaEventRecord(h->ev0, (cudaStream_t)stream));
    inflate_kernel<<<(int)(tiles < cap ? tiles : cap), COSTMAP_THREADS, smem, (cudaStream_t)stream>>>(a);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaEventRecord(h->ev1, (cudaStream_t)stream));
    h->launches++;
    return 0;
}

extern "C" int b200mpc_local_costmap_batch_device(b200mpc_handle *h, int B, int n_beams, const double *scan,
                                                  const double *beam_cos, const double *beam_sin, const double *yaw, double size,
                                                  double resolution, int kh, int kw, uint8_t *out, void *stream) {
    if (!h) return B200MPC_E_ARG;
    if (B < 0 || n_beams < 1 || kh < 1 || kw < 1 || kh > 32 || kw > 32) return set_err(h, B200MPC_E_ARG, "B < 0, n_beams < 1 or kernel size outside [1,32]");
    if (!(size > 0) || !(resolution > 0)) return set_err(h, B200MPC_E_ARG, "size and resolution must be > 0");
    if (B == 0) return 0;
    if (!scan || !beam_cos || !beam_sin || !yaw || !out) return set_err(h, B200MPC_E_ARG, "NULL argument");
    CU_TRY(h, cudaSetDevice(h->device));
    LocalCostmapArgs a;
    const double map_size = size * 2.0; // the publisher passes map_size = costmap_size * 2 (local_costmap_publisher.py:30)
    a.B = B; a.n = n_beams; a.kh = kh; a.kw = kw;
    a.nc = (int)(map_size / resolution);
    if (a.nc < 1 || a.nc > 512) return set_err(h, B200MPC_E_ARG, "grid side int(2*size/resolution) out of range [1,512]");
    a.wpr = (a.nc + 31) / 32;
    a.scan = scan; a.bcos = beam_cos; a.bsin = beam_sin; a.yaw = yaw; a.half = map_size / 2; a.res = resolution;
    a.value = 100; a.out = out;
    const size_t smem = (size_t)LCM_WARPS * 2 * a.nc * a.wpr * 4 + 2048; // + the byte-expansion table of the compile-time instance
    if (smem > 200 * 1024) return set_err(h, B200MPC_E_ARG, "grid too large for the shared-memory staging");
    // the publishers' structuring element on a grid whose rows are whole 16-byte pieces: compile-time passes
    const bool fast = (kh == 10 && kw == 10 && a.nc % 16 == 0 && (reinterpret_cast<size_t>(out) & 15) == 0);
    void (*kern)(const LocalCostmapArgs) = fast ? local_costmap_kernel<10, 10> : local_costmap_kernel<0, 0>;
    if (smem > 48 * 1024) CU_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int grid = (B + LCM_WARPS - 1) / LCM_WARPS;
    const int cap = h->sm_count * 8;
    if (grid > cap) grid = cap;
    CU_TRY(h, cudaEventRecord(h->ev0, (cudaStream_t)stream));
    kern<<<grid, LCM_WARPS * 32, smem, (cudaStream_t)stream>>>(a);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaEventRecord(h->ev1, (cudaStream_t)stream));
    h->launches++;
    return 0;
}

// ---- sensor model and path preprocessing ------------------------------------------------------------------------------
extern "C" int b200mpc_raycast_batch_device(b200mpc_handle *h, int B, int n_beams, const uint32_t *occ_bits, int H, int W,
                                            double origin_x, double origin_y, double resolution, const double *pose,
                                            int pose_stride, double angle_min, double angle_max, double range_min,
                                            double range_max, double step, double *scan, void *stream) {
    if (!h) return B200MPC_E_ARG;
    if (B < 0 || n_beams < 1 || H < 1 || W < 1 || pose_stride < 3) return set_err(h, B200MPC_E_ARG, "B < 0, n_beams < 1, empty map or pose_stride < 3");
    if (!(resolution > 0) || !(step > 0) || !(range_max >= range_min)) return set_err(h, B200MPC_E_ARG, "resolution, step must be > 0 and range_max >= range_min");
    if (B == 0) return 0;
    if (!occ_bits || !pose || !scan) return set_err(h, B200MPC_E_ARG, "NULL argument");
    CU_TRY(h, cudaSetDevice(h->device));
    RaycastArgs a;
    a.B = B; a.n = n_beams; a.H = H; a.W = W; a.wpr = (W + 31) / 32;
    a.nsteps = (int)llrint((range_max - range_min) / step) + 1; // int(round((range_max - range_min) / step)) + 1
    a.occ_bits = occ_bits; a.pose = pose; a.pose_stride = pose_stride;
    a.angle_min = angle_min; a.angle_inc = angle_max - angle_min;
    a.range_min = range_min; a.range_max = range_max; a.step = step;
    a.ox = origin_x; a.oy = origin_y; a.res = resolution; a.scan = scan;
    const size_t smem = (size_t)H * a.wpr * 4;
    if (smem > 200 * 1024) return set_err(h, B200MPC_E_ARG, "map too large for the shared-memory staging (H * ceil(W/32) * 4 bytes <= 200 KB)");
    if (smem > 48 * 1024) CU_TRY(h, cudaFuncSetAttribute(raycast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int grid = (B + RAY_WARPS - 1) / RAY_WARPS;
    const int cap = h->sm_count * 4;
    if (grid > cap) grid = cap;
    CU_TRY(h, cudaEventRecord(h->ev0, (cudaStream_t)stream));
    raycast_kernel<<<grid, RAY_WARPS * 32, smem, (cudaStream_t)stream>>>(a);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaEventRecord(h->ev1, (cudaStream_t)stream));
    h->launches++;
    return 0;
}

extern "C" int b200mpc_headings_batch_device(b200mpc_handle *h, int P, int K, const double *path_xy, double dt, double *heading,
                                             double *velocity, double *omega, void *stream) {
    if (!h) return B200MPC_E_ARG;
    if (P < 0 || K < 2 || !(dt > 0)) return set_err(h, B200MPC_E_ARG, "P < 0, K < 2 or dt <= 0");
    if (P == 0) return 0;
    if (!path_xy || !heading || !velocity || !omega) return set_err(h, B200MPC_E_ARG, "NULL argument");
    CU_TRY(h, cudaSetDevice(h->device));
    HeadingsArgs a;
    a.P = P; a.K = K; a.path_xy = path_xy; a.dt = dt; a.heading = heading; a.velocity = velocity; a.omega = omega;
    long long blocks = ((long long)P * K + 255) / 256;
    const long long cap = (long long)h->sm_count * 8;
    headings_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(a);
    CU_TRY(h, cudaGetLastError());
    h->launches++;
    return 0;
}

// host-buffer variants of the five producers above: stage through the handle's device buffer, blocking
extern "C" int b200mpc_dilate_batch(b200mpc_handle *h, int B, int H, int W, const double *grid, int kh, int kw, uint8_t *out) {
    if (!h) return B200MPC_E_ARG;
    if (B < 0 || H < 1 || W < 1) return set_err(h, B200MPC_E_ARG, "B < 0 or empty grid");
    if (B == 0) return 0;
    if (!grid || !out) return set_err(h, B200MPC_E_ARG, "NULL argument");
    CU_TRY(h, cudaSetDevice(h->device));
    const size_t cells = (size_t)B * H * W;
    int rc = ensure_buf(h, stage_size(cells * 8) + stage_size(cells));
    if (rc) return rc;
    HostStage st{h, h->d_buf, 0};
    cudaError_t e = cudaSuccess;
    double *d_in = stage_in(st, grid, cells, e);
    uint8_t *d_out = stage_in<uint8_t>(st, nullptr, cells, e);
    CU_TRY(h, e);
    rc = b200mpc_dilate_batch_device(h, B, H, W, d_in, kh, kw, d_out, h->stream);
    if (rc) return rc;
    CU_TRY(h, cudaMemcpyAsync(out, d_out, cells, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int b200mpc_inflate_batch(b200mpc_handle *h, int B, int H, int W, const double *grid, const double *inflation_matrix,
                                     int cells_inflation, double *out) {
    if (!h) return B200MPC_E_ARG;
    if (B < 0 || H < 1 || W < 1 || cells_inflation < 0) return set_err(h, B200MPC_E_ARG, "B < 0, empty grid or cells_inflation < 0");
    if (B == 0) return 0;
    if (!grid || !inflation_matrix || !out) return set_err(h, B200MPC_E_ARG, "NULL argument");
    CU_TRY(h, cudaSetDevice(h->device));
    const size_t cells = (size_t)B * H * W, nm = (size_t)(2 * cells_inflation + 1) * (2 * cells_inflation + 1);
    int rc = ensure_buf(h, 2 * stage_size(cells * 8) + stage_size(nm * 8));
    if (rc) return rc;
    HostStage st{h, h->d_buf, 0};
    cudaError_t e = cudaSuccess;
    double *d_in = stage_in(st, grid, cells, e), *d_m = stage_in(st, inflation_matrix, nm, e);
    double *d_out = stage_in<double>(st, nullptr, cells, e);
    CU_TRY(h, e);
    rc = b200mpc_inflate_batch_device(h, B, H, W, d_in, d_m, cells_inflation, d_out, h->stream);
    if (rc) return rc;
    CU_TRY(h, cudaMemcpyAsync(out, d_out, cells * 8, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int b200mpc_local_costmap_batch(b200mpc_handle *h, int B, int n_beams, const double *scan, const double *beam_cos,
                                           const double *beam_sin, const double *yaw, double size, double resolution, int kh,
                                           int kw, uint8_t *out) {
    if (!h) return B200MPC_E_ARG;
    if (B < 0 || n_beams < 1 || !(size > 0) || !(resolution > 0)) return set_err(h, B200MPC_E_ARG, "B < 0, n_beams < 1, size or resolution <= 0");
    if (B == 0) return 0;
    if (!scan || !beam_cos || !beam_sin || !yaw || !out) return set_err(h, B200MPC_E_ARG, "NULL argument");
    CU_TRY(h, cudaSetDevice(h->device));
    const int nc = (int)(size * 2.0 / resolution);
    if (nc < 1 || nc > 512) return set_err(h, B200MPC_E_ARG, "grid side int(2*size/resolution) out of range [1,512]");
    const size_t nb = (size_t)B, cells = nb * nc * nc;
    int rc = ensure_buf(h, stage_size(nb * n_beams * 8) + 2 * stage_size((size_t)n_beams * 8) + stage_size(nb * 8) + stage_size(cells));
    if (rc) return rc;
    HostStage st{h, h->d_buf, 0};
    cudaError_t e = cudaSuccess;
    double *d_scan = stage_in(st, scan, nb * n_beams, e), *d_c = stage_in(st, beam_cos, (size_t)n_beams, e);
    double *d_s = stage_in(st, beam_sin, (size_t)n_beams, e), *d_yaw = stage_in(st, yaw, nb, e);
    uint8_t *d_out = stage_in<uint8_t>(st, nullptr, cells, e);
    CU_TRY(h, e);
    rc = b200mpc_local_costmap_batch_device(h, B, n_beams, d_scan, d_c, d_s, d_yaw, size, resolution, kh, kw, d_out, h->stream);
    if (rc) return rc;
    CU_TRY(h, cudaMemcpyAsync(out, d_out, cells, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int b200mpc_raycast_batch(b200mpc_handle *h, int B, int n_beams, const uint32_t *occ_bits, int H, int W, double origin_x,
                                     double origin_y, double resolution, const double *pose, double angle_min, double angle_max,
                                     double range_min, double range_max, double step, double *scan) {
    if (!h) return B200MPC_E_ARG;
    if (B < 0 || n_beams < 1 || H < 1 || W < 1) return set_err(h, B200MPC_E_ARG, "B < 0, n_beams < 1 or empty map");
    if (B == 0) return 0;
    if (!occ_bits || !pose || !scan) return set_err(h, B200MPC_E_ARG, "NULL argument");
    CU_TRY(h, cudaSetDevice(h->device));
    const size_t nb = (size_t)B, nw = (size_t)H * ((W + 31) / 32);
    int rc = ensure_buf(h, stage_size(nw * 4) + stage_size(nb * 24) + stage_size(nb * n_beams * 8));
    if (rc) return rc;
    HostStage st{h, h->d_buf, 0};
    cudaError_t e = cudaSuccess;
    uint32_t *d_occ = stage_in(st, occ_bits, nw, e);
    double *d_pose = stage_in(st, pose, nb * 3, e);
    double *d_scan = stage_in<double>(st, nullptr, nb * n_beams, e);
    CU_TRY(h, e);
    rc = b200mpc_raycast_batch_device(h, B, n_beams, d_occ, H, W, origin_x, origin_y, resolution, d_pose, 3, angle_min, angle_max,
                                      range_min, range_max, step, d_scan, h->stream);
    if (rc) return rc;
    CU_TRY(h, cudaMemcpyAsync(scan, d_scan, nb * n_beams * 8, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int b200mpc_headings_batch(b200mpc_handle *h, int P, int K, const double *path_xy, double dt, double *heading,
                                      double *velocity, double *omega) {
    if (!h) return B200MPC_E_ARG;
    if (P < 0 || K < 2) return set_err(h, B200MPC_E_ARG, "P < 0 or K < 2");
    if (P == 0) return 0;
    if (!path_xy || !heading || !velocity || !omega) return set_err(h, B200MPC_E_ARG, "NULL argument");
    CU_TRY(h, cudaSetDevice(h->device));
    const size_t np_ = (size_t)P * K, nw = (size_t)P * (K - 1);
    int rc = ensure_buf(h, stage_size(np_ * 16) + 2 * stage_size(np_ * 8) + stage_size(nw * 8));
    if (rc) return rc;
    HostStage st{h, h->d_buf, 0};
    cudaError_t e = cudaSuccess;
    double *d_xy = stage_in(st, path_xy, np_ * 2, e);
    double *d_h = stage_in<double>(st, nullptr, np_, e), *d_v = stage_in<double>(st, nullptr, np_, e);
    double *d_w = stage_in<double>(st, nullptr, nw, e);
    CU_TRY(h, e);
    rc = b200mpc_headings_batch_device(h, P, K, d_xy, dt, d_h, d_v, d_w, h->stream);
    if (rc) return rc;
    CU_TRY(h, cudaMemcpyAsync(heading, d_h, np_ * 8, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaMemcpyAsync(velocity, d_v, np_ * 8, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaMemcpyAsync(omega, d_w, nw * 8, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// ---- multi-GPU host-buffer solve ------------------------------------------------------------------------
// Independent problems shard by batch index: device g gets the contiguous slice [lo_g, hi_g) (sizes differ by at most one),
// one host thread and one handle (its own streams) per device, no collective; every device reads its inputs from and
// writes its results straight into its slice of the caller's arrays — there is no gather copy.
extern "C" int b200mpc_solve_batch_multi(b200mpc_handle **handles, int G, int B, const double *x0, const double *xref,
                                         const double *uref, const double *obs_x, const double *obs_y, int obs_stride,
                                         const double *u_init, double *X_out, double *U_out, double *cost_out,
                                         int32_t *status_out, int32_t *iters_out, int32_t *ls_out) {
    if (!handles || G < 1) return B200MPC_E_ARG;
    for (int g = 0; g < G; g++)
        if (!handles[g]) return B200MPC_E_ARG;
    b200mpc_handle *h0 = handles[0];
    if (B < 0) return set_err(h0, B200MPC_E_ARG, "B < 0");
    for (int g = 1; g < G; g++) {
        if (memcmp(&handles[g]->prm, &h0->prm, sizeof(b200mpc_params)) != 0)
            return set_err(h0, B200MPC_E_ARG, "the handles of a multi-device solve must share one parameter set");
        for (int k = 0; k < g; k++)
            if (handles[k] == handles[g]) return set_err(h0, B200MPC_E_ARG, "a handle appears twice");
    }
    if (B == 0) return 0;
    const int N = h0->prm.N;
    const size_t w_ref = (h0->prm.ref_kind == B200MPC_REF_TRAJ) ? 3 * (size_t)N : 3;
    std::vector<int> rcs(G, 0);
    std::vector<std::thread> th;
    th.reserve(G);
    const int base = B / G, rem = B % G;
    for (int g = 0; g < G; g++) {
        const size_t lo = (size_t)g * base + (size_t)(g < rem ? g : rem);
        const int n = base + (g < rem ? 1 : 0);
        if (n == 0) continue;
        th.emplace_back([=, &rcs]() {
            rcs[g] = b200mpc_solve_batch(handles[g], n, x0 + lo * 3, xref + lo * w_ref, uref ? uref + lo * 2 * N : nullptr,
                                         (obs_x && obs_stride) ? obs_x + lo * obs_stride : obs_x,
                                         (obs_y && obs_stride) ? obs_y + lo * obs_stride : obs_y, obs_stride,
                                         u_init ? u_init + lo * 2 * N : nullptr, X_out ? X_out + lo * 3 * (N + 1) : nullptr,
                                         U_out ? U_out + lo * 2 * N : nullptr, cost_out ? cost_out + lo : nullptr,
                                         status_out ? status_out + lo : nullptr, iters_out ? iters_out + lo : nullptr,
                                         ls_out ? ls_out + lo : nullptr);
        });
    }
    for (auto &t : th) t.join();
    for (int g = 0; g < G; g++)
        if (rcs[g]) return set_err(h0, rcs[g], std::string("device ") + std::to_string(handles[g]->device) + ": " + handles[g]->err);
    return 0;
}
