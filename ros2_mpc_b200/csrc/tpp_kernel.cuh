// tpp_kernel.cuh — large-batch solve kernel: one LANE per problem, horizon streamed through a per-warp workspace.
//
// The warp-per-problem kernel (mpc_solve_kernel in b200mpc.cu) keeps one problem in the registers of a warp; its
// Riccati recursion is serial over the horizon, so 31 of 32 lanes idle during more than half of the issued
// instructions (profiles/r1_v1_*).  For large batches this kernel turns the mapping around: every lane runs the
// complete interior-point method of its own problem, so each warp instruction advances 32 problems and every
// phase — linearisation, Riccati sweeps, trial points — runs at full lane utilisation.
//
// Data layout.  A problem's iterate, Riccati factors and steps (TPP_NF doubles per stage) do not fit on chip for
// enough problems, so they live in an HBM workspace private to the warp:
//     ws[warp][stage k][field f][lane]            (doubles)
// A field row is 32 consecutive doubles = 256 B = two full 128-byte lines, every access of the kernel is such a
// row (fully coalesced, compile-time field offsets from one per-stage pointer, L1 bypassed), and a stage record is
// one contiguous block whose next-needed rows are prefetched into L2 one stage ahead.  Each interior-point
// iteration is three streaming sweeps over the stages:
//     B  backward  k = N..0   read iterate, re-linearise, condense, Riccati recursion        -> K, kf
//     F  forward   k = 0..N   read K, kf (+ iterate), roll the step out                      -> dX, dU,
//                             fraction-to-the-boundary step sizes, directional derivative
//     T  trial     k = N..0   read iterate + step, recover the multiplier step from the stationarity rows of the
//                             Newton system (costate recursion, below), write curr + alpha*step into the OTHER
//                             iterate buffer, and evaluate that point: filter quantities (theta, phi) and all KKT
//                             residual norms
// The cost-to-go matrices P_k, p_k of the Riccati recursion never leave the registers of sweep B.  The textbook
// forward pass needs them for the multiplier step, dlam_k = -(p_k + P_k dx_k); here sweep T, which walks the
// horizon backwards anyway, gets the same quantity from the state rows of the very system being solved,
//     lam_k + dlam_k = A_k' (lam_{k+1} + dlam_{k+1}) - g_k - (Hxx_k + delta_w) dx_k - Hxu_k du_k,
// at the price of one more sin/cos pair per stage — 12 of 56 workspace rows per stage and iteration less traffic.
// The trial point is written speculatively; if the line search accepts it (99 % of first trials) it simply is the
// next iterate and the lane already holds its convergence norms, so there is no separate "accept" or "evaluate"
// sweep.  The two iterate buffers swap roles every trip for the whole warp (so that every row access stays one
// full row); the rare lane whose trial point was rejected copies its iterate across at the end of the trip.
//
// Control flow.  Lanes of a warp solve different problems whose line searches, inertia corrections and iteration
// counts differ.  Every lane carries a phase; one trip of the main loop executes the blocks L(oad), B, F, T in
// that order, each for the lanes that are in that phase.  A regular iteration is one trip.  A lane that needs
// something extra (another trial step size, a second-order correction, a larger delta_w) sits out the blocks it
// does not need in the next trip; it never makes the other 31 lanes wait.  A lane whose problem finishes pulls the
// next problem from a global counter in the next trip.  The per-lane solver state (TppLane) lives in shared
// memory so that the sweeps have the register file to themselves.
//
// The algorithm, constants and every decision are those of the warp-per-problem kernel and of oracle/mpc_oracle.c
// (the CPU checker).  Arithmetic differs at rounding level only: sums over stages are sequential, x/s is
// evaluated as x*(1/s), the barrier term uses one logarithm per stage (log of the product of the four slack
// distances), sin/cos of the RK4 mid- and end-point angles come from the angle-addition formulas, and the
// least-squares multiplier estimate is computed for the unscaled objective and scaled afterwards (it is linear).
#pragma once
// The lane kernels are instantiated per problem family (template parameter SPEC), so that the integrator and the kind of
// reference are compile-time constants and the other family's code leaves the sweeps' loop bodies (instruction cache):
//   SPEC 0  generic (run-time switches)   SPEC 1  RK4 + fixed goal (variants A, B)   SPEC 2  Euler + trajectory (variant C)
//   SPEC 3  RK4 + fixed goal + obstacle cost (variant A).  The generic instance handles the obstacle cost through a run-time
//           switch; instances 1 and 2 carry no obstacle code at all.
#define TPP_SPEC_GENERIC 0
#define TPP_SPEC_RK4_GOAL 1
#define TPP_SPEC_EULER_TRAJ 2
#define TPP_SPEC_RK4_GOAL_OBS 3
#define TPP_IS_EULER(P) ((SPEC == TPP_SPEC_RK4_GOAL || SPEC == TPP_SPEC_RK4_GOAL_OBS) ? false : (SPEC == TPP_SPEC_EULER_TRAJ ? true : ((P).integrator == B200MPC_EULER)))
#define TPP_IS_GOAL(P) ((SPEC == TPP_SPEC_RK4_GOAL || SPEC == TPP_SPEC_RK4_GOAL_OBS) ? true : (SPEC == TPP_SPEC_EULER_TRAJ ? false : ((P).ref_kind == B200MPC_REF_GOAL)))
#define TPP_HAS_OBS(P) (SPEC == TPP_SPEC_RK4_GOAL_OBS ? true : (SPEC == TPP_SPEC_GENERIC ? ((P).obs_form != B200MPC_OBS_NONE) : false))
#ifndef TPP_BWD_EAGER
#define TPP_BWD_EAGER 1 /* backward sweep: read all staged rows at the top of the stage and issue the next copy at once */
#endif

// Workspace rows.  A row holds one PAIR of fields for the 32 lanes of a warp: 32 x double2 = 512 B = four full
// 128-byte lines; every workspace access of the kernel is one 128-bit load/store per lane of such a row.
enum {
    R_X01 = 0, R_X2L0 = 1, R_L12 = 2, R_U = 3, R_S = 4, R_YD = 5, R_VL = 6, R_VU = 7, // iterate: (X0,X1) (X2,lam0) (lam1,lam2) U S yd vL vU
    R_ITER = 8,                                                                      // rows per iterate buffer; two buffers: rows 0-7, 8-15
    R_K = 16,    // (K00,K01) (K02,K10) (K11,K12) (kf0,kf1)
    R_STEP = 20, // Newton step: (dX0,dX1) (dX2,-) (dU0,dU1)
    R_SSTEP = 23, // second-order-correction step, same layout
    R_CS = 26,   // second-order-correction right-hand sides: (cs0,cs1) (cs2,-) (ds0,ds1)
    R_REF = 29,  // per-stage references (trajectory tracking only): (r0,r1) (r2,-) (ub0,ub1)
    R_KB = 32,   // two-sweep kernel (tpp_fused.cuh) only: (kfb0,kfb1), barrier-parameter coefficient of the gain kf
    R_OC = 33,   // obstacle cost only: value, gradient and Hessian of the stage's obstacle sum at the iterate of buffer 0 / 1
                 // (unscaled): (val,gx) (gy,hxx) (hxy,hyy); rows 33-35 belong to iterate buffer 0, rows 36-38 to buffer 1
    TPP_NR_BASE = 33, // rows per stage of the instances without the obstacle cost
    TPP_NR = 39
};
#define TPP_ROW_B 512
// bytes per stage record: the instances that can carry the obstacle cost (generic, RK4 + goal + obstacles) have the cache rows
// (a larger stage stride costs the others 5-7 %: measured 193.7 -> 205 ms per 1 M variant-B problems with 39 rows everywhere)
#define TPP_STAGE_B_OF(SPEC_) ((((SPEC_) == TPP_SPEC_GENERIC || (SPEC_) == TPP_SPEC_RK4_GOAL_OBS) ? (int)TPP_NR : (int)TPP_NR_BASE) * TPP_ROW_B)
#define TPP_STAGE_B TPP_STAGE_B_OF(SPEC)
#define TPP_STAGE_B_MAX (TPP_NR * TPP_ROW_B)
#define TPP_STAGE_SLOTS 11                       /* staging rows per warp: the trial sweep needs 11 */
#define TPP_STAGE_SMEM (TPP_STAGE_SLOTS * TPP_ROW_B) /* bytes of shared-memory staging per warp */

enum { PH_LOAD = 0, PH_B = 1, PH_F = 2, PH_T = 3, PH_BACKTRACK = 4, PH_EXPORT = 5, PH_FIN = 6, PH_DONE = 7 };
enum { BM_NEWTON = 0, BM_SOC = 1, BM_LSQ = 2 };
enum { TM_EVAL = 1, TM_LSQ = 2, TM_STEP = 3, TM_STEP_SOC = 4 };

// workspace rows bypass L1 (each row is used once per sweep; L1 is left to the stack)
__device__ __forceinline__ double2 tpp_ld2(const char *p, int row) {
    return __ldcg(reinterpret_cast<const double2 *>(p + row * TPP_ROW_B));
}
__device__ __forceinline__ void tpp_st2(char *p, int row, double a, double b) {
    __stcg(reinterpret_cast<double2 *>(p + row * TPP_ROW_B), make_double2(a, b));
}
// asynchronous global -> shared copy of this lane's 16 bytes of a row (LDGSTS, L1 bypassed); the data is private to
// the lane, so completion needs only the lane's own cp.async.wait_group — no barrier, no warp synchronisation
__device__ __forceinline__ void tpp_cp16(char *sdst, const char *gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
// L2 prefetch of this lane's 16 bytes of `n` consecutive rows (the warp's 32 lanes cover the rows' four lines each)
#ifndef TPP_L2_PREFETCH
#define TPP_L2_PREFETCH 1
#endif
__device__ __forceinline__ void tpp_l2_prefetch(const char *p, int row0, int n) {
#if TPP_L2_PREFETCH
#pragma unroll
    for (int i = 0; i < n; i++) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + (row0 + i) * TPP_ROW_B));
#endif
}
__device__ __forceinline__ void tpp_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tpp_cp_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ double2 tpp_sld(const char *sb, int slot) {
    return *reinterpret_cast<const double2 *>(sb + slot * TPP_ROW_B);
}
// The staging buffer is single: the next stage's copy is issued as soon as this stage's values are in registers.
// This empty asm makes "in registers" explicit (it consumes the values), so the copy cannot overtake the reads.
__device__ __forceinline__ void tpp_consume(const double2 &a, const double2 &b, const double2 &c, const double2 &d) {
    asm volatile("" ::"d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y), "d"(c.x), "d"(c.y), "d"(d.x), "d"(d.y) : "memory");
}

// Per-lane mode words are loop-invariant, so the compiler would "unswitch" the stage loops into one copy per mode:
// lanes of a warp that are in different modes (a freshly loaded problem next to one in mid-solve) would then run
// their sweeps one after the other, and the code would outgrow the instruction cache.  Passing the mode through an
// empty volatile asm every stage keeps one loop whose mode-specific parts are short, reconverging branches.
// The same is done with the phase word at every block of the main loop, together with a __syncwarp(): otherwise
// the compiler threads the jump "end of block F, phase := T" straight into block T, and the lanes that arrive
// there by different routes execute the sweep as separate groups (measured: 10.6 active lanes per instruction).
__device__ __forceinline__ int tpp_opaque(int v) {
    asm volatile("" : "+r"(v));
    return v;
}
// Reciprocal to within 1 ulp: hardware seed (MUFU.RCP64H) + two Newton steps.  Five instructions instead of the
// ~25 of an IEEE division; the loop bodies have to stay small for the instruction cache.  The operands here
// (slack distances, 2x2 determinants, step components) are normal, finite and non-zero.
__device__ __forceinline__ double tpp_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
}
__device__ __noinline__ double tpp_exp(double a) { return exp(a); }
__device__ __noinline__ double tpp_pow(double a, double b) { return pow(a, b); }
__device__ __noinline__ double tpp_log10(double a) { return log10(a); }

struct TppArgs {
    BatchArgs a;
    double *ws;   // [nwarps][N+1][TPP_NF][32]
    double *filt; // [nwarps][64][32]: 32 filter entries (phi, theta) per lane
    unsigned long long *stats; // [8]: per sweep B, F, T: warp executions and active lanes; trips; warps
    // streamed host-buffer solves (b200mpc_solve_batch): the inputs arrive chunk by chunk while the kernel runs and
    // the results leave chunk by chunk.  All three are NULL for device-buffer solves.
    const unsigned *avail;     // number of leading problems whose inputs are on the device (written by the copy stream)
    unsigned *done;            // [chunks] finished problems per chunk
    unsigned *flags;           // [chunks] host-mapped: set to 1 when the chunk's results are complete in device memory
    int chunk;                 // problems per chunk
    const unsigned *abort;     // host-mapped word: non-zero = the host gave the call up (a copy failed): stop waiting for inputs
    int cta_sync;              // 1: the warps of a CTA run the sweeps in lock-step (instruction cache); 0: every warp on its own
    int obs_smem;              // 1: the dynamic shared memory has room for one obstacle list per warp (behind TPP_SMEM_BYTES)
    int stage_b;               // bytes per stage record of the launched instance (TPP_STAGE_B_OF)
    int b_passes;              // executions of block B per trip (>= 1)
};
#ifndef TPP_STATS
#define TPP_STATS 0
#endif
__device__ __forceinline__ void tpp_stat(const TppArgs &T, int i) {
#if TPP_STATS
    const unsigned m = __activemask();
    if ((threadIdx.x & 31) == __ffs(m) - 1) {
        atomicAdd(T.stats + 2 * i, 1ull);
        atomicAdd(T.stats + 2 * i + 1, (unsigned long long)__popc(m));
    }
#endif
}

// Residual norms of an evaluated point (sweep T -> convergence tests of the same trip).
struct TppNorms {
    double theta, prim_inf, dual_inf, sum_y, sum_z, pmin, pmax, f, slog;
};

// Solver state of one lane (shared memory, one record per thread; the record stride is an odd multiple of 8 bytes
// so that the 64-bit accesses of a half-warp fall into distinct banks).
struct TppLane {
    double theta, f, slog;                         // of the current iterate: constraint violation, objective, sum of log slacks
    double df, mu, theta0, dw, dw_last;            // theta0 = max(1, theta at the starting point)
    double ref_phi, ref_gbd, alpha, a_min, a_z;    // line-search reference values, Newton step sizes
    double alpha_soc, a_z_soc, theta_soc_old;      // second-order correction
    double goal[3];                                // goal state (ref_kind GOAL)
    int b, phase, bmode, tmode;
    int status, iter, ls_extra, n_resto, acceptable_count, ntrial, soc_count, ring;
    unsigned fmask;
    int keep, soc_first, moved;
    int tiny, tiny_last, tiny_flag;                // tiny-step detection (BacktrackingLineSearch::DetectTinyStep and its two flags)
    double ymax_f;                                 // LSQ mode: largest slack-multiplier estimate seen by sweep F
    double dw_b;                                   // two-sweep kernel: delta_w of the factorisation in progress
    int n_eff;                                     // obstacle cost: entries of the problem's obstacle list to walk (ObsList)
    int hslot, hcap;                               // hand-over (BatchArgs::hand_rec): record the problem is exported to; iteration threshold
};
#define TPP_LANE_STRIDE (((sizeof(TppLane) + 7) / 8) | 1) /* in doubles, odd */

// (tpp_sincos_core: b200mpc.cu, shared with the warp kernel)
__device__ __noinline__ void tpp_sincos_far(double x, double *sn, double *cs) { sincos(x, sn, cs); }
__device__ __forceinline__ void tpp_sincos(double x, double &sn, double &cs) {
    if (fabs(x) <= 1e5) tpp_sincos_core(x, sn, cs); // (NaN takes the library routine)
    else tpp_sincos_far(x, &sn, &cs);
}
// sin/cos of th and hw: one out-of-line copy shared by all sweeps (instruction-cache footprint)
struct TppSC { double s0, c0, sh, ch; };
__device__ __noinline__ TppSC tpp_sincos2(double th, double hw) {
    TppSC r;
    tpp_sincos(th, r.s0, r.c0);
    tpp_sincos(hw, r.sh, r.ch);
    return r;
}
__device__ __noinline__ void tpp_sincos1(double th, double *sn, double *cs) { tpp_sincos(th, *sn, *cs); }
// sin/cos of th, th + hw, th + 2 hw (RK4 stage angles) from two sincos evaluations
__device__ __forceinline__ void tpp_trig(double th, double hw, double &s0, double &c0, double &sm, double &cm,
                                         double &se, double &ce) {
    const TppSC t = tpp_sincos2(th, hw);
    s0 = t.s0; c0 = t.c0;
    const double sh = t.sh, ch = t.ch;
    sm = s0 * ch + c0 * sh;
    cm = c0 * ch - s0 * sh;
    const double s2 = 2.0 * sh * ch, c2 = 1.0 - 2.0 * sh * sh;
    se = s0 * c2 + c0 * s2;
    ce = c0 * c2 - s0 * s2;
}

struct TppLin {
    double a13, a23, b11, b12, b21, b22, F0, F1, F2;
    double hxx, hxy, hyy, htt, htv, htw, hvv, hvw, hww;
    double g[5], f;
};

// Obstacle sum of one stage (unscaled): value, gradient, Hessian — read from the stage's cache rows (R_OC), which the
// obstacle block of the kernel (tpp_obstacle_block) fills before the sweeps that need them.
struct TppObs {
    double v, gx, gy, hxx, hxy, hyy;
};
__device__ __forceinline__ TppObs tpp_obs_zero() {
    TppObs o;
    o.v = 0; o.gx = 0; o.gy = 0; o.hxx = 0; o.hxy = 0; o.hyy = 0;
    return o;
}
// cache rows of iterate buffer `buf` at the stage whose record starts at p
__device__ __forceinline__ TppObs tpp_obs_load(const char *p, int buf) {
    const double2 a = tpp_ld2(p, R_OC + 3 * buf), b = tpp_ld2(p, R_OC + 3 * buf + 1), c = tpp_ld2(p, R_OC + 3 * buf + 2);
    TppObs o;
    o.v = a.x; o.gx = a.y; o.gy = b.x; o.hxx = b.y; o.hxy = c.x; o.hyy = c.y;
    return o;
}

// K1+K2 of a stage with controls (k < N): integration step, Jacobian entries, stage cost and its gradient; with
// HESS also the Lagrangian Hessian block (ln = multiplier of the defect X_{k+1} - F(X_k, U_k)).
template <bool HESS, int SPEC>
__device__ __forceinline__ void tpp_lin(const KParams &P, const double r[3], const double ub[2], const double X[3],
                                        const double U[2], const double ln[3], double df, TppLin &o, const TppObs &oc) {
    const double dt = P.dt, th = X[2], v = U[0], w = U[1];
    o.b12 = 0; o.b22 = 0;
    o.htw = 0; o.hvw = 0; o.hww = 0; o.hxy = 0;
    if (TPP_IS_EULER(P)) {
        double sn, cs;
        tpp_sincos1(th, &sn, &cs);
        o.F0 = X[0] + dt * v * cs;
        o.F1 = X[1] + dt * v * sn;
        o.a13 = -dt * v * sn; o.a23 = dt * v * cs;
        o.b11 = dt * cs; o.b21 = dt * sn;
        if (HESS) {
            o.htt = -(ln[0] * (-dt * v * cs) + ln[1] * (-dt * v * sn));
            o.htv = -(ln[0] * (-dt * sn) + ln[1] * (dt * cs));
        }
    } else {
        double s0, c0, sm, cm, se, ce;
        tpp_trig(th, 0.5 * dt * w, s0, c0, sm, cm, se, ce);
        const double h = dt / 6.0;
        const double C = c0 + 4.0 * cm + ce, S = s0 + 4.0 * sm + se;
        const double C1 = 2.0 * cm + ce, S1 = 2.0 * sm + se;
        o.F0 = X[0] + h * v * C;
        o.F1 = X[1] + h * v * S;
        o.a13 = -h * v * S; o.a23 = h * v * C;
        o.b11 = h * C; o.b21 = h * S;
        o.b12 = -h * dt * v * S1; o.b22 = h * dt * v * C1;
        if (HESS) {
            const double C2 = cm + ce, S2 = sm + se;
            o.htt = -(ln[0] * (-h * v * C) + ln[1] * (-h * v * S));
            o.htv = -(ln[0] * (-h * S) + ln[1] * (h * C));
            o.htw = -(ln[0] * (-h * dt * v * C1) + ln[1] * (-h * dt * v * S1));
            o.hvw = -(ln[0] * (-h * dt * S1) + ln[1] * (h * dt * C1));
            o.hww = -(ln[0] * (-h * dt * dt * v * C2) + ln[1] * (-h * dt * dt * v * S2));
        }
    }
    o.F2 = th + dt * w;
    const double e0 = X[0] - r[0], e1 = X[1] - r[1], e2 = X[2] - r[2];
    const double m0 = v - ub[0], m1 = w - ub[1];
    const double er = tpp_exp(-P.kappa * v);
    o.f = e0 * P.Q[0] * e0 + e1 * P.Q[1] * e1 + e2 * P.Q[2] * e2 + m0 * P.R[0] * m0 + m1 * P.R[1] * m1 + er;
    o.g[0] = df * 2.0 * P.Q[0] * e0;
    o.g[1] = df * 2.0 * P.Q[1] * e1;
    o.g[2] = df * 2.0 * P.Q[2] * e2;
    o.g[3] = df * (2.0 * P.R[0] * m0 - P.kappa * er);
    o.g[4] = df * 2.0 * P.R[1] * m1;
    if (HESS) {
        o.hxx = df * 2.0 * P.Q[0];
        o.hyy = df * 2.0 * P.Q[1];
        o.htt = df * 2.0 * P.Q[2] + o.htt;
        o.hvv = df * (2.0 * P.R[0] + P.kappa * P.kappa * er);
        o.hww = df * 2.0 * P.R[1] + o.hww;
    }
    if (TPP_HAS_OBS(P)) {
        o.f += oc.v;
        o.g[0] = fma(df, oc.gx, o.g[0]);
        o.g[1] = fma(df, oc.gy, o.g[1]);
        if (HESS) {
            o.hxx = fma(df, oc.hxx, o.hxx);
            o.hyy = fma(df, oc.hyy, o.hyy);
            o.hxy = df * oc.hxy;
        }
    }
}

// value of the integration step only (second-order-correction defects, restoration roll-out: rare paths)
template <int SPEC>
__device__ __noinline__ void tpp_dyn(const KParams &P, const double X[3], const double U[2], double F[3]) {
    const double dt = P.dt, th = X[2], v = U[0], w = U[1];
    if (TPP_IS_EULER(P)) {
        double sn, cs;
        tpp_sincos1(th, &sn, &cs);
        F[0] = X[0] + dt * v * cs;
        F[1] = X[1] + dt * v * sn;
    } else {
        double s0, c0, sm, cm, se, ce;
        tpp_trig(th, 0.5 * dt * w, s0, c0, sm, cm, se, ce);
        const double h = dt / 6.0;
        F[0] = X[0] + h * v * (c0 + 4.0 * cm + ce);
        F[1] = X[1] + h * v * (s0 + 4.0 * sm + se);
    }
    F[2] = th + dt * w;
}

// per-stage references: the goal (registers) or the tracking reference rows of the stage
template <int SPEC>
__device__ __forceinline__ void tpp_ref(const KParams &P, const double goal[3], const char *p, double r[3], double ub[2]) {
    if (TPP_IS_GOAL(P)) {
        r[0] = goal[0]; r[1] = goal[1]; r[2] = goal[2];
        ub[0] = 0; ub[1] = 0;
    } else {
        const double2 a = tpp_ld2(p, R_REF), b = tpp_ld2(p, R_REF + 1), c = tpp_ld2(p, R_REF + 2);
        r[0] = a.x; r[1] = a.y; r[2] = b.x;
        ub[0] = c.x; ub[1] = c.y;
    }
}


struct TppBwd {
    double gmax, f;
    int ok, bad;
};

// ---- sweep B: condensed stage-wise KKT system, Riccati backward recursion -----------------------------------------
// bmode NEWTON: right-hand side = KKT residual of the current iterate.
// bmode SOC:    same matrix; the defect right-hand sides are advanced in the same sweep from the last trial point
//               (csoc = a*csoc + c(trial), dsoc = a*dsoc + (U_t - S_t)) and stored for the forward sweep.
// bmode LSQ:    least-squares multiplier estimate (W = 0, delta = 1, Sigma = 1, rhs = objective gradient), on the
//               unscaled objective (L.df is 1); also returns the largest gradient entry (objective scaling), the
//               objective and the invalid-number flag of the starting point.
// o.ok = 0 when a condensed Quu block is not positive definite (wrong inertia).
template <int SPEC>
__device__ __forceinline__ void tpp_backward(const KParams &P, const BatchArgs &A, char *wb, char *sb, int cur,
                                             const TppLane &L, TppBwd &o) {
    const int N = P.N;
    const double dt = P.dt;
    const int bmode0 = L.bmode;
    const double dw = (bmode0 != BM_LSQ) ? L.dw : 1.0;
    const double df = L.df, mu = L.mu;
    const double *goal = L.goal;
    const int sfirst = L.soc_first;
    const int srow = sfirst ? R_STEP : R_SSTEP;
    const double at = sfirst ? L.alpha : L.alpha_soc;
    const int co = cur * R_ITER;
    double Xn[3] = {0, 0, 0}, ln[3] = {0, 0, 0};
    double q00 = 0, q01 = 0, q02 = 0, q11 = 0, q12 = 0, q22 = 0, v0 = 0, v1 = 0, v2 = 0;
    double gmax = 0, fs = 0;
    int ok = 1, bad = 0;
    {
        const char *pc = wb + (size_t)N * TPP_STAGE_B + co * TPP_ROW_B;
#pragma unroll
        for (int i = 0; i < R_ITER; i++) tpp_cp16(sb + i * TPP_ROW_B, pc + i * TPP_ROW_B);
        tpp_cp_commit();
    }
#pragma unroll 1
    for (int k = N; k >= 0; --k) {
        const int mode = tpp_opaque(bmode0);
        const bool useW = (mode != BM_LSQ);
        char *p = wb + (size_t)k * TPP_STAGE_B;
        tpp_cp_wait();
        // the stage's rows are read from the staging buffer where they are needed; the copy of the next stage is
        // issued right after the last of them (TPP_BWD_PREFETCH), in front of the Riccati algebra
        const double2 x01 = tpp_sld(sb, R_X01), x2l0 = tpp_sld(sb, R_X2L0), l12 = tpp_sld(sb, R_L12), u2 = tpp_sld(sb, R_U);
#if TPP_BWD_EAGER
        const double2 s2 = tpp_sld(sb, R_S), yd2 = tpp_sld(sb, R_YD), vl2 = tpp_sld(sb, R_VL), vu2 = tpp_sld(sb, R_VU);
        tpp_consume(x01, x2l0, l12, u2);
        tpp_consume(s2, yd2, vl2, vu2);
#endif
        if (k > 1) tpp_l2_prefetch(p - 2 * TPP_STAGE_B + co * TPP_ROW_B, 0, R_ITER);
#define TPP_BWD_PREFETCH()                                                                                   \
    do {                                                                                                     \
        if (k > 0) {                                                                                         \
            const char *pcn = p - TPP_STAGE_B + co * TPP_ROW_B;                                              \
            _Pragma("unroll") for (int i = 0; i < R_ITER; i++) tpp_cp16(sb + i * TPP_ROW_B, pcn + i * TPP_ROW_B); \
            tpp_cp_commit();                                                                                 \
        }                                                                                                    \
    } while (0)
        const double X[3] = {x01.x, x01.y, x2l0.x};
        double lam[3] = {0, 0, 0};
        if (k >= 1) { lam[0] = x2l0.y; lam[1] = l12.x; lam[2] = l12.y; }
        // obstacle sum of the stage at the current iterate (cache rows; zeros outside the stages that carry the cost)
        TppObs oc = tpp_obs_zero();
        if (TPP_HAS_OBS(P)) {
            oc = tpp_obs_load(p, cur);
            if (k > 1) tpp_l2_prefetch(p - 2 * TPP_STAGE_B, R_OC + 3 * cur, 3);
        }
#if TPP_BWD_EAGER
        TPP_BWD_PREFETCH();
#endif
        if (k == N) {
            // terminal stage: no tracking cost, no controls (the obstacle cost of variant A covers it)
#if !TPP_BWD_EAGER
            tpp_consume(x01, x2l0, l12, u2);
            TPP_BWD_PREFETCH();
#endif
            q00 = dw; q01 = 0; q02 = 0; q11 = dw; q12 = 0; q22 = dw;
            if (mode == BM_LSQ) { v0 = 0; v1 = 0; v2 = 0; }
            else { v0 = lam[0]; v1 = lam[1]; v2 = lam[2]; }
            if (TPP_HAS_OBS(P)) {
                const double g0 = df * oc.gx, g1 = df * oc.gy;
                v0 += g0; v1 += g1;
                if (useW) { q00 = fma(df, oc.hxx, q00); q01 = df * oc.hxy; q11 = fma(df, oc.hyy, q11); }
                if (mode == BM_LSQ) {
                    fs += oc.v;
                    if (!isfinite(g0) || !isfinite(g1)) bad = 1;
                    gmax = fmax(gmax, fmax(fabs(g0), fabs(g1)));
                }
            }
        } else {
            const double U[2] = {u2.x, u2.y};
            double r[3], ub[2];
            tpp_ref<SPEC>(P, goal, p, r, ub);
            TppLin q;
            tpp_lin<true, SPEC>(P, r, ub, X, U, ln, df, q, oc);
#if !TPP_BWD_EAGER
            const double2 s2 = tpp_sld(sb, R_S), yd2 = tpp_sld(sb, R_YD), vl2 = tpp_sld(sb, R_VL), vu2 = tpp_sld(sb, R_VU);
            tpp_consume(x01, x2l0, l12, u2);
            tpp_consume(s2, yd2, vl2, vu2);
            TPP_BWD_PREFETCH();
#endif
            const double c0 = Xn[0] - q.F0, c1 = Xn[1] - q.F1, c2 = Xn[2] - q.F2;
            double rx0, rx1, rx2, ru[2], Dsig[2], rs[2], rd[2], rc[3];
            if (mode == BM_LSQ) {
                rx0 = q.g[0]; rx1 = q.g[1]; rx2 = q.g[2];
                ru[0] = q.g[3]; ru[1] = q.g[4];
                Dsig[0] = Dsig[1] = 1.0; rs[0] = rs[1] = 0.0; rd[0] = rd[1] = 0.0;
                rc[0] = rc[1] = rc[2] = 0.0;
                fs += q.f;
                if (!isfinite(c0) || !isfinite(c1) || !isfinite(c2) || !isfinite(q.g[3]) || !isfinite(q.g[4])) bad = 1;
                gmax = fmax(gmax, fmax(fabs(q.g[3]), fabs(q.g[4])));
                if (k >= 1) {
                    if (!isfinite(q.g[0]) || !isfinite(q.g[1]) || !isfinite(q.g[2])) bad = 1;
                    gmax = fmax(gmax, fmax(fabs(q.g[0]), fmax(fabs(q.g[1]), fabs(q.g[2]))));
                }
            } else {
                rx0 = q.g[0] + lam[0] - ln[0];
                rx1 = q.g[1] + lam[1] - ln[1];
                rx2 = q.g[2] + lam[2] - (q.a13 * ln[0] + q.a23 * ln[1] + ln[2]);
                const double S[2] = {s2.x, s2.y}, yd[2] = {yd2.x, yd2.y}, vL[2] = {vl2.x, vl2.y}, vU[2] = {vu2.x, vu2.y};
                ru[0] = q.g[3] - (q.b11 * ln[0] + q.b21 * ln[1]) + yd[0];
                ru[1] = q.g[4] - (q.b12 * ln[0] + q.b22 * ln[1] + dt * ln[2]) + yd[1];
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const double isl = tpp_rcp(S[i] - P.sL[i]), isu = tpp_rcp(P.sU[i] - S[i]);
                    rs[i] = -yd[i] - mu * isl + mu * isu;
                    Dsig[i] = vL[i] * isl + vU[i] * isu + dw;
                    rd[i] = U[i] - S[i];
                }
                rc[0] = c0; rc[1] = c1; rc[2] = c2;
                if (mode == BM_SOC) {
                    // defects of the last trial point curr + at*step
                    const double2 d01 = tpp_ld2(p, srow), d2 = tpp_ld2(p, srow + 1), du2 = tpp_ld2(p, srow + 2);
                    double2 dsp = make_double2(rd[0], rd[1]), cs01 = make_double2(rc[0], rc[1]), cs2 = make_double2(rc[2], 0.0);
                    if (!sfirst) { dsp = tpp_ld2(p, R_CS + 2); cs01 = tpp_ld2(p, R_CS); cs2 = tpp_ld2(p, R_CS + 1); }
                    const double du[2] = {du2.x, du2.y}, rdp[2] = {dsp.x, dsp.y}, base[3] = {cs01.x, cs01.y, cs2.x};
                    // trial state of the successor stage (its step rows are still in the workspace)
                    const double2 n01 = tpp_ld2(p + TPP_STAGE_B, srow), n2 = tpp_ld2(p + TPP_STAGE_B, srow + 1);
                    const double Xtn[3] = {Xn[0] + at * n01.x, Xn[1] + at * n01.y, Xn[2] + at * n2.x};
                    double Xt[3], Ut[2], St[2], Ft[3];
                    Xt[0] = X[0] + at * d01.x; Xt[1] = X[1] + at * d01.y; Xt[2] = X[2] + at * d2.x;
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        Ut[i] = U[i] + at * du[i];
                        St[i] = S[i] + at * (du[i] + rdp[i]);
                        rd[i] = at * rdp[i] + (Ut[i] - St[i]);
                    }
                    tpp_dyn<SPEC>(P, Xt, Ut, Ft);
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        rc[i] = at * base[i] + (Xtn[i] - Ft[i]);
                    }
                    tpp_st2(p, R_CS, rc[0], rc[1]); tpp_st2(p, R_CS + 1, rc[2], 0.0); tpp_st2(p, R_CS + 2, rd[0], rd[1]);
                }
            }
            const double hxx = useW ? q.hxx : 0.0, hyy = useW ? q.hyy : 0.0, hxy = useW ? q.hxy : 0.0;
            const double htt = useW ? q.htt : 0.0, htv = useW ? q.htv : 0.0, htw = useW ? q.htw : 0.0;
            const double hvv = useW ? q.hvv : 0.0, hvw = useW ? q.hvw : 0.0, hww = useW ? q.hww : 0.0;
            const double a = q.a13, b = q.a23, b11 = q.b11, b12 = q.b12, b21 = q.b21, b22 = q.b22;
            const double d0 = -rc[0], d1 = -rc[1], d2 = -rc[2];
            // w = P d + p
            const double w0 = q00 * d0 + q01 * d1 + q02 * d2 + v0;
            const double w1 = q01 * d0 + q11 * d1 + q12 * d2 + v1;
            const double w2 = q02 * d0 + q12 * d1 + q22 * d2 + v2;
            // t = P[:,2] + a P[:,0] + b P[:,1]
            const double t0 = q02 + a * q00 + b * q01;
            const double t1 = q12 + a * q01 + b * q11;
            const double t2 = q22 + a * q02 + b * q12;
            // Qxx = Hxx + A'PA
            const double x00 = hxx + dw + q00, x01_ = hxy + q01, x11 = hyy + dw + q11;
            const double x02 = t0, x12 = t1, x22 = htt + dw + t2 + a * t0 + b * t1;
            // PB columns
            const double e0 = b11 * q00 + b21 * q01, e1 = b11 * q01 + b21 * q11;
            const double f0 = b12 * q00 + b22 * q01 + dt * q02, f1 = b12 * q01 + b22 * q11 + dt * q12,
                         f2 = b12 * q02 + b22 * q12 + dt * q22;
            // Qux = Hux + B'PA
            const double u00 = e0, u01 = e1, u02 = htv + b11 * t0 + b21 * t1;
            const double u10 = f0, u11 = f1, u12 = htw + b12 * t0 + b22 * t1 + dt * t2;
            // Quu = Huu + Dsig + B'PB
            const double r00 = hvv + dw + Dsig[0] + b11 * e0 + b21 * e1;
            const double r01 = hvw + b11 * f0 + b21 * f1;
            const double r11 = hww + dw + Dsig[1] + b12 * f0 + b22 * f1 + dt * f2;
            const bool first = (k == 0);
            const double gx0 = (first ? 0.0 : rx0) + w0;
            const double gx1 = (first ? 0.0 : rx1) + w1;
            const double gx2 = (first ? 0.0 : rx2) + a * w0 + b * w1 + w2;
            const double qu0 = ru[0] + Dsig[0] * rd[0] + rs[0];
            const double qu1 = ru[1] + Dsig[1] * rd[1] + rs[1];
            const double gu0 = qu0 + b11 * w0 + b21 * w1;
            const double gu1 = qu1 + b12 * w0 + b22 * w1 + dt * w2;
            const double det = r00 * r11 - r01 * r01;
            if (!(r00 > 0.0) || !(det > 0.0)) ok = 0;
            const double idet = tpp_rcp(det);
            const double i00 = r11 * idet, i01 = -r01 * idet, i11 = r00 * idet;
            const double K00 = -(i00 * u00 + i01 * u10), K01 = -(i00 * u01 + i01 * u11), K02 = -(i00 * u02 + i01 * u12);
            const double K10 = -(i01 * u00 + i11 * u10), K11 = -(i01 * u01 + i11 * u11), K12 = -(i01 * u02 + i11 * u12);
            const double k0 = -(i00 * gu0 + i01 * gu1), k1 = -(i01 * gu0 + i11 * gu1);
            tpp_st2(p, R_K, K00, K01); tpp_st2(p, R_K + 1, K02, K10); tpp_st2(p, R_K + 2, K11, K12);
            tpp_st2(p, R_K + 3, k0, k1);
            // P' = Qxx + Qux'K (symmetrised), p' = gx + Qux' kf
            q00 = x00 + u00 * K00 + u10 * K10;
            q11 = x11 + u01 * K01 + u11 * K11;
            q22 = x22 + u02 * K02 + u12 * K12;
            q01 = x01_ + 0.5 * ((u00 * K01 + u10 * K11) + (u01 * K00 + u11 * K10));
            q02 = x02 + 0.5 * ((u00 * K02 + u10 * K12) + (u02 * K00 + u12 * K10));
            q12 = x12 + 0.5 * ((u01 * K02 + u11 * K12) + (u02 * K01 + u12 * K11));
            v0 = gx0 + u00 * k0 + u10 * k1;
            v1 = gx1 + u01 * k0 + u11 * k1;
            v2 = gx2 + u02 * k0 + u12 * k1;
        }
        Xn[0] = X[0]; Xn[1] = X[1]; Xn[2] = X[2];
        ln[0] = lam[0]; ln[1] = lam[1]; ln[2] = lam[2];
    }
#undef TPP_BWD_PREFETCH
    o.ok = ok; o.bad = bad; o.gmax = gmax; o.f = fs;
}

struct TppFwd {
    double a_max, a_z, gbd, ymax;
    int bad, tiny;
};

// staging of the forward sweep: slots 0-3 = K, kf rows of stage k, 4-7 = U, S, vL, vU of stage k,
// 8-9 = (X0,X1) (X2,lam0) of stage k+1
template <int SPEC>
__device__ __forceinline__ void tpp_forward_stage(char *sb, const char *p, int co, bool has_next) {
#pragma unroll
    for (int i = 0; i < 4; i++) tpp_cp16(sb + i * TPP_ROW_B, p + (R_K + i) * TPP_ROW_B);
    const char *pc = p + co * TPP_ROW_B;
    tpp_cp16(sb + 4 * TPP_ROW_B, pc + R_U * TPP_ROW_B);
    tpp_cp16(sb + 5 * TPP_ROW_B, pc + R_S * TPP_ROW_B);
    tpp_cp16(sb + 6 * TPP_ROW_B, pc + R_VL * TPP_ROW_B);
    tpp_cp16(sb + 7 * TPP_ROW_B, pc + R_VU * TPP_ROW_B);
    if (has_next) {
        tpp_cp16(sb + 8 * TPP_ROW_B, pc + TPP_STAGE_B + R_X01 * TPP_ROW_B);
        tpp_cp16(sb + 9 * TPP_ROW_B, pc + TPP_STAGE_B + R_X2L0 * TPP_ROW_B);
    }
    tpp_cp_commit();
}

// ---- sweep F: forward roll-out of the step; step sizes and directional derivative ----------------------------------
// (the multiplier step is recovered by sweep T; in LSQ mode ymax covers the slack-multiplier estimate only)
template <int SPEC>
__device__ __forceinline__ void tpp_forward(const KParams &P, const BatchArgs &A, char *wb, char *sb, int cur,
                                            const TppLane &L, TppFwd &o) {
    const int N = P.N;
    const double dt = P.dt, mu = L.mu, tau = fmax(TAU_MIN, 1.0 - L.mu), df = L.df;
    const double *goal = L.goal;
    const int bmode0 = L.bmode;
    const int co = cur * R_ITER;
    const int orow = (bmode0 == BM_SOC) ? R_SSTEP : R_STEP;
    double y0 = 0, y1 = 0, y2 = 0;
    double a_max = 1.0, a_z = 1.0, gbd = 0, ymax = 0;
    int bad = 0;
    int big = 0;      // a primal component of the Newton step exceeds tiny_step_tol (relative)
    double c2 = 0;    // squared 2-norm of the primal infeasibility of the current iterate
    double X[3];
    {
        const double2 a = tpp_ld2(wb + co * TPP_ROW_B, R_X01), b = tpp_ld2(wb + co * TPP_ROW_B, R_X2L0);
        X[0] = a.x; X[1] = a.y; X[2] = b.x;
    }
    tpp_forward_stage<SPEC>(sb, wb, co, N > 0);
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
        const int mode = tpp_opaque(bmode0);
        char *p = wb + (size_t)k * TPP_STAGE_B;
        tpp_cp_wait();
        const double2 k0_ = tpp_sld(sb, 0), k1_ = tpp_sld(sb, 1), k2_ = tpp_sld(sb, 2), kf_ = tpp_sld(sb, 3);
        const double2 u2 = tpp_sld(sb, 4), s2 = tpp_sld(sb, 5), vl2 = tpp_sld(sb, 6), vu2 = tpp_sld(sb, 7);
        const double2 xn01 = tpp_sld(sb, 8), xn2 = tpp_sld(sb, 9);
        tpp_consume(k0_, k1_, k2_, kf_);
        tpp_consume(u2, s2, vl2, vu2);
        tpp_consume(xn01, xn2, xn2, xn2);
        if (k < N) tpp_forward_stage<SPEC>(sb, p + TPP_STAGE_B, co, k + 1 < N);
        if (k + 2 <= N) {
            tpp_l2_prefetch(p + 2 * TPP_STAGE_B, R_K, 4);
            tpp_l2_prefetch(p + 2 * TPP_STAGE_B + co * TPP_ROW_B, R_U, 2);
            tpp_l2_prefetch(p + 2 * TPP_STAGE_B + co * TPP_ROW_B, R_VL, 2);
            if (k + 3 <= N) tpp_l2_prefetch(p + 3 * TPP_STAGE_B + co * TPP_ROW_B, R_X01, 2);
        }
        if (!isfinite(y0) || !isfinite(y1) || !isfinite(y2)) bad = 1;
        tpp_st2(p, orow, y0, y1); tpp_st2(p, orow + 1, y2, 0.0);
        if (k >= 1 && (!(fabs(y0) <= TINY_STEP_TOL * (1.0 + fabs(X[0]))) || !(fabs(y1) <= TINY_STEP_TOL * (1.0 + fabs(X[1]))) ||
                       !(fabs(y2) <= TINY_STEP_TOL * (1.0 + fabs(X[2])))))
            big = 1;
        // obstacle gradient of the stage at the current iterate (directional derivative of the objective)
        TppObs oc = tpp_obs_zero();
        if (TPP_HAS_OBS(P) && mode == BM_NEWTON) {
            const double2 a = tpp_ld2(p, R_OC + 3 * cur), b = tpp_ld2(p, R_OC + 3 * cur + 1);
            oc.gx = a.y; oc.gy = b.x;
            if (k + 2 <= N) tpp_l2_prefetch(p + 2 * TPP_STAGE_B, R_OC + 3 * cur, 2);
            if (k == N) gbd += df * (oc.gx * y0 + oc.gy * y1);
        }
        if (k < N) {
            const double du0 = kf_.x + k0_.x * y0 + k0_.y * y1 + k1_.x * y2;
            const double du1 = kf_.y + k1_.y * y0 + k2_.x * y1 + k2_.y * y2;
            tpp_st2(p, orow + 2, du0, du1);
            if (!isfinite(du0) || !isfinite(du1)) bad = 1;
            const double du[2] = {du0, du1};
            const double U[2] = {u2.x, u2.y}, Sv[2] = {s2.x, s2.y}, vLv[2] = {vl2.x, vl2.y}, vUv[2] = {vu2.x, vu2.y};
            const double Xn[3] = {xn01.x, xn01.y, xn2.x};
            double r[3], ub[2];
            tpp_ref<SPEC>(P, goal, p, r, ub);
            const double ln0[3] = {0, 0, 0};
            TppLin q;
            tpp_lin<false, SPEC>(P, r, ub, X, U, ln0, df, q, oc);
            double rc0, rc1, rc2, rdv[2] = {U[0] - Sv[0], U[1] - Sv[1]};
            if (mode == BM_LSQ) {
                rc0 = rc1 = rc2 = 0;
                ymax = fmax(ymax, fmax(fabs(du0), fabs(du1))); // dyd = Sigma*dS + rs with Sigma = 1, rd = rs = 0
            } else {
                if (mode == BM_SOC) {
                    const double2 c01 = tpp_ld2(p, R_CS), c2 = tpp_ld2(p, R_CS + 1), d2 = tpp_ld2(p, R_CS + 2);
                    rc0 = c01.x; rc1 = c01.y; rc2 = c2.x;
                    rdv[0] = d2.x; rdv[1] = d2.y;
                } else {
                    rc0 = Xn[0] - q.F0; rc1 = Xn[1] - q.F1; rc2 = Xn[2] - q.F2;
                }
                if (mode == BM_NEWTON && k >= 1) gbd += q.g[0] * y0 + q.g[1] * y1 + q.g[2] * y2;
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const double S = Sv[i], vL = vLv[i], vU = vUv[i];
                    const double ds = du[i] + rdv[i];
                    const double sl = S - P.sL[i], su = P.sU[i] - S;
                    const double isl = tpp_rcp(sl), isu = tpp_rcp(su);
                    if (ds != 0.0) a_max = fmin(a_max, tau * ((ds < 0) ? -sl : su) * tpp_rcp(ds));
                    const double dvL = mu * isl - vL - vL * isl * ds;
                    const double dvU = mu * isu - vU + vU * isu * ds;
                    if (dvL < 0) a_z = fmin(a_z, -tau * vL * tpp_rcp(dvL));
                    if (dvU < 0) a_z = fmin(a_z, -tau * vU * tpp_rcp(dvU));
                    if (!isfinite(ds) || !isfinite(dvL) || !isfinite(dvU)) bad = 1;
                    if (mode == BM_NEWTON) gbd += (-mu * isl + mu * isu) * ds + q.g[3 + i] * du[i];
                    if (!(fabs(du[i]) <= TINY_STEP_TOL * (1.0 + fabs(U[i]))) || !(fabs(ds) <= TINY_STEP_TOL * (1.0 + fabs(S)))) big = 1;
                    c2 += rdv[i] * rdv[i];
                }
                c2 += rc0 * rc0 + rc1 * rc1 + rc2 * rc2;
            }
            const double n0 = y0 + q.a13 * y2 + q.b11 * du0 + q.b12 * du1 - rc0;
            const double n1 = y1 + q.a23 * y2 + q.b21 * du0 + q.b22 * du1 - rc1;
            const double n2 = y2 + dt * du1 - rc2;
            y0 = n0; y1 = n1; y2 = n2;
            X[0] = Xn[0]; X[1] = Xn[1]; X[2] = Xn[2];
        }
    }
    o.a_max = a_max; o.a_z = a_z; o.gbd = gbd; o.ymax = ymax; o.bad = bad;
    o.tiny = (!big && sqrt(c2) <= 1e-4) ? 1 : 0;
}

struct TppTrial {
    double th, phi, ymax, dymax; // dymax: largest multiplier step (defect and slack-equality multipliers), step modes
    int bad;
    TppNorms n;
};

// staging of the trial sweep: slots 0-7 = current iterate of stage k, 8-10 = step rows of stage k
__device__ __forceinline__ void tpp_trial_stage(char *sb, const char *p, int co, int srow) {
#pragma unroll
    for (int i = 0; i < R_ITER; i++) tpp_cp16(sb + i * TPP_ROW_B, p + (co + i) * TPP_ROW_B);
#pragma unroll
    for (int i = 0; i < 3; i++) tpp_cp16(sb + (R_ITER + i) * TPP_ROW_B, p + (srow + i) * TPP_ROW_B);
    tpp_cp_commit();
}

// State rows of the Newton system at stage k (1 <= k < N), linearised at the current iterate:
//     Lam_k = A_k' Lam_{k+1} - g_k - (Hxx_k + dw) dx_k - Hxu_k du_k,      Lam = lam + dlam (full-step multipliers).
// lo = multipliers of the defect of stage k+1 at the current iterate (they weight the second derivatives of F),
// useW = 0 for the least-squares multiplier estimate (W = 0, dw = 1, unscaled objective: pass df = 1).
template <int SPEC>
__device__ __forceinline__ void tpp_costate(const KParams &P, const double r[3], const double X[3], const double U[2],
                                            const double lo[3], const double Ln[3], const double dX[3],
                                            const double dU[2], double df, double dw, bool useW, double Lk[3],
                                            const TppObs &oc) {
    const double dt = P.dt, th = X[2], v = U[0], w = U[1];
    double a13, a23, htt, htv, htw = 0;
    if (TPP_IS_EULER(P)) {
        double sn, cs;
        tpp_sincos1(th, &sn, &cs);
        a13 = -dt * v * sn; a23 = dt * v * cs;
        htt = dt * v * (lo[0] * cs + lo[1] * sn);
        htv = dt * (lo[0] * sn - lo[1] * cs);
    } else {
        double s0, c0, sm, cm, se, ce;
        tpp_trig(th, 0.5 * dt * w, s0, c0, sm, cm, se, ce);
        const double h = dt / 6.0;
        const double C = c0 + 4.0 * cm + ce, S = s0 + 4.0 * sm + se;
        const double C1 = 2.0 * cm + ce, S1 = 2.0 * sm + se;
        a13 = -h * v * S; a23 = h * v * C;
        htt = h * v * (lo[0] * C + lo[1] * S);
        htv = h * (lo[0] * S - lo[1] * C);
        htw = h * dt * v * (lo[0] * C1 + lo[1] * S1);
    }
    const double g0 = df * 2.0 * P.Q[0] * (X[0] - r[0]);
    const double g1 = df * 2.0 * P.Q[1] * (X[1] - r[1]);
    const double g2 = df * 2.0 * P.Q[2] * (X[2] - r[2]);
    double hxx = dw, hyy = dw, hth = dw;
    if (useW) {
        hxx += df * 2.0 * P.Q[0];
        hyy += df * 2.0 * P.Q[1];
        hth += df * 2.0 * P.Q[2] + htt;
    } else {
        htv = 0; htw = 0;
    }
    Lk[0] = Ln[0] - g0 - hxx * dX[0];
    Lk[1] = Ln[1] - g1 - hyy * dX[1];
    Lk[2] = Ln[2] + a13 * Ln[0] + a23 * Ln[1] - g2 - hth * dX[2] - htv * dU[0] - htw * dU[1];
    if (TPP_HAS_OBS(P)) {
        // obstacle sum of the stage: gradient, and (Newton system only) its Hessian block
        Lk[0] -= df * oc.gx;
        Lk[1] -= df * oc.gy;
        if (useW) {
            Lk[0] -= df * (oc.hxx * dX[0] + oc.hxy * dX[1]);
            Lk[1] -= df * (oc.hxy * dX[0] + oc.hyy * dX[1]);
        }
    }
}

// ---- sweep T: write curr + alpha*step into the other iterate buffer and evaluate that point -------------------------
// tmode EVAL:        new = current (after the restoration stand-in).
// tmode LSQ:         new = current with the least-squares multiplier estimate (scaled by df; or zeros) for lam, yd.
// tmode STEP(_SOC):  new = current + alpha*(dX,dU,dS,dlam,dyd) and bound multipliers + a_z*(dvL,dvU), clamped.
// The multiplier step dlam is not stored by the other sweeps: it comes from the costate recursion (tpp_costate).
template <int SPEC>
__device__ __forceinline__ void tpp_trial(const KParams &P, const BatchArgs &A, char *wb, char *sb, int cur,
                                          const TppLane &L, TppTrial &o) {
    const int N = P.N;
    const double dt = P.dt, mu = L.mu, df = L.df, dw = L.dw;
    const double *goal = L.goal;
    const int tmode0 = L.tmode, keep0 = L.keep;
    const bool soc0 = (tmode0 == TM_STEP_SOC);
    const double alpha = soc0 ? L.alpha_soc : L.alpha, a_z = soc0 ? L.a_z_soc : L.a_z;
    const int co = cur * R_ITER, no = R_ITER - co;
    const int srow = soc0 ? R_SSTEP : R_STEP;
    const double ikap = 1.0 / KAPPA_SIGMA;
    double Xn[3] = {0, 0, 0}, ln[3] = {0, 0, 0};
    double lo[3] = {0, 0, 0}, Lam[3] = {0, 0, 0}; // multipliers of stage k+1: at the current iterate / after a full step
    double th = 0, slog = 0, pi = 0, di = 0, sy = 0, sz = 0, pmin = 1e300, pmax = -1e300, fs = 0, ymax = 0, dymax = 0;
    int bad = 0;
    tpp_trial_stage(sb, wb + (size_t)N * TPP_STAGE_B, co, srow);
#pragma unroll 1
    for (int k = N; k >= 0; --k) {
        const int mode = tpp_opaque(tmode0);
        const bool step = (mode >= TM_STEP), soc = (mode == TM_STEP_SOC);
        const bool lsq = (mode == TM_LSQ) && keep0;
        char *p = wb + (size_t)k * TPP_STAGE_B;
        char *pw = p + no * TPP_ROW_B;
        tpp_cp_wait();
        const double2 x01 = tpp_sld(sb, R_X01), x2l0 = tpp_sld(sb, R_X2L0), l12 = tpp_sld(sb, R_L12), u2 = tpp_sld(sb, R_U);
        const double2 s2 = tpp_sld(sb, R_S), yd2 = tpp_sld(sb, R_YD), vl2 = tpp_sld(sb, R_VL), vu2 = tpp_sld(sb, R_VU);
        const double2 dx01 = tpp_sld(sb, R_ITER), dx2 = tpp_sld(sb, R_ITER + 1), du2 = tpp_sld(sb, R_ITER + 2);
        tpp_consume(x01, x2l0, l12, u2);
        tpp_consume(s2, yd2, vl2, vu2);
        tpp_consume(dx01, dx2, du2, du2);
        if (k > 0) tpp_trial_stage(sb, p - TPP_STAGE_B, co, srow);
        if (k > 1) {
            tpp_l2_prefetch(p - 2 * TPP_STAGE_B + co * TPP_ROW_B, 0, R_ITER);
            tpp_l2_prefetch(p - 2 * TPP_STAGE_B, srow, 3);
        }
        double r[3], ub[2];
        tpp_ref<SPEC>(P, goal, p, r, ub);
        double X[3] = {x01.x, x01.y, x2l0.x};
        double lam[3] = {0, 0, 0};
        const double duv[2] = {du2.x, du2.y};
        // obstacle sums of the stage: at the current iterate (costate recursion) and at the point being evaluated (the
        // obstacle block has put them into the cache rows of the other buffer)
        TppObs occ = tpp_obs_zero(), oct = tpp_obs_zero();
        if (TPP_HAS_OBS(P)) {
            if (step || lsq) occ = tpp_obs_load(p, cur);
            oct = tpp_obs_load(p, 1 - cur);
            if (k > 1) tpp_l2_prefetch(p - 2 * TPP_STAGE_B, R_OC, 6);
        }
        if (k >= 1) {
            lam[0] = x2l0.y; lam[1] = l12.x; lam[2] = l12.y;
            const double lcur[3] = {lam[0], lam[1], lam[2]};
            const double dX[3] = {dx01.x, dx01.y, dx2.x};
            if (step || lsq) {
                const double dwk = lsq ? 1.0 : dw;
                double Lk[3];
                if (k == N) {
                    // terminal stage: no tracking cost, no controls
                    Lk[0] = -dwk * dX[0]; Lk[1] = -dwk * dX[1]; Lk[2] = -dwk * dX[2];
                    if (TPP_HAS_OBS(P)) {
                        const double dfv = lsq ? 1.0 : df;
                        Lk[0] -= dfv * occ.gx;
                        Lk[1] -= dfv * occ.gy;
                        if (!lsq) {
                            Lk[0] -= dfv * (occ.hxx * dX[0] + occ.hxy * dX[1]);
                            Lk[1] -= dfv * (occ.hxy * dX[0] + occ.hyy * dX[1]);
                        }
                    }
                } else {
                    const double Uc[2] = {u2.x, u2.y};
                    tpp_costate<SPEC>(P, r, X, Uc, lo, Lam, dX, duv, lsq ? 1.0 : df, dwk, !lsq, Lk, occ);
                }
                Lam[0] = Lk[0]; Lam[1] = Lk[1]; Lam[2] = Lk[2];
                if (!isfinite(Lk[0]) || !isfinite(Lk[1]) || !isfinite(Lk[2])) bad = 1;
                if (step) {
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        X[i] = fma(alpha, dX[i], X[i]); // (the obstacle block forms the same point with the same operation)
                        const double dl = Lk[i] - lam[i];
                        dymax = fmax(dymax, fabs(dl));
                        lam[i] += alpha * dl;
                    }
                } else {
                    ymax = fmax(ymax, fmax(fabs(Lk[0]), fmax(fabs(Lk[1]), fabs(Lk[2]))));
#pragma unroll
                    for (int i = 0; i < 3; i++) lam[i] = df * Lk[i];
                }
            } else if (mode == TM_LSQ) {
                lam[0] = lam[1] = lam[2] = 0.0;
            }
            lo[0] = lcur[0]; lo[1] = lcur[1]; lo[2] = lcur[2];
        }
        tpp_st2(pw, R_X01, X[0], X[1]); tpp_st2(pw, R_X2L0, X[2], lam[0]); tpp_st2(pw, R_L12, lam[1], lam[2]);
        if (k < N) {
            double U[2] = {u2.x, u2.y}, S[2] = {s2.x, s2.y}, yd[2] = {yd2.x, yd2.y}, vL[2] = {vl2.x, vl2.y}, vU[2] = {vu2.x, vu2.y};
            double rdv[2] = {U[0] - S[0], U[1] - S[1]};
            if (soc) {
                const double2 d2 = tpp_ld2(p, R_CS + 2);
                rdv[0] = d2.x; rdv[1] = d2.y;
            }
#pragma unroll
            for (int i = 0; i < 2; i++) {
                if (step) {
                    const double du = duv[i];
                    const double ds = du + rdv[i];
                    const double isl = tpp_rcp(S[i] - P.sL[i]), isu = tpp_rcp(P.sU[i] - S[i]);
                    const double Dsig = vL[i] * isl + vU[i] * isu + dw;
                    const double rs = -yd[i] - mu * isl + mu * isu;
                    const double dvL = mu * isl - vL[i] - vL[i] * isl * ds;
                    const double dvU = mu * isu - vU[i] + vU[i] * isu * ds;
                    U[i] += alpha * du;
                    S[i] += alpha * ds;
                    const double dyd = Dsig * ds + rs;
                    dymax = fmax(dymax, fabs(dyd));
                    yd[i] += alpha * dyd;
                    vL[i] += a_z * dvL;
                    vU[i] += a_z * dvU;
                    const double ml = mu * tpp_rcp(S[i] - P.sL[i]), mu_u = mu * tpp_rcp(P.sU[i] - S[i]);
                    vL[i] = fmax(fmin(vL[i], KAPPA_SIGMA * ml), ml * ikap);
                    vU[i] = fmax(fmin(vU[i], KAPPA_SIGMA * mu_u), mu_u * ikap);
                } else if (mode == TM_LSQ) {
                    yd[i] = keep0 ? df * duv[i] : 0.0;
                }
            }
            tpp_st2(pw, R_U, U[0], U[1]); tpp_st2(pw, R_S, S[0], S[1]); tpp_st2(pw, R_YD, yd[0], yd[1]);
            tpp_st2(pw, R_VL, vL[0], vL[1]); tpp_st2(pw, R_VU, vU[0], vU[1]);
            TppLin q;
            tpp_lin<false, SPEC>(P, r, ub, X, U, ln, df, q, oct);
            const double c[3] = {Xn[0] - q.F0, Xn[1] - q.F1, Xn[2] - q.F2};
            fs += q.f;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                th += fabs(c[i]);
                pi = fmax(pi, fabs(c[i]));
            }
            if (k >= 1) {
                const double r0 = q.g[0] + lam[0] - ln[0];
                const double r1 = q.g[1] + lam[1] - ln[1];
                const double r2 = q.g[2] + lam[2] - (q.a13 * ln[0] + q.a23 * ln[1] + ln[2]);
                di = fmax(di, fmax(fabs(r0), fmax(fabs(r1), fabs(r2))));
                sy += fabs(lam[0]) + fabs(lam[1]) + fabs(lam[2]);
            }
            const double ru0 = q.g[3] - (q.b11 * ln[0] + q.b21 * ln[1]) + yd[0];
            const double ru1 = q.g[4] - (q.b12 * ln[0] + q.b22 * ln[1] + dt * ln[2]) + yd[1];
            di = fmax(di, fmax(fabs(ru0), fabs(ru1)));
            double prod = 1.0;
            bool inside = true;
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const double rd = U[i] - S[i];
                const double sl = S[i] - P.sL[i], su = P.sU[i] - S[i];
                th += fabs(rd);
                pi = fmax(pi, fabs(rd));
                prod *= sl * su;
                inside = inside && (sl > 0.0) && (su > 0.0);
                di = fmax(di, fabs(-yd[i] - vL[i] + vU[i]));
                sy += fabs(yd[i]);
                sz += fabs(vL[i]) + fabs(vU[i]);
                const double pl = sl * vL[i], pu = su * vU[i];
                pmin = fmin(pmin, fmin(pl, pu));
                pmax = fmax(pmax, fmax(pl, pu));
            }
            // barrier term: one logarithm per stage; a slack outside its bounds gives NaN like log(negative) would
            slog += log(inside ? prod : -1.0);
        } else if (k >= 1) {
            // terminal state: no tracking cost; its stationarity residual is the multiplier itself (plus the gradient of
            // the obstacle sum, where the terminal stage carries it)
            double r0 = lam[0], r1 = lam[1];
            if (TPP_HAS_OBS(P)) {
                r0 = fma(df, oct.gx, r0);
                r1 = fma(df, oct.gy, r1);
                fs += oct.v;
            }
            di = fmax(di, fmax(fabs(r0), fmax(fabs(r1), fabs(lam[2]))));
            sy += fabs(lam[0]) + fabs(lam[1]) + fabs(lam[2]);
        }
        Xn[0] = X[0]; Xn[1] = X[1]; Xn[2] = X[2];
        ln[0] = lam[0]; ln[1] = lam[1]; ln[2] = lam[2];
    }
    o.th = th;
    o.phi = df * fs - mu * slog;
    o.ymax = ymax; o.dymax = dymax; o.bad = bad;
    o.n.theta = th; o.n.prim_inf = pi; o.n.dual_inf = di; o.n.sum_y = sy; o.n.sum_z = sz;
    o.n.pmin = pmin; o.n.pmax = pmax; o.n.f = fs; o.n.slog = slog;
}

// ---- obstacle block: the obstacle sums the next sweep needs, for the lanes in `mask` ---------------------------------------
// One problem at a time, by the whole warp with lane <-> stage (obstacle_eval of b200mpc.cu, the routine of the warp kernel):
// the problem's obstacle list is read straight from the caller's arrays (every lane reads the same entry: one broadcast
// transaction, served by L1 after the first stage touched it), the stage's position comes from the lane's column of the
// workspace,  point = X of buffer `cur` (+ alpha * the step in rows `srow`, when srow >= 0),  and value / gradient / Hessian
// go to the cache rows of buffer `dst`.  A lane-per-problem walk would have to re-read each lane's 2.5 KB list for every
// stage (or tile the stages); this way a list is read once per evaluation, the padding fold (n_eff) is per problem, and
// lanes that are not about to evaluate a point cost nothing.
// copy_rows: the point does not move (multiplier estimate): the cache rows are copied instead.
// the first n entries of a problem's obstacle list -> the warp's list buffer in shared memory (olist; NULL: stay in global)
__device__ __forceinline__ void tpp_stage_list(const double *&gx, const double *&gy, int n, double *olist, int Mpad) {
    if (!olist) return;
    const int lane = threadIdx.x & 31;
    __syncwarp();
    for (int i = lane; i < n; i += 32) { olist[i] = gx[i]; olist[Mpad + i] = gy[i]; }
    __syncwarp();
    gx = olist; gy = olist + Mpad;
}

template <int SPEC>
__device__ __noinline__ void tpp_obstacle_block(const KParams &P, const BatchArgs &A, char *wbase, unsigned mask, int cur,
                                                int dst, int b_mine, int neff_mine, double alpha_mine, int srow_mine,
                                                int copy_mine, double *olist) {
    const int lane = threadIdx.x & 31, N = P.N;
    const int co = cur * R_ITER;
    const int Mpad = (P.M + 3) & ~3;
    for (unsigned m = mask; m; m &= m - 1) {
        const int j = __ffs(m) - 1;
        const size_t b = (size_t)__shfl_sync(FULL, b_mine, j);
        const int ne = __shfl_sync(FULL, neff_mine, j);
        const double al = __shfl_sync(FULL, alpha_mine, j);
        const int sr = __shfl_sync(FULL, srow_mine, j);
        const int cp = __shfl_sync(FULL, copy_mine, j);
        const double *gx = A.ox + (size_t)A.obs_stride * b, *gy = A.oy + (size_t)A.obs_stride * b;
        if (const unsigned rest = m & (m - 1)) {
            // the next problem's list on its way into L2 while this one is evaluated (lanes 0-15: x, 16-31: y; 128 B each)
            const size_t bn = (size_t)__shfl_sync(FULL, b_mine, __ffs(rest) - 1);
            const double *pn = ((lane < 16) ? A.ox : A.oy) + (size_t)A.obs_stride * bn + (lane & 15) * 16;
            if ((lane & 15) * 16 < P.M) asm volatile("prefetch.global.L2 [%0];" ::"l"(pn));
        }
        if (!cp) tpp_stage_list(gx, gy, ne, olist, Mpad);
        const double w0 = 1.0 + (double)(P.M - ne);
        for (int k = lane; k <= N; k += 32) {
            char *p = wbase + (size_t)k * TPP_STAGE_B + j * 16;
            if (cp) {
                const double2 a = tpp_ld2(p, R_OC + 3 * cur), c = tpp_ld2(p, R_OC + 3 * cur + 1), d = tpp_ld2(p, R_OC + 3 * cur + 2);
                tpp_st2(p, R_OC + 3 * dst, a.x, a.y); tpp_st2(p, R_OC + 3 * dst + 1, c.x, c.y); tpp_st2(p, R_OC + 3 * dst + 2, d.x, d.y);
                continue;
            }
            const double2 x = tpp_ld2(p + co * TPP_ROW_B, R_X01);
            double px = x.x, py = x.y;
            if (sr >= 0) {
                const double2 d = tpp_ld2(p, sr);
                px = fma(al, d.x, px);
                py = fma(al, d.y, py);
            }
            double o6[6] = {0, 0, 0, 0, 0, 0};
            if (k >= P.obs_k0 && k <= P.obs_k1) obstacle_eval(P.obs_form, P.obs_c, P.inv_r2, gx, gy, ne, w0, px, py, o6);
            tpp_st2(p, R_OC + 3 * dst, o6[0], o6[1]); tpp_st2(p, R_OC + 3 * dst + 1, o6[2], o6[3]);
            tpp_st2(p, R_OC + 3 * dst + 2, o6[4], o6[5]);
        }
    }
    __syncwarp();
}

// ---- filter (32 entries per lane, kept in the workspace) --------------------------------------------------------------
__device__ __forceinline__ bool tpp_filter_ok(const double *fl, unsigned mask, double phi, double th) {
    while (mask) {
        const int i = __ffs(mask) - 1;
        mask &= mask - 1;
        const double fp = fl[(2 * i) * 32], ft = fl[(2 * i + 1) * 32];
        if (!(cmp_le(phi, fp, fp) || cmp_le(th, ft, ft))) return false;
    }
    return true;
}

__device__ __forceinline__ void tpp_filter_add(double *fl, TppLane &L, double phi, double th) {
    unsigned mask = L.fmask;
    unsigned m = mask;
    while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1;
        if (fl[(2 * i) * 32] >= phi && fl[(2 * i + 1) * 32] >= th) mask &= ~(1u << i); // dominated by the new entry
    }
    int slot;
    if (~mask) slot = __ffs(~mask) - 1;
    else { slot = L.ring & 31; L.ring++; }
    fl[(2 * slot) * 32] = phi;
    fl[(2 * slot + 1) * 32] = th;
    L.fmask = mask | (1u << slot);
}

// FilterLSAcceptor::CheckAcceptabilityOfTrialPoint (same tests as ls_acceptable of the warp kernel)
__device__ __forceinline__ bool tpp_ls_acceptable(const TppLane &L, const double *fl, double alpha_test, double phi_t,
                                                  double th_t, bool &ftype_armijo) {
    ftype_armijo = false;
    if (!isfinite(th_t) || !isfinite(phi_t)) return false;
    if (th_t > 1e4 * L.theta0) return false;
    const double gbd = L.ref_gbd, theta = L.theta, phi = L.ref_phi;
    const bool ftype = (gbd < 0) && (alpha_test * tpp_pow(-gbd, S_PHI) > DELTA_LS * tpp_pow(theta, S_THETA));
    const bool armijo = cmp_le(phi_t - phi, ETA_PHI * alpha_test * gbd, phi);
    ftype_armijo = ftype && armijo;
    bool ok;
    if (alpha_test > 0 && ftype && theta <= 1e-4 * L.theta0) {
        ok = armijo;
    } else {
        if (phi_t > phi) {
            double bas = 1.0;
            if (fabs(phi) > 10.0) bas = tpp_log10(fabs(phi));
            if (tpp_log10(phi_t - phi) > OBJ_MAX_INC + bas) return false;
        }
        ok = cmp_le(th_t, (1 - GAMMA_THETA) * theta, theta) || cmp_le(phi_t - phi, -GAMMA_PHI * theta, phi);
    }
    if (!ok) return false;
    return tpp_filter_ok(fl, L.fmask, phi_t, th_t);
}

// Restoration stand-in (same rule as the warp kernel and the oracle).  The line search gave up on the Newton direction
// (alpha < alpha_min).  The filter is augmented with the current point, then
//   stage 1: along the direction to the closed-form feasible point (U = S, X rolled out; S fixed) the longest step
//            t = 1, 1/2, ... is taken whose point passes IPOPT's restoration acceptance (finite, theta <= kappa_resto * theta,
//            no excessive objective increase, acceptable to the augmented filter);
//   stage 2: if there is none, the family "plan shrunk towards standing still": S_l = u_c + l (S - u_c), U = S_l, X rolled out,
//            l = 1/2, 1/4, ..., 0; the member with the lowest barrier objective is taken.
// The multipliers restart.
// tpp_resto_eval: theta, objective and log-barrier sum of  current + t * direction  (direction (dX, dU) in the R_SSTEP rows;
// the slack step is dU + (U - S) when move_s, else the slacks stay).
template <int SPEC>
__device__ __noinline__ void tpp_resto_eval(const KParams &P, const char *wb, int co, const double *goal, double t, int move_s,
                                            double *th_out, double *f_out, double *slog_out) {
    const int N = P.N;
    double th = 0, fs = 0, slog = 0;
    double X[3];
    {
        const char *pc = wb + co * TPP_ROW_B;
        const double2 a = tpp_ld2(pc, R_X01), b = tpp_ld2(pc, R_X2L0);
        X[0] = a.x; X[1] = a.y; X[2] = b.x; // stage 0 does not move
    }
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
        const char *p = wb + (size_t)k * TPP_STAGE_B;
        const char *pc = p + co * TPP_ROW_B;
        const double2 u2 = tpp_ld2(pc, R_U), s2 = tpp_ld2(pc, R_S), du2 = tpp_ld2(p, R_SSTEP + 2);
        const double2 n01 = tpp_ld2(pc + TPP_STAGE_B, R_X01), n2 = tpp_ld2(pc + TPP_STAGE_B, R_X2L0);
        const double2 d01 = tpp_ld2(p + TPP_STAGE_B, R_SSTEP), d2 = tpp_ld2(p + TPP_STAGE_B, R_SSTEP + 1);
        const double U[2] = {u2.x + t * du2.x, u2.y + t * du2.y};
        double S[2] = {s2.x, s2.y};
        if (move_s) {
            S[0] = s2.x + t * (du2.x + (u2.x - s2.x));
            S[1] = s2.y + t * (du2.y + (u2.y - s2.y));
            const double l0 = S[0] - P.sL[0], l1 = P.sU[0] - S[0], l2 = S[1] - P.sL[1], l3 = P.sU[1] - S[1];
            const bool inside = (l0 > 0.0) && (l1 > 0.0) && (l2 > 0.0) && (l3 > 0.0);
            slog += log(inside ? (l0 * l1) * (l2 * l3) : -1.0);
        }
        const double Xn[3] = {n01.x + t * d01.x, n01.y + t * d01.y, n2.x + t * d2.x};
        double r[3], ub[2];
        tpp_ref<SPEC>(P, goal, p, r, ub);
        const double ln0[3] = {0, 0, 0};
        TppLin q;
        tpp_lin<false, SPEC>(P, r, ub, X, U, ln0, 1.0, q, tpp_obs_zero());
        th += fabs(Xn[0] - q.F0) + fabs(Xn[1] - q.F1) + fabs(Xn[2] - q.F2) + fabs(U[0] - S[0]) + fabs(U[1] - S[1]);
        fs += q.f;
        X[0] = Xn[0]; X[1] = Xn[1]; X[2] = Xn[2];
    }
    *th_out = th;
    *f_out = fs;
    *slog_out = slog;
}

// restoration direction -> R_SSTEP rows: (Xr - X) with Xr the roll-out of the target slacks S_l, (S_l - U);
// lam < 0: S_l = S exactly (stage 1).  Returns the largest bound multiplier.
template <int SPEC>
__device__ __noinline__ double tpp_resto_direction(const KParams &P, char *wb, int co, double lam, double uc0, double uc1) {
    const int N = P.N;
    double zm = 0;
    double y[3];
    const double2 a = tpp_ld2(wb + co * TPP_ROW_B, R_X01), b = tpp_ld2(wb + co * TPP_ROW_B, R_X2L0);
    y[0] = a.x; y[1] = a.y; y[2] = b.x;
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
        char *p = wb + (size_t)k * TPP_STAGE_B;
        const char *pc = p + co * TPP_ROW_B;
        const double2 x01 = tpp_ld2(pc, R_X01), x2 = tpp_ld2(pc, R_X2L0);
        tpp_st2(p, R_SSTEP, (k == 0) ? 0.0 : y[0] - x01.x, (k == 0) ? 0.0 : y[1] - x01.y);
        tpp_st2(p, R_SSTEP + 1, (k == 0) ? 0.0 : y[2] - x2.x, 0.0);
        if (k < N) {
            const double2 u = tpp_ld2(pc, R_U), sv = tpp_ld2(pc, R_S), vl = tpp_ld2(pc, R_VL), vu = tpp_ld2(pc, R_VU);
            zm = fmax(zm, fmax(fmax(vl.x, vl.y), fmax(vu.x, vu.y)));
            double S[2] = {sv.x, sv.y};
            if (lam >= 0.0) { S[0] = uc0 + lam * (sv.x - uc0); S[1] = uc1 + lam * (sv.y - uc1); }
            tpp_st2(p, R_SSTEP + 2, S[0] - u.x, S[1] - u.y);
            double F[3];
            tpp_dyn<SPEC>(P, y, S, F);
            y[0] = F[0]; y[1] = F[1]; y[2] = F[2];
        }
    }
    return zm;
}

template <int SPEC>
__device__ __forceinline__ void tpp_restore(const KParams &P, char *wb, double *fl, int cur, TppLane &L) {
    const int N = P.N;
    if (L.theta <= 1e-10 || L.n_resto >= MAX_RESTO) { L.status = B200MPC_RESTORATION_FAILED; L.phase = PH_FIN; return; }
    tpp_filter_add(fl, L, L.ref_phi - GAMMA_PHI * L.theta, (1 - GAMMA_THETA) * L.theta);
    const int co = cur * R_ITER, no = R_ITER - co;
    const double zm = tpp_resto_direction<SPEC>(P, wb, co, -1.0, 0.0, 0.0);
    double t_acc = 0.0;
    int move_s = 0;
    const double phi_ref = L.ref_phi;
    for (double t = 1.0; t >= RESTO_T_MIN; t *= 0.5) {
        double th_r, f_r, sl_r;
        tpp_resto_eval<SPEC>(P, wb, co, L.goal, t, 0, &th_r, &f_r, &sl_r);
        const double phi_r = L.df * f_r - L.mu * L.slog; // the slacks do not move: the barrier term is the current one
        if (!isfinite(th_r) || !isfinite(phi_r)) continue;
        if (!(th_r <= KAPPA_RESTO * L.theta)) continue;
        if (phi_r > phi_ref) {
            double bas = 1.0;
            if (fabs(phi_ref) > 10.0) bas = tpp_log10(fabs(phi_ref));
            if (tpp_log10(phi_r - phi_ref) > OBJ_MAX_INC + bas) continue;
        }
        if (!tpp_filter_ok(fl, L.fmask, phi_r, th_r)) continue;
        t_acc = t;
        break;
    }
    if (t_acc == 0.0) {
        double uc[2];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const double lo = P.sL[i], hi = P.sU[i];
            const double pl = fmin(BOUND_PUSH * fmax(1.0, fabs(lo)), BOUND_FRAC * (hi - lo));
            const double pu = fmin(BOUND_PUSH * fmax(1.0, fabs(hi)), BOUND_FRAC * (hi - lo));
            uc[i] = fmin(fmax(0.0, lo + pl), hi - pu);
        }
        double lam = 0.5, best = __longlong_as_double(0x7ff0000000000000ll), lam_best = -1.0;
        for (;;) {
            (void)tpp_resto_direction<SPEC>(P, wb, co, lam, uc[0], uc[1]);
            double th_r, f_r, sl_r;
            tpp_resto_eval<SPEC>(P, wb, co, L.goal, 1.0, 1, &th_r, &f_r, &sl_r);
            const double phi_r = L.df * f_r - L.mu * sl_r;
            if (isfinite(th_r) && isfinite(phi_r) && (th_r <= KAPPA_RESTO * L.theta) && (phi_r < best) &&
                tpp_filter_ok(fl, L.fmask, phi_r, th_r)) {
                best = phi_r; lam_best = lam;
            }
            if (lam == 0.0) break;
            lam = (lam * 0.5 >= RESTO_T_MIN) ? lam * 0.5 : 0.0;
        }
        if (lam_best >= 0.0) {
            (void)tpp_resto_direction<SPEC>(P, wb, co, lam_best, uc[0], uc[1]);
            t_acc = 1.0;
            move_s = 1;
        }
    }
    if (t_acc == 0.0) { L.status = B200MPC_RESTORATION_FAILED; L.phase = PH_FIN; return; }
    const bool reset = zm > 1e3;
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
        const char *p = wb + (size_t)k * TPP_STAGE_B;
        const char *pc = p + co * TPP_ROW_B;
        char *pw = wb + (size_t)k * TPP_STAGE_B + no * TPP_ROW_B;
        const double2 x01 = tpp_ld2(pc, R_X01), x2 = tpp_ld2(pc, R_X2L0), d01 = tpp_ld2(p, R_SSTEP), d2 = tpp_ld2(p, R_SSTEP + 1);
        tpp_st2(pw, R_X01, x01.x + t_acc * d01.x, x01.y + t_acc * d01.y);
        tpp_st2(pw, R_X2L0, x2.x + t_acc * d2.x, 0.0);
        tpp_st2(pw, R_L12, 0.0, 0.0);
        if (k < N) {
            const double2 u = tpp_ld2(pc, R_U), du = tpp_ld2(p, R_SSTEP + 2), s = tpp_ld2(pc, R_S), a = tpp_ld2(pc, R_VL), b = tpp_ld2(pc, R_VU);
            double S[2] = {s.x, s.y};
            if (move_s) { S[0] = s.x + t_acc * (du.x + (u.x - s.x)); S[1] = s.y + t_acc * (du.y + (u.y - s.y)); }
            tpp_st2(pw, R_U, u.x + t_acc * du.x, u.y + t_acc * du.y); tpp_st2(pw, R_S, S[0], S[1]); tpp_st2(pw, R_YD, 0.0, 0.0);
            tpp_st2(pw, R_VL, reset ? 1.0 : a.x, reset ? 1.0 : a.y);
            tpp_st2(pw, R_VU, reset ? 1.0 : b.x, reset ? 1.0 : b.y);
        }
    }
    L.moved = 1;
    L.n_resto++;
    L.iter++;
    L.tmode = TM_EVAL;
    L.phase = PH_T;
}

// Restoration stand-in of the instances with the obstacle cost: the same rule, but every candidate needs the obstacle sums of
// all stages, so the WARP evaluates the candidates of one problem at a time (lane <-> stage, sums by butterfly reductions)
// while the problem's own lane builds the directions, keeps its filter and takes the decisions.
// tpp_resto_eval_coop: theta, objective and log-barrier sum of  current + t * direction  of the problem in column wbj.
template <int SPEC>
__device__ __forceinline__ void tpp_resto_eval_coop(const KParams &P, const double *gx, const double *gy, int n_eff, const char *wbj,
                                                    int co, const double goal[3], double t, int move_s, double &th_out,
                                                    double &f_out, double &slog_out) {
    const int lane = threadIdx.x & 31, N = P.N;
    const double w0 = 1.0 + (double)(P.M - n_eff);
    double th = 0, fs = 0, slog = 0;
    for (int k = lane; k <= N; k += 32) {
        const char *p = wbj + (size_t)k * TPP_STAGE_B;
        const char *pc = p + co * TPP_ROW_B;
        const double2 x01 = tpp_ld2(pc, R_X01), x2 = tpp_ld2(pc, R_X2L0), dx01 = tpp_ld2(p, R_SSTEP), dx2 = tpp_ld2(p, R_SSTEP + 1);
        const double X[3] = {x01.x + t * dx01.x, x01.y + t * dx01.y, x2.x + t * dx2.x}; // (the direction rows of stage 0 are zero)
        TppObs oc = tpp_obs_zero();
        if (k >= P.obs_k0 && k <= P.obs_k1) {
            double o6[6];
            obstacle_eval(P.obs_form, P.obs_c, P.inv_r2, gx, gy, n_eff, w0, X[0], X[1], o6);
            oc.v = o6[0];
        }
        if (k < N) {
            const double2 u2 = tpp_ld2(pc, R_U), s2 = tpp_ld2(pc, R_S), du2 = tpp_ld2(p, R_SSTEP + 2);
            const double2 n01 = tpp_ld2(pc + TPP_STAGE_B, R_X01), n2 = tpp_ld2(pc + TPP_STAGE_B, R_X2L0);
            const double2 d01 = tpp_ld2(p + TPP_STAGE_B, R_SSTEP), d2 = tpp_ld2(p + TPP_STAGE_B, R_SSTEP + 1);
            const double U[2] = {u2.x + t * du2.x, u2.y + t * du2.y};
            double S[2] = {s2.x, s2.y};
            if (move_s) {
                S[0] = s2.x + t * (du2.x + (u2.x - s2.x));
                S[1] = s2.y + t * (du2.y + (u2.y - s2.y));
                const double l0 = S[0] - P.sL[0], l1 = P.sU[0] - S[0], l2 = S[1] - P.sL[1], l3 = P.sU[1] - S[1];
                const bool inside = (l0 > 0.0) && (l1 > 0.0) && (l2 > 0.0) && (l3 > 0.0);
                slog += log(inside ? (l0 * l1) * (l2 * l3) : -1.0);
            }
            const double Xn[3] = {n01.x + t * d01.x, n01.y + t * d01.y, n2.x + t * d2.x};
            double r[3], ub[2];
            tpp_ref<SPEC>(P, goal, p, r, ub);
            const double ln0[3] = {0, 0, 0};
            TppLin q;
            tpp_lin<false, SPEC>(P, r, ub, X, U, ln0, 1.0, q, oc);
            th += fabs(Xn[0] - q.F0) + fabs(Xn[1] - q.F1) + fabs(Xn[2] - q.F2) + fabs(U[0] - S[0]) + fabs(U[1] - S[1]);
            fs += q.f;
        } else {
            fs += oc.v;
        }
    }
    __syncwarp();
    th_out = wsum(th); f_out = wsum(fs); slog_out = wsum(slog);
}

template <int SPEC>
__device__ __noinline__ void tpp_restore_obs(const KParams &P, const BatchArgs &A, char *wbase, double *fl, int cur, TppLane &L,
                                             unsigned mask, double *olist) {
    const int lane = threadIdx.x & 31, N = P.N;
    const int Mpad = (P.M + 3) & ~3;
    const int co = cur * R_ITER, no = R_ITER - co;
    double uc[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const double lo = P.sL[i], hi = P.sU[i];
        const double pl = fmin(BOUND_PUSH * fmax(1.0, fabs(lo)), BOUND_FRAC * (hi - lo));
        const double pu = fmin(BOUND_PUSH * fmax(1.0, fabs(hi)), BOUND_FRAC * (hi - lo));
        uc[i] = fmin(fmax(0.0, lo + pl), hi - pu);
    }
    for (unsigned m = mask; m; m &= m - 1) {
        const int j = __ffs(m) - 1;
        char *wbj = wbase + j * 16; // the problem's column of the workspace
        int go = 0;
        double zm = 0;
        if (lane == j) {
            if (L.theta <= 1e-10 || L.n_resto >= MAX_RESTO) { L.status = B200MPC_RESTORATION_FAILED; L.phase = PH_FIN; }
            else {
                tpp_filter_add(fl, L, L.ref_phi - GAMMA_PHI * L.theta, (1 - GAMMA_THETA) * L.theta);
                zm = tpp_resto_direction<SPEC>(P, wbj, co, -1.0, 0.0, 0.0);
                go = 1;
            }
        }
        __syncwarp();
        if (!__shfl_sync(FULL, go, j)) continue;
        const size_t pb = (size_t)__shfl_sync(FULL, L.b, j);
        const int ne = __shfl_sync(FULL, L.n_eff, j);
        const double theta = __shfl_sync(FULL, L.theta, j), df = __shfl_sync(FULL, L.df, j), mu = __shfl_sync(FULL, L.mu, j);
        const double slog_cur = __shfl_sync(FULL, L.slog, j), phi_ref = __shfl_sync(FULL, L.ref_phi, j);
        const double goal[3] = {__shfl_sync(FULL, L.goal[0], j), __shfl_sync(FULL, L.goal[1], j), __shfl_sync(FULL, L.goal[2], j)};
        const double *gx = A.ox + (size_t)A.obs_stride * pb, *gy = A.oy + (size_t)A.obs_stride * pb;
        tpp_stage_list(gx, gy, ne, olist, Mpad);
        double t_acc = 0.0;
        int move_s = 0;
        for (double t = 1.0; t >= RESTO_T_MIN; t *= 0.5) {
            double th_r, f_r, sl_r;
            tpp_resto_eval_coop<SPEC>(P, gx, gy, ne, wbj, co, goal, t, 0, th_r, f_r, sl_r);
            const double phi_r = df * f_r - mu * slog_cur; // the slacks do not move: the barrier term is the current one
            int ok = 0;
            if (lane == j) {
                ok = isfinite(th_r) && isfinite(phi_r) && (th_r <= KAPPA_RESTO * theta);
                if (ok && phi_r > phi_ref) {
                    double bas = 1.0;
                    if (fabs(phi_ref) > 10.0) bas = tpp_log10(fabs(phi_ref));
                    if (tpp_log10(phi_r - phi_ref) > OBJ_MAX_INC + bas) ok = 0;
                }
                if (ok && !tpp_filter_ok(fl, L.fmask, phi_r, th_r)) ok = 0;
            }
            if (__shfl_sync(FULL, ok, j)) { t_acc = t; break; }
        }
        if (t_acc == 0.0) {
            double lam = 0.5, best = __longlong_as_double(0x7ff0000000000000ll), lam_best = -1.0;
            for (;;) {
                if (lane == j) (void)tpp_resto_direction<SPEC>(P, wbj, co, lam, uc[0], uc[1]);
                __syncwarp();
                double th_r, f_r, sl_r;
                tpp_resto_eval_coop<SPEC>(P, gx, gy, ne, wbj, co, goal, 1.0, 1, th_r, f_r, sl_r);
                const double phi_r = df * f_r - mu * sl_r;
                int ok = 0;
                if (lane == j)
                    ok = isfinite(th_r) && isfinite(phi_r) && (th_r <= KAPPA_RESTO * theta) && (phi_r < best) &&
                         tpp_filter_ok(fl, L.fmask, phi_r, th_r);
                if (__shfl_sync(FULL, ok, j)) { best = phi_r; lam_best = lam; }
                if (lam == 0.0) break;
                lam = (lam * 0.5 >= RESTO_T_MIN) ? lam * 0.5 : 0.0;
            }
            if (lam_best >= 0.0) {
                if (lane == j) (void)tpp_resto_direction<SPEC>(P, wbj, co, lam_best, uc[0], uc[1]);
                __syncwarp();
                t_acc = 1.0;
                move_s = 1;
            }
        }
        if (t_acc == 0.0) {
            if (lane == j) { L.status = B200MPC_RESTORATION_FAILED; L.phase = PH_FIN; }
            continue;
        }
        // the accepted point goes to the other iterate buffer (stage-parallel), its obstacle sums to that buffer's cache rows
        const bool reset = __shfl_sync(FULL, zm, j) > 1e3;
        for (int k = lane; k <= N; k += 32) {
            const char *p = wbj + (size_t)k * TPP_STAGE_B;
            const char *pc = p + co * TPP_ROW_B;
            char *pw = wbj + (size_t)k * TPP_STAGE_B + no * TPP_ROW_B;
            const double2 x01 = tpp_ld2(pc, R_X01), x2 = tpp_ld2(pc, R_X2L0), d01 = tpp_ld2(p, R_SSTEP), d2 = tpp_ld2(p, R_SSTEP + 1);
            tpp_st2(pw, R_X01, fma(t_acc, d01.x, x01.x), fma(t_acc, d01.y, x01.y));
            tpp_st2(pw, R_X2L0, x2.x + t_acc * d2.x, 0.0);
            tpp_st2(pw, R_L12, 0.0, 0.0);
            if (k < N) {
                const double2 u = tpp_ld2(pc, R_U), du = tpp_ld2(p, R_SSTEP + 2), sv = tpp_ld2(pc, R_S), a = tpp_ld2(pc, R_VL), c = tpp_ld2(pc, R_VU);
                double S[2] = {sv.x, sv.y};
                if (move_s) { S[0] = sv.x + t_acc * (du.x + (u.x - sv.x)); S[1] = sv.y + t_acc * (du.y + (u.y - sv.y)); }
                tpp_st2(pw, R_U, u.x + t_acc * du.x, u.y + t_acc * du.y); tpp_st2(pw, R_S, S[0], S[1]); tpp_st2(pw, R_YD, 0.0, 0.0);
                tpp_st2(pw, R_VL, reset ? 1.0 : a.x, reset ? 1.0 : a.y);
                tpp_st2(pw, R_VU, reset ? 1.0 : c.x, reset ? 1.0 : c.y);
            }
        }
        __syncwarp();
        tpp_obstacle_block<SPEC>(P, A, wbase, 1u << j, cur, 1 - cur, L.b, L.n_eff, t_acc, R_SSTEP, 0, olist);
        if (lane == j) {
            L.moved = 1;
            L.n_resto++;
            L.iter++;
            L.tmode = TM_EVAL;
            L.phase = PH_T;
        }
    }
    __syncwarp();
}

// Top of an interior-point iteration: convergence tests, barrier update; leaves the lane in phase B (Newton) or
// finishes the problem.  n = residual norms of the (new) current iterate.
__device__ __forceinline__ void tpp_iterate_top(const KParams &P, TppLane &L, const TppNorms &n) {
    const int N = P.N;
    L.theta = n.theta; L.f = n.f; L.slog = n.slog;
    if (L.theta0 < 0) L.theta0 = fmax(1.0, n.theta);
    const double df = L.df;
    const double sd = fmax(S_MAX, (n.sum_y + n.sum_z) / (double)(9 * N)) / S_MAX;
    const double sc = fmax(S_MAX, n.sum_z / (double)(4 * N)) / S_MAX;
    const double compl0 = fmax(fabs(n.pmin), fabs(n.pmax));
    const double E0 = fmax(n.dual_inf / sd, fmax(n.prim_inf, compl0 / sc));
    L.phase = PH_FIN;
    if (!isfinite(E0)) { L.status = B200MPC_INVALID_NUMBER_DETECTED; return; }
    if (E0 <= P.tol && n.dual_inf / df <= 1.0 && n.prim_inf <= 1e-4 && compl0 / df <= 1e-4) {
        L.status = B200MPC_SOLVE_SUCCEEDED;
        return;
    }
    if (P.acceptable_iter > 0 && E0 <= P.acceptable_tol && n.dual_inf / df <= 1e10 && n.prim_inf <= 1e-2 &&
        compl0 / df <= 1e-2) {
        L.acceptable_count++;
        if (L.acceptable_count >= P.acceptable_iter) { L.status = B200MPC_SOLVED_TO_ACCEPTABLE_LEVEL; return; }
    } else {
        L.acceptable_count = 0;
    }
    if (L.iter >= P.max_iter) { L.status = B200MPC_MAXITER_EXCEEDED; return; }
    // barrier parameter update (a tiny step in two consecutive iterations forces a decrease; when mu cannot decrease any
    // more: Search_Direction_Becomes_Too_Small)
    double mu = L.mu;
    bool changed = false;
    bool tflag = L.tiny_flag != 0;
    L.tiny_flag = 0;
    for (;;) {
        const double cm = fmax(fabs(n.pmax - mu), fabs(n.pmin - mu));
        const double Emu = fmax(n.dual_inf / sd, fmax(n.prim_inf, cm / sc));
        if (!(Emu <= K_EPS * mu) && !tflag) break;
        const double nm = fmax(fmin(K_MU * mu, tpp_pow(mu, TH_MU)), P.mu_floor);
        if (nm == mu) {
            if (tflag) { L.status = B200MPC_SEARCH_DIRECTION_TOO_SMALL; return; }
            break;
        }
        mu = nm;
        changed = true;
        tflag = false;
    }
    if (changed) {
        L.mu = mu;
        L.fmask = 0;
    }
    L.dw = 0.0;
    L.bmode = BM_NEWTON;
    L.phase = PH_B;
}

// With TPP_SYNC the warps of a CTA run the sweeps in lock-step (a __syncthreads() in front of every block), so the
// SM's instruction cache holds one loop body at a time instead of the whole kernel.
#ifndef TPP_THREADS
#define TPP_THREADS 384
#endif
#ifndef TPP_SYNC
#define TPP_SYNC 1
#endif
#ifndef TPP_EXP_NO_HANDOVER
#define TPP_EXP_NO_HANDOVER 0
#endif
// lock-step is a run-time choice only for the instances that can carry the obstacle cost
#define TPP_CTA_SYNC ((SPEC == TPP_SPEC_RK4_GOAL || SPEC == TPP_SPEC_EULER_TRAJ) ? true : (T.cta_sync != 0))
#define TPP_HAND (TPP_EXP_NO_HANDOVER ? false : (A.hand_rec != nullptr))
// cta_sync == 2 (instances with a run-time choice only): lock-step among the warps that share a scheduler (warp index mod 4:
// three of the twelve warps) — they share that scheduler's instruction buffer, and a barrier of three waits for less
// imbalance than a barrier of twelve.  Named barriers 1-4, 32 * (warps / 4) threads each.
static_assert((TPP_THREADS / 32) % 4 == 0, "the scheduler-group barriers count on the same number of warps per scheduler");
// (cta_sync == 3: two groups, even and odd warps — six warps on two schedulers each; measured against the groups of three)
__device__ __forceinline__ void tpp_group_sync(int wid, int mode) {
    const int gm = (mode == 3) ? 1 : 3;
    asm volatile("bar.sync %0, %1;" ::"r"(1 + (wid & gm)), "r"((TPP_THREADS / 32 / (gm + 1)) * 32) : "memory");
}
__device__ __forceinline__ bool tpp_group_and(int wid, int mode, bool pred) {
    int r;
    const int gm = (mode == 3) ? 1 : 3;
    asm volatile("{\n.reg .pred p, q;\nsetp.ne.s32 q, %3, 0;\nbar.red.and.pred p, %1, %2, q;\nselp.s32 %0, 1, 0, p;\n}"
                 : "=r"(r) : "r"(1 + (wid & gm)), "r"((TPP_THREADS / 32 / (gm + 1)) * 32), "r"((int)pred) : "memory");
    return r != 0;
}
// (compile-time off for the instances without the obstacle cost: measured there, the group barrier is slower than the CTA's,
// 205.9 vs 186 ms per 1 M problems, and the mere run-time choice costs those instances 112 B of spills and 13 %)
#ifndef TPP_STATIC_GROUP
#define TPP_STATIC_GROUP 0 /* build knob: 2 / 3 = the instances without the obstacle cost use the group barrier of that mode */
#endif
#define TPP_GROUP_SYNC ((SPEC == TPP_SPEC_RK4_GOAL || SPEC == TPP_SPEC_EULER_TRAJ) ? (TPP_STATIC_GROUP != 0) : (T.cta_sync >= 2))
#define TPP_GROUP_MODE ((SPEC == TPP_SPEC_RK4_GOAL || SPEC == TPP_SPEC_EULER_TRAJ) ? TPP_STATIC_GROUP : T.cta_sync)
#if TPP_SYNC == 1
#define TPP_BLOCK_SYNC()                          \
    do {                                          \
        if (TPP_GROUP_SYNC) tpp_group_sync(wid, TPP_GROUP_MODE);  \
        else if (TPP_CTA_SYNC) __syncthreads();   \
        else __syncwarp();                        \
    } while (0)
#else
#define TPP_BLOCK_SYNC() __syncwarp() /* TPP_SYNC == 2: the CTA only meets once per trip (at the exit test) */
#endif
#define TPP_SMEM_BYTES ((size_t)(TPP_THREADS / 32) * TPP_STAGE_SMEM + (size_t)TPP_THREADS * TPP_LANE_STRIDE * sizeof(double))

// ---- straggler hand-over (BatchArgs::hand_rec).  Out of line: inlined, the export code costs the sweeps registers
// (measured: +80 B of spills per thread and 8 % of the throughput of variant B, whether or not a problem is ever exported).
// Once per trip, whole warp: (1) the records of the lanes that decided to leave (phase PH_EXPORT: iterate rows of buffer
// myco, filter, solver scalars); a lane that finds the record buffer full stays and re-evaluates its point (one trip);
// (2) the threshold of the next trip: hand_iter iterations, 0 once the work queue is empty and the warp has thinned out.
__device__ __noinline__ void tpp_hand_trip(const KParams &P, const TppArgs &T, char *wbase, size_t gw, TppLane &L, int myco) {
    const BatchArgs &A = T.a;
    const int N = P.N, lane = threadIdx.x & 31;
    const int stage_b = T.stage_b;
    int ex = (L.phase == PH_EXPORT);
    if (ex) {
        const unsigned slot = atomicAdd(A.hand_count, 1u);
        if (slot < (unsigned)A.hand_cap) {
            L.hslot = (int)slot;
        } else {
            ex = 0;
            L.hslot = -2; // no further attempts
            L.tmode = TM_EVAL;
            L.phase = PH_T;
        }
    }
    for (unsigned m = __ballot_sync(FULL, ex); m; m &= m - 1) {
        const int j = __ffs(m) - 1;
        const int co = __shfl_sync(FULL, myco, j);
        double *hr = A.hand_rec + (size_t)__shfl_sync(FULL, L.hslot, j) * HAND_REC(N);
        for (int k = lane; k <= N; k += 32) {
            const char *pc = wbase + (size_t)k * stage_b + co * TPP_ROW_B + j * 16;
            double2 *o = reinterpret_cast<double2 *>(hr + 16 * k);
#pragma unroll
            for (int f = 0; f < R_ITER; f++) o[f] = tpp_ld2(pc, f);
        }
        const double *flj = T.filt + gw * (64 * 32) + j;
        hr[HAND_FILT(N) + lane] = flj[lane * 32];
        hr[HAND_FILT(N) + 32 + lane] = flj[(32 + lane) * 32];
    }
    if (ex) {
        double *q = A.hand_rec + (size_t)L.hslot * HAND_REC(N) + HAND_SCAL(N);
        q[HS_B] = (double)L.b; q[HS_MU] = L.mu; q[HS_DF] = L.df; q[HS_THETA0] = L.theta0; q[HS_DWLAST] = L.dw_last;
        q[HS_ITER] = (double)L.iter; q[HS_LS] = (double)L.ls_extra; q[HS_NRESTO] = (double)L.n_resto;
        q[HS_ACCEPT] = (double)L.acceptable_count; q[HS_TINYLAST] = (double)L.tiny_last;
        q[HS_TINYFLAG] = (double)L.tiny_flag; q[HS_FMASK] = (double)L.fmask; q[HS_RING] = (double)L.ring;
        L.phase = PH_LOAD;
    }
    if (A.hand_thin >= 0) { // (off by default: the work-queue counter is not even looked at then)
        const int ph = L.phase;
        const unsigned act = __ballot_sync(FULL, ph != PH_DONE && ph != PH_LOAD);
        const bool thin = __popc(act) <= A.hand_thin && *reinterpret_cast<volatile unsigned *>(A.counter) >= (unsigned)A.B;
        L.hcap = (L.hslot == -2) ? 0x7fffffff : (thin ? 0 : A.hand_iter);
    } else if (L.hslot == -2) {
        L.hcap = 0x7fffffff;
    }
}

template <int SPEC>
__global__ void __launch_bounds__(TPP_THREADS, B200MPC_TPP_MIN_CTAS) mpc_solve_tpp_kernel(const KParams P, const TppArgs T) {
    extern __shared__ __align__(16) char tpp_smem[];
    const BatchArgs &A = T.a;
    const int N = P.N;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t gw = (size_t)blockIdx.x * (blockDim.x >> 5) + wid;
    char *wbase = reinterpret_cast<char *>(T.ws) + gw * ((size_t)(N + 1) * TPP_STAGE_B); // the warp's workspace
    char *wb = wbase + lane * 16;                                                        // this lane's column of it
    double *fl = T.filt + gw * (64 * 32) + lane;
    char *sb = tpp_smem + (size_t)wid * TPP_STAGE_SMEM + lane * 16;
    TppLane &L = *reinterpret_cast<TppLane *>(tpp_smem + (size_t)(TPP_THREADS / 32) * TPP_STAGE_SMEM +
                                              (size_t)threadIdx.x * TPP_LANE_STRIDE * sizeof(double));
    L.phase = PH_LOAD;
    L.b = -1;
    L.moved = 0;
    // obstacle cost: the warp's list buffer (the list of the problem whose sums the warp is forming)
    double *olist = (TPP_HAS_OBS(P) && T.obs_smem) ? reinterpret_cast<double *>(tpp_smem + TPP_SMEM_BYTES) + (size_t)wid * 2 * ((P.M + 3) & ~3) : nullptr;
    int cur = 0; // warp-uniform: the buffer holding the current iterates during this trip

    for (;;) {
        // ---- block L: pull the next problems; the warp writes each starting point together (lane <-> stage) ----
        __syncwarp();
        int newb = -1;
        if (tpp_opaque(L.phase) == PH_LOAD) {
            const int b = (int)atomicAdd(A.counter, 1u);
            if (b >= A.B) {
                L.phase = PH_DONE;
            } else {
                newb = b;
                L.b = b;
                bool gone = false; // streamed call given up by the host
                if (T.avail) {
                    // streamed inputs: wait until the copy stream has delivered this problem (only ever at the start
                    // of a batch: the copies run 20x faster than the problems are consumed)
                    // (the abort word lives in host memory: it is looked at every 1024th round — tens of thousands of waiting
                    // lanes reading it over PCIe every round slowed the very input copies they wait for: 204 -> 217 ms)
                    unsigned spins = 0;
                    while (*reinterpret_cast<const volatile unsigned *>(T.avail) <= (unsigned)b) {
                        if ((++spins & 1023u) == 0 && T.abort && *reinterpret_cast<const volatile unsigned *>(T.abort)) { gone = true; break; }
                        __nanosleep(500);
                    }
                    __threadfence();
                }
                L.goal[0] = L.goal[1] = L.goal[2] = 0;
                if (TPP_IS_GOAL(P)) {
                    const double *xr = A.xref + 3 * (size_t)b;
                    L.goal[0] = __ldcg(xr); L.goal[1] = __ldcg(xr + 1); L.goal[2] = __ldcg(xr + 2);
                }
                L.status = B200MPC_MAXITER_EXCEEDED;
                L.iter = 0; L.ls_extra = 0; L.n_resto = 0; L.acceptable_count = 0; L.ntrial = 0; L.soc_count = 0;
                L.ring = 0; L.fmask = 0; L.keep = 0; L.soc_first = 1; L.moved = 0;
                L.tiny = 0; L.tiny_last = 0; L.tiny_flag = 0;
                L.df = 1.0; L.mu = P.mu_init;
                L.theta0 = -1; L.dw = 0; L.dw_last = 0;
                L.alpha = 0; L.a_z = 0; L.alpha_soc = 0; L.a_z_soc = 0; L.a_min = 0; L.theta_soc_old = 0;
                L.ref_phi = 0; L.ref_gbd = 0;
                L.f = 0; L.slog = 0; L.theta = 0;
                L.bmode = BM_LSQ;
                L.tmode = TM_LSQ;
                L.phase = PH_B;
                L.hslot = -1;
                L.hcap = (A.hand_iter_tail > 0 && (long long)b >= (long long)A.B - (long long)(gridDim.x * blockDim.x)) ? A.hand_iter_tail : A.hand_iter;
                if (gone) { L.phase = PH_DONE; newb = -1; }
            }
        }
        __syncwarp();
        for (unsigned m = __ballot_sync(FULL, newb >= 0); m; m &= m - 1) {
            const int j = __ffs(m) - 1;                      // slot (lane) that receives the problem
            const size_t b = (size_t)__shfl_sync(FULL, newb, j);
            // inputs bypass L1 (a line may straddle two chunks of a streamed batch)
            const double x00 = __ldcg(A.x0 + 3 * b), x01 = __ldcg(A.x0 + 3 * b + 1), x02 = __ldcg(A.x0 + 3 * b + 2);
            const double2 *ui = A.u_init ? reinterpret_cast<const double2 *>(A.u_init + b * 2 * N) : nullptr;
            for (int k = lane; k <= N; k += 32) {
                char *p = wbase + (size_t)k * TPP_STAGE_B + j * 16;
                char *pc = p + cur * R_ITER * TPP_ROW_B;
                tpp_st2(pc, R_X01, (k == 0) ? x00 : 0.0, (k == 0) ? x01 : 0.0);
                tpp_st2(pc, R_X2L0, (k == 0) ? x02 : 0.0, 0.0);
                tpp_st2(pc, R_L12, 0.0, 0.0);
                if (k < N) {
                    double2 u = make_double2(0.0, 0.0);
                    if (ui) u = __ldcg(ui + k);
                    const double uv[2] = {u.x, u.y};
                    double sv[2];
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        // slack initialisation: s = d(x) pushed into the interior; bound multipliers 1
                        const double lo = P.sL[i], hi = P.sU[i];
                        const double pl = fmin(BOUND_PUSH * fmax(1.0, fabs(lo)), BOUND_FRAC * (hi - lo));
                        const double pu = fmin(BOUND_PUSH * fmax(1.0, fabs(hi)), BOUND_FRAC * (hi - lo));
                        double sx = uv[i];
                        if (sx < lo + pl) sx = lo + pl;
                        if (sx > hi - pu) sx = hi - pu;
                        sv[i] = sx;
                    }
                    tpp_st2(pc, R_U, uv[0], uv[1]); tpp_st2(pc, R_S, sv[0], sv[1]); tpp_st2(pc, R_YD, 0.0, 0.0);
                    tpp_st2(pc, R_VL, 1.0, 1.0); tpp_st2(pc, R_VU, 1.0, 1.0);
                    if (!TPP_IS_GOAL(P)) {
                        const double *xr = A.xref + b * 3 * N + 3 * k;
                        const double *ur = A.uref + b * 2 * N + 2 * k;
                        tpp_st2(p, R_REF, __ldcg(xr), __ldcg(xr + 1)); tpp_st2(p, R_REF + 1, __ldcg(xr + 2), 0.0);
                        tpp_st2(p, R_REF + 2, __ldcg(ur), __ldcg(ur + 1));
                    }
                }
            }
        }
        __syncwarp();
        if (TPP_HAS_OBS(P)) {
            // new problems: length of the obstacle list without its padding, obstacle sums of the starting point
            const unsigned nm = __ballot_sync(FULL, newb >= 0);
            for (unsigned m = nm; m; m &= m - 1) {
                const int j = __ffs(m) - 1;
                const size_t b = (size_t)__shfl_sync(FULL, newb, j);
                const ObsList OL = obstacle_list_setup(A.ox + (size_t)A.obs_stride * b, A.oy + (size_t)A.obs_stride * b, P.M, lane);
                if (lane == j) L.n_eff = OL.n_eff;
            }
            __syncwarp();
            if (nm) tpp_obstacle_block<SPEC>(P, A, wbase, nm, cur, cur, L.b, L.n_eff, 0.0, -1, 0, olist);
        }
#if TPP_SYNC
        if (TPP_GROUP_SYNC ? tpp_group_and(wid, TPP_GROUP_MODE, L.phase == PH_DONE)
                           : (TPP_CTA_SYNC ? (bool)__syncthreads_and(L.phase == PH_DONE) : (bool)__all_sync(FULL, L.phase == PH_DONE))) break;
#else
        if (__all_sync(FULL, L.phase == PH_DONE)) break;
#endif

        // ---- block B ----  (b_passes > 1: a lane whose factorisation has the wrong inertia repeats the sweep with the next
        // delta_w in the same trip instead of costing itself a whole trip: with the obstacle cost half of the iterations need
        // an inertia correction, and block B is a small part of a trip there)
#pragma unroll 1
        for (int bpass = 0; bpass < T.b_passes; bpass++) {
        TPP_BLOCK_SYNC();
        if (tpp_opaque(L.phase) == PH_B) {
            TppBwd r;
            tpp_stat(T, 0);
            tpp_backward<SPEC>(P, A, wb, sb, cur, L, r);
            const int bmode = L.bmode;
            if (bmode == BM_LSQ) {
                // objective scaling from the gradient at the starting point; invalid-number check
                L.f = r.f;
                if (r.bad || !isfinite(r.f)) {
                    L.status = B200MPC_INVALID_NUMBER_DETECTED;
                    L.phase = PH_FIN;
                } else {
                    if (r.gmax > 100.0) L.df = fmax(100.0 / r.gmax, 1e-8);
                    if (r.ok) L.phase = PH_F;
                    else { L.keep = 0; L.tmode = TM_LSQ; L.phase = PH_T; }
                }
            } else if (r.ok) {
                if (bmode == BM_NEWTON && L.dw > 0.0) L.dw_last = L.dw;
                L.phase = PH_F;
            } else if (bmode == BM_SOC) {
                L.phase = PH_BACKTRACK; // correction abandoned: continue with the Newton direction
            } else {
                // inertia correction
                const double dw = L.dw, dwl = L.dw_last;
                double nd;
                if (dw == 0.0) nd = (dwl == 0.0) ? DW_INIT : fmax(DW_MIN, dwl * DW_DEC);
                else nd = (dwl == 0.0 || 1e5 * dwl < dw) ? dw * DW_INC_FIRST : dw * DW_INC;
                L.dw = nd;
                if (nd > DW_MAX) { L.status = B200MPC_ERROR_IN_STEP_COMPUTATION; L.phase = PH_FIN; }
            }
        }
        }

        // ---- block F ----
        TPP_BLOCK_SYNC();
        if (tpp_opaque(L.phase) == PH_F) {
            TppFwd f;
            tpp_stat(T, 1);
            tpp_forward<SPEC>(P, A, wb, sb, cur, L, f);
            const int bmode = L.bmode;
            if (bmode == BM_LSQ) {
                // the estimate is kept if its largest entry is <= 1e3; sweep T adds the defect multipliers
                L.ymax_f = f.ymax;
                L.keep = (L.df * f.ymax <= 1e3) ? 1 : 0;
                L.tmode = TM_LSQ;
                L.phase = PH_T;
            } else if (bmode == BM_NEWTON) {
                if (f.bad) {
                    L.status = B200MPC_ERROR_IN_STEP_COMPUTATION;
                    L.phase = PH_FIN;
                } else {
                    const double theta = L.theta;
                    L.ref_phi = L.df * L.f - L.mu * L.slog;
                    L.ref_gbd = f.gbd;
                    double a_min = GAMMA_THETA;
                    if (f.gbd < 0) {
                        a_min = fmin(GAMMA_THETA, GAMMA_PHI * theta / (-f.gbd));
                        if (theta <= 1e-4 * L.theta0) a_min = fmin(a_min, DELTA_LS * tpp_pow(theta, S_THETA) / tpp_pow(-f.gbd, S_PHI));
                    }
                    L.a_min = a_min * ALPHA_MIN_FRAC;
                    L.alpha = f.a_max;
                    L.a_z = f.a_z;
                    L.ntrial = 0;
                    L.tiny = f.tiny;
                    if (!f.tiny) { L.tiny_flag = 0; L.tiny_last = 0; }
                    L.tmode = TM_STEP;
                    L.phase = PH_T;
                }
            } else {
                L.alpha_soc = f.a_max;
                L.a_z_soc = f.a_z;
                L.tmode = TM_STEP_SOC;
                L.phase = PH_T;
            }
        }

        // ---- block O: obstacle sums of the points sweep T is about to evaluate (into the cache rows of the other buffer) ----
        if (TPP_HAS_OBS(P)) {
            TPP_BLOCK_SYNC(); // (without this meeting: 736.5 vs 740.7 ms per 262 144 problems — immaterial)
            const int tm = L.tmode;
            const unsigned om = __ballot_sync(FULL, tpp_opaque(L.phase) == PH_T);
            if (om) {
                const int srow = (tm == TM_STEP) ? R_STEP : ((tm == TM_STEP_SOC) ? R_SSTEP : -1);
                const double al = (tm == TM_STEP_SOC) ? L.alpha_soc : L.alpha;
                // multiplier estimate / evaluation after a restoration: the point is the current one, whose sums are cached
                const int cp = (tm == TM_LSQ || tm == TM_EVAL) ? 1 : 0;
                tpp_obstacle_block<SPEC>(P, A, wbase, om, cur, 1 - cur, L.b, L.n_eff, al, srow, cp, olist);
            }
        }

        // ---- block T ----
        TPP_BLOCK_SYNC();
        if (tpp_opaque(L.phase) == PH_T) {
            TppTrial t;
            tpp_stat(T, 2);
            tpp_trial<SPEC>(P, A, wb, sb, cur, L, t);
            const int tm = L.tmode;
            bool accepted = false;
            if (tm == TM_EVAL) {
                accepted = true;
            } else if (tm == TM_LSQ) {
                // discard the estimate (and write zeros in the next trip) if a defect multiplier exceeds the limit
                if (L.keep && !(L.df * t.ymax <= 1e3)) L.keep = 0;
                else accepted = true;
            } else if (t.bad) {
                L.status = B200MPC_ERROR_IN_STEP_COMPUTATION;
                L.phase = PH_FIN;
            } else {
                const bool soc = (tm == TM_STEP_SOC);
                if (soc || L.ntrial++ > 0) L.ls_extra++;
                bool fa;
                bool ok = tpp_ls_acceptable(L, fl, L.alpha, t.phi, t.th, fa);
                if (!soc && L.tiny) {
                    // tiny step: the full step is accepted without a line search (unless it cannot be evaluated)
                    L.tiny = 0;
                    if (isfinite(t.th) && isfinite(t.phi)) {
                        ok = true;
                        if (L.tiny_last) L.tiny_flag = 1;
                        L.tiny_last = (t.dymax < TINY_STEP_Y_TOL) ? 1 : 0;
                    } else {
                        L.tiny_flag = 0; L.tiny_last = 0;
                    }
                }
                if (ok) {
                    if (!fa) tpp_filter_add(fl, L, L.ref_phi - GAMMA_PHI * L.theta, (1 - GAMMA_THETA) * L.theta);
                    L.iter++;
                    accepted = true;
                } else if (!soc) {
                    if (L.ntrial == 1 && P.max_soc > 0 && isfinite(t.th) && t.th >= L.theta) {
                        // second-order correction, first round: right-hand sides from this trial point
                        L.soc_count = 0;
                        L.soc_first = 1;
                        L.theta_soc_old = t.th;
                        L.bmode = BM_SOC;
                        L.phase = PH_B;
                    } else {
                        L.phase = PH_BACKTRACK;
                    }
                } else {
                    L.soc_count++;
                    if (L.soc_count < P.max_soc && t.th <= KAPPA_SOC * L.theta_soc_old) {
                        L.theta_soc_old = t.th;
                        L.soc_first = 0;
                        L.bmode = BM_SOC;
                        L.phase = PH_B;
                    } else {
                        L.phase = PH_BACKTRACK;
                    }
                }
            }
            if (accepted) {
                L.moved = 1;
                // a straggler leaves for the warp kernel here, at the top of an iteration, with its state as it stands
                bool leave = false;
                if (TPP_HAND && L.iter >= L.hcap) { L.phase = PH_EXPORT; leave = true; }
                // top of the next iteration: convergence tests, barrier update
                if (!leave) tpp_iterate_top(P, L, t.n);
            }
        }

        // ---- rare: backtracking / restoration stand-in ----
        __syncwarp();
        {
            int resto = 0;
            if (tpp_opaque(L.phase) == PH_BACKTRACK) {
                L.alpha *= 0.5;
                if (L.alpha < L.a_min) {
                    resto = 1;
                } else {
                    L.tmode = TM_STEP;
                    L.phase = PH_T;
                }
            }
            if (TPP_HAS_OBS(P)) {
                const unsigned rm = __ballot_sync(FULL, resto);
                if (rm) tpp_restore_obs<SPEC>(P, A, wbase, fl, cur, L, rm, olist);
            } else if (resto) {
                tpp_restore<SPEC>(P, wb, fl, cur, L);
            }
        }

        // ---- result store and release of finished lanes; iterate copy for lanes that did not move ----
        // Both are done by the whole warp for one lane at a time (lane <-> stage): one round trip per event instead
        // of one per stage, and stage-contiguous result stores.
        __syncwarp();
        {
            const int ph = tpp_opaque(L.phase);
            const bool fin = (ph == PH_FIN);
            const bool cpy = (ph == PH_B || ph == PH_F || ph == PH_T) && !L.moved;
            const int myco = (L.moved ? (1 - cur) : cur) * R_ITER;
            for (unsigned m = __ballot_sync(FULL, fin); m; m &= m - 1) {
                const int j = __ffs(m) - 1;
                const size_t b = (size_t)__shfl_sync(FULL, L.b, j);
                const int co = __shfl_sync(FULL, myco, j);
                double *xo = A.X + b * 3 * (N + 1);
                double2 *uo = reinterpret_cast<double2 *>(A.U + b * 2 * N);
                for (int k = lane; k <= N; k += 32) {
                    const char *pc = wbase + (size_t)k * TPP_STAGE_B + co * TPP_ROW_B + j * 16;
                    const double2 a = tpp_ld2(pc, R_X01), c = tpp_ld2(pc, R_X2L0);
                    double2 u = make_double2(0.0, 0.0);
                    if (k < N) u = tpp_ld2(pc, R_U);
                    xo[3 * k] = a.x; xo[3 * k + 1] = a.y; xo[3 * k + 2] = c.x;
                    if (k < N) uo[k] = u;
                }
            }
            // hand-over records: iterate rows of the new current buffer, filter, solver scalars
            if (TPP_HAND && (A.hand_thin >= 0 || __any_sync(FULL, ph == PH_EXPORT))) tpp_hand_trip(P, T, wbase, gw, L, myco);
            if (fin) {
                const size_t b = (size_t)L.b;
                if (A.cost) A.cost[b] = L.f;
                A.status[b] = L.status;
                if (A.iters) A.iters[b] = L.iter;
                if (A.ls) A.ls[b] = L.ls_extra;
                L.phase = PH_LOAD;
            }
            if (T.done && __any_sync(FULL, fin)) {
                // streamed results: the lane that completes a chunk raises the chunk's flag in host memory; the host
                // then copies the chunk out while the kernel keeps running
                __threadfence(); // every lane: its share of the cooperative result stores
                __syncwarp();
                if (fin) {
                    const int c = L.b / T.chunk;
                    const unsigned n_in = (unsigned)min(T.chunk, A.B - c * T.chunk);
                    if (atomicAdd(T.done + c, 1u) + 1u == n_in) {
                        __threadfence_system();
                        *reinterpret_cast<volatile unsigned *>(T.flags + c) = 1u;
                        __threadfence_system();
                    }
                }
            }
            // the iterate stays where it is but the warp's buffers swap: copy it across (rejected trial point,
            // inertia-correction retry)
            for (unsigned m = __ballot_sync(FULL, cpy); m; m &= m - 1) {
                const int j = __ffs(m) - 1;
                const int co = cur * R_ITER, no = R_ITER - co;
                for (int k = lane; k <= N; k += 32) {
                    const char *pc = wbase + (size_t)k * TPP_STAGE_B + co * TPP_ROW_B + j * 16;
                    char *pw = wbase + (size_t)k * TPP_STAGE_B + no * TPP_ROW_B + j * 16;
                    double2 v[R_ITER];
#pragma unroll
                    for (int f = 0; f < R_ITER; f++) v[f] = tpp_ld2(pc, f);
#pragma unroll
                    for (int f = 0; f < R_ITER; f++) tpp_st2(pw, f, v[f].x, v[f].y);
                    if (TPP_HAS_OBS(P)) {
                        char *ps = wbase + (size_t)k * TPP_STAGE_B + j * 16;
#pragma unroll
                        for (int f = 0; f < 3; f++) {
                            const double2 c = tpp_ld2(ps, R_OC + 3 * cur + f);
                            tpp_st2(ps, R_OC + 3 * (1 - cur) + f, c.x, c.y);
                        }
                    }
                }
            }
            L.moved = 0;
        }
        cur ^= 1;
        tpp_stat(T, 3);
    }
}
