// tpp_fused.cuh — lane-per-problem kernel with TWO streaming sweeps per interior-point iteration.
//
// tpp_kernel.cuh runs three sweeps per iteration: B (backward: linearise, Riccati), F (forward: roll the step out) and
// T (backward: apply the step, evaluate the new point).  T of iteration i and B of iteration i+1 walk the horizon in the
// same direction over the same point — the trial point T has just written is, 99 % of the time, the iterate B reads
// back a moment later.  Here they are one sweep, TB: at every stage the new iterate is formed, written and evaluated,
// and — from the values still in registers — linearised and pushed through the Riccati step of the NEXT iteration.
// Per stage and iteration this removes the 8 iterate rows B used to read (44 -> 37 workspace rows), one of the four
// sin/cos pairs and one of the three lock-step barriers.
//
// What makes the fusion legal:
//  * the Riccati matrices do not depend on the barrier parameter; the right-hand side does, affinely.  The new mu is
//    only known after the sweep (it follows from the norms of the new point), so the vector recursion is carried twice,
//    (v, kf) = (v_a, kf_a) + mu (v_b, kf_b), and sweep F forms kf = kf_a + mu kf_b (one more gain row, R_KB);
//  * the first factorisation of an iteration always uses delta_w = 0;
//  * anything else that is only known after the sweep — trial point rejected, problem converged, wrong inertia —
//    makes the B-part's output unused (1 % / once per problem) or costs the lane a repeat in HOLD mode (the T-part
//    copies the iterate unchanged; this is also how second-order-correction right-hand sides are prepared).
// Trip = blocks L(oad), TB, F.  Everything else (workspace layout, staging, lane state, filter, restoration stand-in,
// result stores) is shared with tpp_kernel.cuh; the algorithm and its decisions are unchanged.
#pragma once
#ifndef TPPF_L2_PREFETCH
#define TPPF_L2_PREFETCH 0 /* the two-sweep kernel is instruction-fetch bound: the L2 prefetches cost more than they hide */
#endif
__device__ __forceinline__ void tppf_l2_prefetch(const char *p, int row0, int n) {
#if TPPF_L2_PREFETCH
    tpp_l2_prefetch(p, row0, n);
#endif
}

// workspace row R_KB = (kfb0, kfb1): barrier-parameter coefficient of the feed-forward gain (tpp_kernel.cuh)
enum { PHF_LOAD = 0, PHF_TB = 1, PHF_F = 2, PHF_BACKTRACK = 4, PHF_FIN = 6, PHF_DONE = 7 };
enum { TMF_INIT = 0, TMF_EVAL = 1, TMF_LSQ = 2, TMF_STEP = 3, TMF_STEP_SOC = 4, TMF_HOLD = 5 };

// ---- sweep F (as tpp_forward, gains kf = kf_a + mu kf_b, iterate read from buffer `buf`) ----------------------------
template <int SPEC>
__device__ __forceinline__ void tppf_forward_stage(char *sb, const char *p, int co, bool has_next) {
#pragma unroll
    for (int i = 0; i < 4; i++) tpp_cp16(sb + i * TPP_ROW_B, p + (R_K + i) * TPP_ROW_B);
    tpp_cp16(sb + 4 * TPP_ROW_B, p + R_KB * TPP_ROW_B);
    const char *pc = p + co * TPP_ROW_B;
    tpp_cp16(sb + 5 * TPP_ROW_B, pc + R_U * TPP_ROW_B);
    tpp_cp16(sb + 6 * TPP_ROW_B, pc + R_S * TPP_ROW_B);
    tpp_cp16(sb + 7 * TPP_ROW_B, pc + R_VL * TPP_ROW_B);
    tpp_cp16(sb + 8 * TPP_ROW_B, pc + R_VU * TPP_ROW_B);
    if (has_next) {
        tpp_cp16(sb + 9 * TPP_ROW_B, pc + TPP_STAGE_B + R_X01 * TPP_ROW_B);
        tpp_cp16(sb + 10 * TPP_ROW_B, pc + TPP_STAGE_B + R_X2L0 * TPP_ROW_B);
    }
    tpp_cp_commit();
}

template <int SPEC>
__device__ __forceinline__ void tppf_forward(const KParams &P, char *wb, char *sb, int buf, const TppLane &L, TppFwd &o) {
    const int N = P.N;
    const double dt = P.dt, mu = L.mu, tau = fmax(TAU_MIN, 1.0 - L.mu), df = L.df;
    const double *goal = L.goal;
    const int bmode0 = L.bmode;
    const int co = buf * R_ITER;
    const int orow = (bmode0 == BM_SOC) ? R_SSTEP : R_STEP;
    double y0 = 0, y1 = 0, y2 = 0;
    double a_max = 1.0, a_z = 1.0, gbd = 0, ymax = 0;
    int bad = 0;
    double X[3];
    {
        const double2 a = tpp_ld2(wb + co * TPP_ROW_B, R_X01), b = tpp_ld2(wb + co * TPP_ROW_B, R_X2L0);
        X[0] = a.x; X[1] = a.y; X[2] = b.x;
    }
    tppf_forward_stage<SPEC>(sb, wb, co, N > 0);
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
        const int mode = tpp_opaque(bmode0);
        char *p = wb + (size_t)k * TPP_STAGE_B;
        tpp_cp_wait();
        const double2 k0_ = tpp_sld(sb, 0), k1_ = tpp_sld(sb, 1), k2_ = tpp_sld(sb, 2), kfa_ = tpp_sld(sb, 3), kfb_ = tpp_sld(sb, 4);
        const double2 u2 = tpp_sld(sb, 5), s2 = tpp_sld(sb, 6), vl2 = tpp_sld(sb, 7), vu2 = tpp_sld(sb, 8);
        const double2 xn01 = tpp_sld(sb, 9), xn2 = tpp_sld(sb, 10);
        tpp_consume(k0_, k1_, k2_, kfa_);
        tpp_consume(kfb_, u2, s2, vl2);
        tpp_consume(vu2, xn01, xn2, xn2);
        if (k < N) tppf_forward_stage<SPEC>(sb, p + TPP_STAGE_B, co, k + 1 < N);
        if (k + 2 <= N) {
            tppf_l2_prefetch(p + 2 * TPP_STAGE_B, R_K, 4);
            tppf_l2_prefetch(p + 2 * TPP_STAGE_B, R_KB, 1);
            tppf_l2_prefetch(p + 2 * TPP_STAGE_B + co * TPP_ROW_B, R_U, 2);
            tppf_l2_prefetch(p + 2 * TPP_STAGE_B + co * TPP_ROW_B, R_VL, 2);
            if (k + 3 <= N) tppf_l2_prefetch(p + 3 * TPP_STAGE_B + co * TPP_ROW_B, R_X01, 2);
        }
        if (!isfinite(y0) || !isfinite(y1) || !isfinite(y2)) bad = 1;
        tpp_st2(p, orow, y0, y1); tpp_st2(p, orow + 1, y2, 0.0);
        if (k < N) {
            const double kf0 = kfa_.x + mu * kfb_.x, kf1 = kfa_.y + mu * kfb_.y;
            const double du0 = kf0 + k0_.x * y0 + k0_.y * y1 + k1_.x * y2;
            const double du1 = kf1 + k1_.y * y0 + k2_.x * y1 + k2_.y * y2;
            tpp_st2(p, orow + 2, du0, du1);
            if (!isfinite(du0) || !isfinite(du1)) bad = 1;
            const double du[2] = {du0, du1};
            const double U[2] = {u2.x, u2.y}, Sv[2] = {s2.x, s2.y}, vLv[2] = {vl2.x, vl2.y}, vUv[2] = {vu2.x, vu2.y};
            const double Xn[3] = {xn01.x, xn01.y, xn2.x};
            double r[3], ub[2];
            tpp_ref<SPEC>(P, goal, p, r, ub);
            const double ln0[3] = {0, 0, 0};
            TppLin q;
            tpp_lin<false, SPEC>(P, r, ub, X, U, ln0, df, q, tpp_obs_zero());
            double rc0, rc1, rc2, rdv[2] = {U[0] - Sv[0], U[1] - Sv[1]};
            if (mode == BM_LSQ) {
                rc0 = rc1 = rc2 = 0;
                ymax = fmax(ymax, fmax(fabs(du0), fabs(du1)));
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
                }
            }
            const double n0 = y0 + q.a13 * y2 + q.b11 * du0 + q.b12 * du1 - rc0;
            const double n1 = y1 + q.a23 * y2 + q.b21 * du0 + q.b22 * du1 - rc1;
            const double n2 = y2 + dt * du1 - rc2;
            y0 = n0; y1 = n1; y2 = n2;
            X[0] = Xn[0]; X[1] = Xn[1]; X[2] = Xn[2];
        }
    }
    o.a_max = a_max; o.a_z = a_z; o.gbd = gbd; o.ymax = ymax; o.bad = bad;
}

// ---- sweep TB ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tppf_stage(char *sb, const char *p, int co, int srow) {
#pragma unroll
    for (int i = 0; i < R_ITER; i++) tpp_cp16(sb + i * TPP_ROW_B, p + (co + i) * TPP_ROW_B);
#pragma unroll
    for (int i = 0; i < 3; i++) tpp_cp16(sb + (R_ITER + i) * TPP_ROW_B, p + (srow + i) * TPP_ROW_B);
    tpp_cp_commit();
}

// T-part (tmode): INIT / EVAL / HOLD  new = current;  LSQ  new = current with the multiplier estimate;
//                 STEP(_SOC)  new = current + alpha*step, multiplier step from the costate recursion.
// B-part (bmode): NEWTON / SOC / LSQ as tpp_backward, on the new point, right-hand side split in (1, mu) parts.
template <int SPEC>
__device__ __forceinline__ void tppf_trial_backward(const KParams &P, char *wb, char *sb, int cur, const TppLane &L, double dwb,
                                                    TppTrial &o, TppBwd &ob) {
    const int N = P.N;
    const double dt = P.dt, mu = L.mu, df = L.df, dw = L.dw;
    const double *goal = L.goal;
    const int tmode0 = L.tmode, keep0 = L.keep, bmode0 = L.bmode;
    const bool soc0 = (tmode0 == TMF_STEP_SOC);
    const double alpha = soc0 ? L.alpha_soc : L.alpha, a_z = soc0 ? L.a_z_soc : L.a_z;
    const int co = cur * R_ITER, no = R_ITER - co;
    const int srow = soc0 ? R_SSTEP : R_STEP;
    const double ikap = 1.0 / KAPPA_SIGMA;
    // B-part, second-order correction: trial point of the rejected Newton / previous correction step
    const int sfirst = L.soc_first;
    const int brow = sfirst ? R_STEP : R_SSTEP;
    const double at = sfirst ? L.alpha : L.alpha_soc;
    double Xn[3] = {0, 0, 0}, ln[3] = {0, 0, 0};
    double lo[3] = {0, 0, 0}, Lam[3] = {0, 0, 0};
    double th = 0, slog = 0, pi = 0, di = 0, sy = 0, sz = 0, pmin = 1e300, pmax = -1e300, fs = 0, ymax = 0;
    double chk = 0.0, bchk = 0.0;
    // Riccati state: cost-to-go matrix, its vector in (1, mu) parts
    double q00 = 0, q01 = 0, q02 = 0, q11 = 0, q12 = 0, q22 = 0, v0 = 0, v1 = 0, v2 = 0, z0 = 0, z1 = 0, z2 = 0;
    double gmax = 0;
    int ok = 1;
    tppf_stage(sb, wb + (size_t)N * TPP_STAGE_B, co, srow);
#pragma unroll 1
    for (int k = N; k >= 0; --k) {
        const int mode = tpp_opaque(tmode0), bmode = tpp_opaque(bmode0);
        const bool step = (mode == TMF_STEP || mode == TMF_STEP_SOC), soc = (mode == TMF_STEP_SOC);
        const bool lsq = (mode == TMF_LSQ) && keep0;
        const bool useW = (bmode != BM_LSQ);
        const double dwk2 = useW ? dwb : 1.0;
        char *p = wb + (size_t)k * TPP_STAGE_B;
        char *pw = p + no * TPP_ROW_B;
        tpp_cp_wait();
        const double2 x01 = tpp_sld(sb, R_X01), x2l0 = tpp_sld(sb, R_X2L0), l12 = tpp_sld(sb, R_L12), u2 = tpp_sld(sb, R_U);
        const double2 s2 = tpp_sld(sb, R_S), yd2 = tpp_sld(sb, R_YD), vl2 = tpp_sld(sb, R_VL), vu2 = tpp_sld(sb, R_VU);
        const double2 dx01 = tpp_sld(sb, R_ITER), dx2 = tpp_sld(sb, R_ITER + 1), du2 = tpp_sld(sb, R_ITER + 2);
        tpp_consume(x01, x2l0, l12, u2);
        tpp_consume(s2, yd2, vl2, vu2);
        tpp_consume(dx01, dx2, du2, du2);
        if (k > 0) tppf_stage(sb, p - TPP_STAGE_B, co, srow);
        if (k > 1) {
            tppf_l2_prefetch(p - 2 * TPP_STAGE_B + co * TPP_ROW_B, 0, R_ITER);
            tppf_l2_prefetch(p - 2 * TPP_STAGE_B, srow, 3);
        }
        double r[3], ub[2];
        tpp_ref<SPEC>(P, goal, p, r, ub);
        // ================= T-part: the new iterate of stage k =================
        const double Xc[3] = {x01.x, x01.y, x2l0.x}; // current iterate (the second-order correction needs it below)
        double X[3] = {Xc[0], Xc[1], Xc[2]};
        double lam[3] = {0, 0, 0};
        const double duv[2] = {du2.x, du2.y};
        if (k >= 1) {
            lam[0] = x2l0.y; lam[1] = l12.x; lam[2] = l12.y;
            const double lcur[3] = {lam[0], lam[1], lam[2]};
            const double dX[3] = {dx01.x, dx01.y, dx2.x};
            if (step || lsq) {
                const double dwk = lsq ? 1.0 : dw;
                double Lk[3];
                if (k == N) {
                    Lk[0] = -dwk * dX[0]; Lk[1] = -dwk * dX[1]; Lk[2] = -dwk * dX[2];
                } else {
                    const double Uc[2] = {u2.x, u2.y};
                    tpp_costate<SPEC>(P, r, X, Uc, lo, Lam, dX, duv, lsq ? 1.0 : df, dwk, !lsq, Lk, tpp_obs_zero());
                }
                Lam[0] = Lk[0]; Lam[1] = Lk[1]; Lam[2] = Lk[2];
                chk = fma(0.0, (Lk[0] + Lk[1]) + Lk[2], chk); // 0*inf = NaN and NaN sticks: one test after the sweep
                if (step) {
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        X[i] += alpha * dX[i];
                        lam[i] += alpha * (Lk[i] - lam[i]);
                    }
                } else {
                    ymax = fmax(ymax, fmax(fabs(Lk[0]), fmax(fabs(Lk[1]), fabs(Lk[2]))));
#pragma unroll
                    for (int i = 0; i < 3; i++) lam[i] = df * Lk[i];
                }
            } else if (mode == TMF_LSQ) {
                lam[0] = lam[1] = lam[2] = 0.0;
            }
            lo[0] = lcur[0]; lo[1] = lcur[1]; lo[2] = lcur[2];
        }
        tpp_st2(pw, R_X01, X[0], X[1]); tpp_st2(pw, R_X2L0, X[2], lam[0]); tpp_st2(pw, R_L12, lam[1], lam[2]);
        if (k == N) {
            if (k >= 1) {
                // terminal state: no cost; its stationarity residual is the multiplier itself
                di = fmax(di, fmax(fabs(lam[0]), fmax(fabs(lam[1]), fabs(lam[2]))));
                sy += fabs(lam[0]) + fabs(lam[1]) + fabs(lam[2]);
            }
            // B-part, terminal stage: no cost, no controls
            q00 = dwk2; q01 = 0; q02 = 0; q11 = dwk2; q12 = 0; q22 = dwk2;
            z0 = z1 = z2 = 0;
            if (bmode == BM_LSQ) { v0 = 0; v1 = 0; v2 = 0; }
            else { v0 = lam[0]; v1 = lam[1]; v2 = lam[2]; }
        } else {
            double U[2] = {u2.x, u2.y}, S[2] = {s2.x, s2.y}, yd[2] = {yd2.x, yd2.y}, vL[2] = {vl2.x, vl2.y}, vU[2] = {vu2.x, vu2.y};
            const double Uc[2] = {U[0], U[1]}, Sc[2] = {S[0], S[1]};
            double rdv[2] = {U[0] - S[0], U[1] - S[1]};
            if (soc) {
                const double2 d2 = tpp_ld2(p, R_CS + 2);
                rdv[0] = d2.x; rdv[1] = d2.y;
            }
            double isl2[2], isu2[2]; // reciprocal slack distances of the new point (T-part clamps, B-part Sigma)
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
                    yd[i] += alpha * (Dsig * ds + rs);
                    vL[i] += a_z * dvL;
                    vU[i] += a_z * dvU;
                    isl2[i] = tpp_rcp(S[i] - P.sL[i]); isu2[i] = tpp_rcp(P.sU[i] - S[i]);
                    const double ml = mu * isl2[i], mu_u = mu * isu2[i];
                    vL[i] = fmax(fmin(vL[i], KAPPA_SIGMA * ml), ml * ikap);
                    vU[i] = fmax(fmin(vU[i], KAPPA_SIGMA * mu_u), mu_u * ikap);
                } else {
                    isl2[i] = tpp_rcp(S[i] - P.sL[i]); isu2[i] = tpp_rcp(P.sU[i] - S[i]);
                    if (mode == TMF_LSQ) yd[i] = keep0 ? df * duv[i] : 0.0;
                }
            }
            tpp_st2(pw, R_U, U[0], U[1]); tpp_st2(pw, R_S, S[0], S[1]); tpp_st2(pw, R_YD, yd[0], yd[1]);
            tpp_st2(pw, R_VL, vL[0], vL[1]); tpp_st2(pw, R_VU, vU[0], vU[1]);
            TppLin q;
            tpp_lin<true, SPEC>(P, r, ub, X, U, ln, df, q, tpp_obs_zero());
            const double c[3] = {Xn[0] - q.F0, Xn[1] - q.F1, Xn[2] - q.F2};
            fs += q.f;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                th += fabs(c[i]);
                pi = fmax(pi, fabs(c[i]));
            }
            double rx0 = 0, rx1 = 0, rx2 = 0;
            if (k >= 1) {
                rx0 = q.g[0] + lam[0] - ln[0];
                rx1 = q.g[1] + lam[1] - ln[1];
                rx2 = q.g[2] + lam[2] - (q.a13 * ln[0] + q.a23 * ln[1] + ln[2]);
                di = fmax(di, fmax(fabs(rx0), fmax(fabs(rx1), fabs(rx2))));
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
            slog += log(inside ? prod : -1.0);
            // ================= B-part: Riccati step of the next factorisation at the new point =================
            double ra0, ra1, ra2, qa[2], qb[2], Dsig[2], rc[3];
            if (bmode == BM_LSQ) {
                ra0 = q.g[0]; ra1 = q.g[1]; ra2 = q.g[2];
                qa[0] = q.g[3]; qa[1] = q.g[4]; qb[0] = qb[1] = 0.0;
                Dsig[0] = Dsig[1] = 1.0;
                rc[0] = rc[1] = rc[2] = 0.0;
                bchk = fma(0.0, ((c[0] + c[1]) + c[2]) + (q.g[3] + q.g[4]), bchk);
                gmax = fmax(gmax, fmax(fabs(q.g[3]), fabs(q.g[4])));
                if (k >= 1) {
                    bchk = fma(0.0, (q.g[0] + q.g[1]) + q.g[2], bchk);
                    gmax = fmax(gmax, fmax(fabs(q.g[0]), fmax(fabs(q.g[1]), fabs(q.g[2]))));
                }
            } else {
                ra0 = rx0; ra1 = rx1; ra2 = rx2;
                double rd[2] = {U[0] - S[0], U[1] - S[1]};
                rc[0] = c[0]; rc[1] = c[1]; rc[2] = c[2];
#pragma unroll
                for (int i = 0; i < 2; i++) Dsig[i] = vL[i] * isl2[i] + vU[i] * isu2[i] + dwb;
                if (bmode == BM_SOC) {
                    // (HOLD mode: X, U, S are the current iterate) defects of the last trial point curr + at*step
                    const double2 d01 = tpp_ld2(p, brow), d2 = tpp_ld2(p, brow + 1), dub = tpp_ld2(p, brow + 2);
                    double2 dsp = make_double2(rd[0], rd[1]), cs01 = make_double2(rc[0], rc[1]), cs2 = make_double2(rc[2], 0.0);
                    if (!sfirst) { dsp = tpp_ld2(p, R_CS + 2); cs01 = tpp_ld2(p, R_CS); cs2 = tpp_ld2(p, R_CS + 1); }
                    const double dub_[2] = {dub.x, dub.y}, rdp[2] = {dsp.x, dsp.y}, base[3] = {cs01.x, cs01.y, cs2.x};
                    const double2 n01 = tpp_ld2(p + TPP_STAGE_B, brow), n2 = tpp_ld2(p + TPP_STAGE_B, brow + 1);
                    const double Xtn[3] = {Xn[0] + at * n01.x, Xn[1] + at * n01.y, Xn[2] + at * n2.x};
                    double Xt[3], Ut[2], St[2], Ft[3];
                    Xt[0] = Xc[0] + at * d01.x; Xt[1] = Xc[1] + at * d01.y; Xt[2] = Xc[2] + at * d2.x;
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        Ut[i] = Uc[i] + at * dub_[i];
                        St[i] = Sc[i] + at * (dub_[i] + rdp[i]);
                        rd[i] = at * rdp[i] + (Ut[i] - St[i]);
                    }
                    tpp_dyn<SPEC>(P, Xt, Ut, Ft);
#pragma unroll
                    for (int i = 0; i < 3; i++) rc[i] = at * base[i] + (Xtn[i] - Ft[i]);
                    tpp_st2(p, R_CS, rc[0], rc[1]); tpp_st2(p, R_CS + 1, rc[2], 0.0); tpp_st2(p, R_CS + 2, rd[0], rd[1]);
                }
                // qu = ru + Dsig*rd + rs,  rs = -yd + mu (1/su - 1/sl)
                qa[0] = ru0 + Dsig[0] * rd[0] - yd[0]; qa[1] = ru1 + Dsig[1] * rd[1] - yd[1];
                qb[0] = isu2[0] - isl2[0]; qb[1] = isu2[1] - isl2[1];
            }
            const double hxx = useW ? q.hxx : 0.0, hyy = useW ? q.hyy : 0.0;
            const double htt = useW ? q.htt : 0.0, htv = useW ? q.htv : 0.0, htw = useW ? q.htw : 0.0;
            const double hvv = useW ? q.hvv : 0.0, hvw = useW ? q.hvw : 0.0, hww = useW ? q.hww : 0.0;
            const double a = q.a13, b = q.a23, b11 = q.b11, b12 = q.b12, b21 = q.b21, b22 = q.b22;
            const double d0 = -rc[0], d1 = -rc[1], d2 = -rc[2];
            const double w0 = q00 * d0 + q01 * d1 + q02 * d2 + v0;
            const double w1 = q01 * d0 + q11 * d1 + q12 * d2 + v1;
            const double w2 = q02 * d0 + q12 * d1 + q22 * d2 + v2;
            const double t0 = q02 + a * q00 + b * q01;
            const double t1 = q12 + a * q01 + b * q11;
            const double t2 = q22 + a * q02 + b * q12;
            const double x00 = hxx + dwk2 + q00, x01_ = q01, x11 = hyy + dwk2 + q11;
            const double x02 = t0, x12 = t1, x22 = htt + dwk2 + t2 + a * t0 + b * t1;
            const double e0 = b11 * q00 + b21 * q01, e1 = b11 * q01 + b21 * q11;
            const double f0 = b12 * q00 + b22 * q01 + dt * q02, f1 = b12 * q01 + b22 * q11 + dt * q12,
                         f2 = b12 * q02 + b22 * q12 + dt * q22;
            const double u00 = e0, u01 = e1, u02 = htv + b11 * t0 + b21 * t1;
            const double u10 = f0, u11 = f1, u12 = htw + b12 * t0 + b22 * t1 + dt * t2;
            const double r00 = hvv + dwk2 + Dsig[0] + b11 * e0 + b21 * e1;
            const double r01 = hvw + b11 * f0 + b21 * f1;
            const double r11 = hww + dwk2 + Dsig[1] + b12 * f0 + b22 * f1 + dt * f2;
            const bool first = (k == 0);
            const double gx0 = (first ? 0.0 : ra0) + w0;
            const double gx1 = (first ? 0.0 : ra1) + w1;
            const double gx2 = (first ? 0.0 : ra2) + a * w0 + b * w1 + w2;
            const double gu0 = qa[0] + b11 * w0 + b21 * w1;
            const double gu1 = qa[1] + b12 * w0 + b22 * w1 + dt * w2;
            // mu-coefficient of the right-hand side: d = 0, rx = 0, qu = qb
            const double hx2 = a * z0 + b * z1 + z2;
            const double hu0 = qb[0] + b11 * z0 + b21 * z1;
            const double hu1 = qb[1] + b12 * z0 + b22 * z1 + dt * z2;
            const double det = r00 * r11 - r01 * r01;
            if (!(r00 > 0.0) || !(det > 0.0)) ok = 0;
            const double idet = tpp_rcp(det);
            const double i00 = r11 * idet, i01 = -r01 * idet, i11 = r00 * idet;
            const double K00 = -(i00 * u00 + i01 * u10), K01 = -(i00 * u01 + i01 * u11), K02 = -(i00 * u02 + i01 * u12);
            const double K10 = -(i01 * u00 + i11 * u10), K11 = -(i01 * u01 + i11 * u11), K12 = -(i01 * u02 + i11 * u12);
            const double k0 = -(i00 * gu0 + i01 * gu1), k1 = -(i01 * gu0 + i11 * gu1);
            const double m0 = -(i00 * hu0 + i01 * hu1), m1 = -(i01 * hu0 + i11 * hu1);
            tpp_st2(p, R_K, K00, K01); tpp_st2(p, R_K + 1, K02, K10); tpp_st2(p, R_K + 2, K11, K12);
            tpp_st2(p, R_K + 3, k0, k1); tpp_st2(p, R_KB, m0, m1);
            q00 = x00 + u00 * K00 + u10 * K10;
            q11 = x11 + u01 * K01 + u11 * K11;
            q22 = x22 + u02 * K02 + u12 * K12;
            q01 = x01_ + 0.5 * ((u00 * K01 + u10 * K11) + (u01 * K00 + u11 * K10));
            q02 = x02 + 0.5 * ((u00 * K02 + u10 * K12) + (u02 * K00 + u12 * K10));
            q12 = x12 + 0.5 * ((u01 * K02 + u11 * K12) + (u02 * K01 + u12 * K11));
            v0 = gx0 + u00 * k0 + u10 * k1;
            v1 = gx1 + u01 * k0 + u11 * k1;
            v2 = gx2 + u02 * k0 + u12 * k1;
            const double y0_ = z0 + u00 * m0 + u10 * m1;
            const double y1_ = z1 + u01 * m0 + u11 * m1;
            const double y2_ = hx2 + u02 * m0 + u12 * m1;
            z0 = y0_; z1 = y1_; z2 = y2_;
        }
        Xn[0] = X[0]; Xn[1] = X[1]; Xn[2] = X[2];
        ln[0] = lam[0]; ln[1] = lam[1]; ln[2] = lam[2];
    }
    o.th = th;
    o.phi = df * fs - mu * slog;
    o.ymax = ymax; o.bad = !isfinite(chk);
    o.n.theta = th; o.n.prim_inf = pi; o.n.dual_inf = di; o.n.sum_y = sy; o.n.sum_z = sz;
    o.n.pmin = pmin; o.n.pmax = pmax; o.n.f = fs; o.n.slog = slog;
    ob.ok = ok; ob.bad = !isfinite(bchk); ob.gmax = gmax; ob.f = fs;
}

// Top of an interior-point iteration for the two-sweep kernel: convergence tests and barrier update only.  Returns
// false when the problem is finished (status set, phase FIN).
__device__ __forceinline__ bool tppf_iterate_top(const KParams &P, TppLane &L, const TppNorms &n) {
    const double dw = L.dw;
    const int bm = L.bmode;
    tpp_iterate_top(P, L, n); // (sets phase = PH_B / PH_FIN, dw = 0, bmode = NEWTON: restored / overridden here)
    L.dw = dw;
    L.bmode = bm;
    if (L.phase == PH_FIN) { L.phase = PHF_FIN; return false; }
    L.phase = PHF_TB;
    return true;
}

template <int SPEC>
__global__ void __launch_bounds__(TPP_THREADS, B200MPC_TPP_MIN_CTAS) mpc_solve_tppf_kernel(const KParams P, const TppArgs T) {
    extern __shared__ __align__(16) char tpp_smem[];
    const BatchArgs &A = T.a;
    const int N = P.N;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t gw = (size_t)blockIdx.x * (blockDim.x >> 5) + wid;
    char *wbase = reinterpret_cast<char *>(T.ws) + gw * ((size_t)(N + 1) * TPP_STAGE_B);
    char *wb = wbase + lane * 16;
    double *fl = T.filt + gw * (64 * 32) + lane;
    char *sb = tpp_smem + (size_t)wid * TPP_STAGE_SMEM + lane * 16;
    TppLane &L = *reinterpret_cast<TppLane *>(tpp_smem + (size_t)(TPP_THREADS / 32) * TPP_STAGE_SMEM +
                                              (size_t)threadIdx.x * TPP_LANE_STRIDE * sizeof(double));
    L.phase = PHF_LOAD;
    L.b = -1;
    L.moved = 0;
    int cur = 0; // warp-uniform: the buffer holding the current iterates at the start of this trip

    for (;;) {
        // ---- block L: pull the next problems; the warp writes each starting point together (lane <-> stage) ----
        __syncwarp();
        int newb = -1;
        if (tpp_opaque(L.phase) == PHF_LOAD) {
            const int b = (int)atomicAdd(A.counter, 1u);
            if (b >= A.B) {
                L.phase = PHF_DONE;
            } else {
                newb = b;
                L.b = b;
                if (T.avail) {
                    unsigned spins = 0;
                    while (*reinterpret_cast<const volatile unsigned *>(T.avail) <= (unsigned)b) {
                        if ((++spins & 1023u) == 0 && T.abort && *reinterpret_cast<const volatile unsigned *>(T.abort)) break; // call given up by the host
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
                L.tiny = 0; L.tiny_last = 0; L.tiny_flag = 0; // (the two-sweep kernel does not detect tiny steps)
                L.df = 1.0; L.mu = P.mu_init;
                L.theta0 = -1; L.dw = 0; L.dw_last = 0; L.dw_b = 0;
                L.alpha = 0; L.a_z = 0; L.alpha_soc = 0; L.a_z_soc = 0; L.a_min = 0; L.theta_soc_old = 0;
                L.ref_phi = 0; L.ref_gbd = 0; L.ymax_f = 0;
                L.f = 0; L.slog = 0; L.theta = 0;
                L.bmode = BM_LSQ;
                L.tmode = TMF_INIT;
                L.phase = PHF_TB;
            }
        }
        __syncwarp();
        for (unsigned m = __ballot_sync(FULL, newb >= 0); m; m &= m - 1) {
            const int j = __ffs(m) - 1;
            const size_t b = (size_t)__shfl_sync(FULL, newb, j);
            const double x00 = __ldcg(A.x0 + 3 * b), x01 = __ldcg(A.x0 + 3 * b + 1), x02 = __ldcg(A.x0 + 3 * b + 2);
            const double2 *ui = A.u_init ? reinterpret_cast<const double2 *>(A.u_init + b * 2 * N) : nullptr;
            for (int k = lane; k <= N; k += 32) {
                char *p = wbase + (size_t)k * TPP_STAGE_B + j * 16;
                char *pc = p + cur * R_ITER * TPP_ROW_B;
                tpp_st2(pc, R_X01, (k == 0) ? x00 : 0.0, (k == 0) ? x01 : 0.0);
                tpp_st2(pc, R_X2L0, (k == 0) ? x02 : 0.0, 0.0);
                tpp_st2(pc, R_L12, 0.0, 0.0);
                // (step rows are read by the first TB sweep in INIT mode: their content is not used, but must be finite-free safe)
                if (k < N) {
                    double2 u = make_double2(0.0, 0.0);
                    if (ui) u = __ldcg(ui + k);
                    const double uv[2] = {u.x, u.y};
                    double sv[2];
#pragma unroll
                    for (int i = 0; i < 2; i++) {
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
        if (__syncthreads_and(L.phase == PHF_DONE)) break;

        // ---- block TB ----
        TPP_BLOCK_SYNC();
        if (tpp_opaque(L.phase) == PHF_TB) {
            TppTrial t;
            TppBwd r;
            tpp_stat(T, 2);
            const int bm = L.bmode;
            const double dwb = (bm == BM_SOC) ? L.dw : L.dw_b;
            tppf_trial_backward<SPEC>(P, wb, sb, cur, L, dwb, t, r);
            const int tm = L.tmode;
            bool use_b = false; // the point of this sweep is the (new) current iterate: the B-part's output counts
            if (tm == TMF_INIT || tm == TMF_HOLD) {
                L.moved = 1;
                use_b = true;
            } else if (tm == TMF_EVAL) {
                L.moved = 1;
                use_b = tppf_iterate_top(P, L, t.n);
            } else if (tm == TMF_LSQ) {
                if (L.keep && !(L.df * t.ymax <= 1e3)) {
                    L.keep = 0; // discard the estimate: repeat the sweep with zeros
                } else {
                    L.moved = 1;
                    use_b = tppf_iterate_top(P, L, t.n);
                }
            } else if (t.bad) {
                L.status = B200MPC_ERROR_IN_STEP_COMPUTATION;
                L.phase = PHF_FIN;
            } else {
                const bool soc = (tm == TMF_STEP_SOC);
                if (soc || L.ntrial++ > 0) L.ls_extra++;
                bool fa;
                if (tpp_ls_acceptable(L, fl, L.alpha, t.phi, t.th, fa)) {
                    if (!fa) tpp_filter_add(fl, L, L.ref_phi - GAMMA_PHI * L.theta, (1 - GAMMA_THETA) * L.theta);
                    L.iter++;
                    L.moved = 1;
                    use_b = tppf_iterate_top(P, L, t.n);
                } else if (!soc) {
                    if (L.ntrial == 1 && P.max_soc > 0 && isfinite(t.th) && t.th >= L.theta) {
                        // second-order correction, first round: right-hand sides from this trial point
                        L.soc_count = 0;
                        L.soc_first = 1;
                        L.theta_soc_old = t.th;
                        L.bmode = BM_SOC;
                        L.tmode = TMF_HOLD;
                    } else {
                        L.phase = PHF_BACKTRACK;
                    }
                } else {
                    L.soc_count++;
                    if (L.soc_count < P.max_soc && t.th <= KAPPA_SOC * L.theta_soc_old) {
                        L.theta_soc_old = t.th;
                        L.soc_first = 0;
                        L.bmode = BM_SOC;
                        L.tmode = TMF_HOLD;
                    } else {
                        L.phase = PHF_BACKTRACK;
                    }
                }
            }
            if (use_b) {
                // what block B of the three-sweep kernel does with the outcome of the factorisation
                if (bm == BM_LSQ) {
                    L.f = r.f;
                    if (r.bad || !isfinite(r.f)) {
                        L.status = B200MPC_INVALID_NUMBER_DETECTED;
                        L.phase = PHF_FIN;
                    } else {
                        if (r.gmax > 100.0) L.df = fmax(100.0 / r.gmax, 1e-8);
                        if (r.ok) L.phase = PHF_F;
                        else { L.keep = 0; L.tmode = TMF_LSQ; L.bmode = BM_NEWTON; L.dw_b = 0.0; }
                    }
                } else if (r.ok) {
                    if (bm == BM_NEWTON) {
                        if (L.dw_b > 0.0) L.dw_last = L.dw_b;
                        L.dw = L.dw_b;
                    }
                    L.phase = PHF_F;
                } else if (bm == BM_SOC) {
                    L.phase = PHF_BACKTRACK; // correction abandoned: continue with the Newton direction
                } else {
                    // inertia correction: repeat the sweep on the same point with a larger delta_w
                    const double dw = L.dw_b, dwl = L.dw_last;
                    double nd;
                    if (dw == 0.0) nd = (dwl == 0.0) ? DW_INIT : fmax(DW_MIN, dwl * DW_DEC);
                    else nd = (dwl == 0.0 || 1e5 * dwl < dw) ? dw * DW_INC_FIRST : dw * DW_INC;
                    L.dw_b = nd;
                    L.tmode = TMF_HOLD;
                    if (nd > DW_MAX) { L.status = B200MPC_ERROR_IN_STEP_COMPUTATION; L.phase = PHF_FIN; }
                }
            }
        }

        // ---- block F (on the new iterate, buffer 1 - cur) ----
        TPP_BLOCK_SYNC();
        if (tpp_opaque(L.phase) == PHF_F) {
            TppFwd f;
            tpp_stat(T, 1);
            tppf_forward<SPEC>(P, wb, sb, 1 - cur, L, f);
            const int bmode = L.bmode;
            L.phase = PHF_TB;
            L.dw_b = 0.0;
            if (bmode == BM_LSQ) {
                L.ymax_f = f.ymax;
                L.keep = (L.df * f.ymax <= 1e3) ? 1 : 0;
                L.tmode = TMF_LSQ;
                L.bmode = BM_NEWTON;
            } else if (bmode == BM_NEWTON) {
                if (f.bad) {
                    L.status = B200MPC_ERROR_IN_STEP_COMPUTATION;
                    L.phase = PHF_FIN;
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
                    L.tmode = TMF_STEP;
                }
            } else {
                L.alpha_soc = f.a_max;
                L.a_z_soc = f.a_z;
                L.tmode = TMF_STEP_SOC;
                L.bmode = BM_NEWTON;
            }
        }

        // ---- rare: backtracking / restoration stand-in ----
        __syncwarp();
        if (tpp_opaque(L.phase) == PHF_BACKTRACK) {
            L.alpha *= 0.5;
            L.bmode = BM_NEWTON;
            L.dw_b = 0.0;
            if (L.alpha < L.a_min) {
                // buffer `cur` holds the current iterate (a HOLD sweep only copied it); the restored point goes to 1 - cur
                tpp_restore<SPEC>(P, wb, fl, cur, L);
                if (L.phase == PH_FIN) L.phase = PHF_FIN;
                else { L.tmode = TMF_EVAL; L.phase = PHF_TB; }
            } else {
                L.tmode = TMF_STEP;
                L.phase = PHF_TB;
            }
        }

        // ---- result store and release of finished lanes; iterate copy for lanes that did not move ----
        __syncwarp();
        {
            const int ph = tpp_opaque(L.phase);
            const bool fin = (ph == PHF_FIN);
            const bool cpy = (ph == PHF_TB) && !L.moved;
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
            if (fin) {
                const size_t b = (size_t)L.b;
                if (A.cost) A.cost[b] = L.f;
                A.status[b] = L.status;
                if (A.iters) A.iters[b] = L.iter;
                if (A.ls) A.ls[b] = L.ls_extra;
                L.phase = PHF_LOAD;
            }
            if (T.done && __any_sync(FULL, fin)) {
                __threadfence();
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
                }
            }
            L.moved = 0;
        }
        cur ^= 1;
        tpp_stat(T, 3);
    }
}
