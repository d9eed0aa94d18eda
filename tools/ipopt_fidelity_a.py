"""Evidence table for the obstacle-active variant A (task 3): on config-3 problems, does a first-order
(KKT) point exist where the oracle's interior-point emulation gives up, and does the oracle find the same point when
it does converge?

For every sampled problem (cold start, as the reference):
  1. oracle/mpc_oracle.c (IPOPT emulation, closed-form restoration stand-in) -> status, cost, iterations, and the KKT
     certificate of the point it RETURNED (oracle.kkt_certificate: defect, scaled stationarity, complementarity);
  2. an independent third-party solve of the same NLP in single-shooting form (controls only, states rolled out with the
     RK4 closed form, exact gradient by the adjoint recursion from orc_eval's stage Jacobians): scipy L-BFGS-B from the
     same cold start (U = 0) and, for problems the oracle failed on, also from the oracle's returned U;
  3. scipy SLSQP on the multiple-shooting form from the reference's cold start (X = 0, U = 0), N as given.
Output: one JSON line per problem + a summary line (profiles/r2_ipopt_fidelity_variant_a.jsonl).
CPU only; uses the oracle (test infrastructure), never the product library."""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rollout(p, x0, U):
    """X_{k+1} = RK4(X_k, U_k) in the closed form of SURVEY App. B."""
    N, dt = p.N, p.dt
    X = np.zeros((N + 1, 3)); X[0] = x0
    for k in range(N):
        x, y, th = X[k]; v, w = U[k]
        tm, te = th + 0.5 * dt * w, th + dt * w
        C = np.cos(th) + 4 * np.cos(tm) + np.cos(te)
        S = np.sin(th) + 4 * np.sin(tm) + np.sin(te)
        X[k + 1] = (x + dt * v * C / 6.0, y + dt * v * S / 6.0, te)
    return X


def make_ss(O, p, x0, goal, ox, oy):
    N = p.N

    def fg(u):
        U = u.reshape(N, 2)
        X = rollout(p, x0, U)
        e = O.evaluate(p, x0, goal, X, U, obs_x=ox, obs_y=oy)
        if not np.isfinite(e["f"]):
            return 1e300, np.zeros(2 * N)
        g, st = e["grad"], e["stages"]
        lam = np.zeros((N + 2, 3)); gu = np.zeros((N, 2))
        for k in range(N, 0, -1):
            gk = g[3 * (k - 1):3 * (k - 1) + 3]
            if k == N:
                lam[k] = -gk
            else:
                A = np.array([[1, 0, st[k, 0]], [0, 1, st[k, 1]], [0, 0, 1]])
                lam[k] = A.T @ lam[k + 1] - gk
        for k in range(N):
            b11, b12, b21, b22 = st[k, 2:6]
            Bm = np.array([[b11, b12], [b21, b22], [0, p.dt]])
            gu[k] = g[3 * N + 2 * k:3 * N + 2 * k + 2] - Bm.T @ lam[k + 1]
        return e["f"], gu.ravel()
    return fg


def work(args):
    i, x0, goal, ox, oy, params = args
    from scipy.optimize import minimize
    from oracle import oracle as O
    p = O.variant_params("A", params)
    N = p.N
    t = time.time()
    r = O.solve(p, x0, goal, obs_x=ox, obs_y=oy)
    rec = {"i": int(i), "oracle_status": int(r["status"]), "oracle_iters": int(r["stats"]["iters"]),
           "oracle_resto": int(r["stats"]["n_resto"]), "oracle_cost": float(r["cost"])}
    Xo, Uo = r["X"].T.copy(), r["U"].T.copy()
    if np.all(np.isfinite(Xo)) and np.all(np.isfinite(Uo)):
        kc = O.kkt_certificate(p, x0, goal, Xo, Uo, obs_x=ox, obs_y=oy)
        rec["oracle_point"] = {k: float(kc[k]) for k in ("defect", "scaled", "complementarity")}
    fg = make_ss(O, p, x0, goal, ox, oy)
    bounds = [(p.u_lo[0], p.u_hi[0]), (p.u_lo[1], p.u_hi[1])] * N
    starts = {"cold": np.zeros(2 * N)}
    if r["status"] not in (0, 1) and np.all(np.isfinite(Uo)):
        starts["from_oracle_point"] = np.clip(Uo.ravel(), [b[0] for b in bounds], [b[1] for b in bounds])
    for name, u0 in starts.items():
        f0, _ = fg(u0)
        if not np.isfinite(f0) or f0 >= 1e300:
            rec["lbfgsb_" + name] = {"start_invalid": True}
            continue
        res = minimize(fg, u0, jac=True, method="L-BFGS-B", bounds=bounds,
                       options={"maxiter": 3000, "maxfun": 20000, "ftol": 1e-15, "gtol": 1e-9, "maxcor": 30})
        U = res.x.reshape(N, 2)
        X = rollout(p, x0, U)
        kc = O.kkt_certificate(p, x0, goal, X, U, obs_x=ox, obs_y=oy)
        rec["lbfgsb_" + name] = {"cost": float(res.fun), "nit": int(res.nit), "scaled_stationarity": float(kc["scaled"]),
                                 "complementarity": float(kc["complementarity"]), "kkt_point": bool(kc["scaled"] <= 1e-6 and kc["complementarity"] <= 1e-6)}
    # SLSQP on the multiple-shooting form from the reference's cold start
    def f_ms(z):
        X = np.vstack([x0, z[:3 * N].reshape(N, 3)]); U = z[3 * N:].reshape(N, 2)
        e = O.evaluate(p, x0, goal, X, U, obs_x=ox, obs_y=oy)
        if not np.isfinite(e["f"]):
            return 1e300, np.zeros(5 * N)
        return e["f"], e["grad"]

    def c_ms(z):
        X = np.vstack([x0, z[:3 * N].reshape(N, 3)]); U = z[3 * N:].reshape(N, 2)
        return O.evaluate(p, x0, goal, X, U, obs_x=ox, obs_y=oy)["c"].ravel()

    z0 = np.zeros(5 * N)
    f0, _ = f_ms(z0)
    if np.isfinite(f0) and f0 < 1e300:
        lo = np.r_[np.full(3 * N, -np.inf), np.tile([p.u_lo[0], p.u_lo[1]], N)]
        hi = np.r_[np.full(3 * N, np.inf), np.tile([p.u_hi[0], p.u_hi[1]], N)]
        try:
            res = minimize(f_ms, z0, jac=True, method="SLSQP", bounds=list(zip(lo, hi)),
                           constraints=[{"type": "eq", "fun": c_ms}], options={"maxiter": 400, "ftol": 1e-13})
            X = np.vstack([x0, res.x[:3 * N].reshape(N, 3)]); U = res.x[3 * N:].reshape(N, 2)
            kc = O.kkt_certificate(p, x0, goal, X, U, obs_x=ox, obs_y=oy)
            rec["slsqp_cold"] = {"success": bool(res.success), "cost": float(res.fun), "nit": int(res.nit), "defect": float(kc["defect"]),
                                 "scaled_stationarity": float(kc["scaled"]),
                                 "kkt_point": bool(kc["defect"] <= 1e-6 and kc["scaled"] <= 1e-6 and kc["complementarity"] <= 1e-6)}
        except Exception as ex:  # noqa: BLE001
            rec["slsqp_cold"] = {"error": str(ex)[:80]}
    else:
        rec["slsqp_cold"] = {"start_invalid": True}
    rec["seconds"] = time.time() - t
    return rec


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 640
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r2_ipopt_fidelity_variant_a.jsonl")
    from ros2_mpc_b200 import load_params, synth
    params = load_params()
    w = synth.robots_on_map(B=4096, seed=0, params=params)
    # a first pass with the oracle picks the sample: every failing problem among the first 2048 up to n/2, the rest converged
    from oracle import oracle as O
    p = O.variant_params("A", params)
    r = O.solve_batch(p, w["x0"][:2048], w["goal"][:2048], obs_x=w["obs_x"][:2048], obs_y=w["obs_y"][:2048])
    bad = np.nonzero(~np.isin(r["status"], (0, 1)))[0]
    good = np.nonzero(np.isin(r["status"], (0, 1)))[0]
    rng = np.random.default_rng(5)
    pick = np.r_[rng.permutation(bad)[:n // 2], rng.permutation(good)[:n - min(len(bad), n // 2)]]
    jobs = [(i, w["x0"][i], w["goal"][i], w["obs_x"][i], w["obs_y"][i], params) for i in pick]
    nproc = int(os.environ.get("NPROC", "5"))
    recs = []
    with mp.Pool(nproc) as pool, open(out, "w") as f:
        for rec in pool.imap_unordered(work, jobs, chunksize=4):
            recs.append(rec)
            f.write(json.dumps(rec) + "\n"); f.flush()
        # ---- summary ----
        S = {"summary": True, "problems": len(recs), "workload": "config 3 (map_carto, seed 0), variant A, cold start"}
        by = {}
        for rec in recs:
            s = rec["oracle_status"]
            b = by.setdefault(s, {"n": 0, "oracle_point_is_kkt_1e-6": 0, "lbfgsb_cold_finds_kkt": 0, "lbfgsb_cold_same_cost_1e-5": 0,
                                  "lbfgsb_cold_lower_cost": 0, "slsqp_cold_finds_kkt": 0, "slsqp_same_cost_1e-5": 0,
                                  "cold_start_invalid": 0, "lbfgsb_from_oracle_point_kkt": 0, "from_oracle_point_cost_drop_rel_median": []})
            b["n"] += 1
            op = rec.get("oracle_point")
            if op and op["defect"] <= 1e-6 and op["scaled"] <= 1e-6 and op["complementarity"] <= 1e-6:
                b["oracle_point_is_kkt_1e-6"] += 1
            lc = rec.get("lbfgsb_cold", {})
            if lc.get("start_invalid"):
                b["cold_start_invalid"] += 1
            if lc.get("kkt_point"):
                b["lbfgsb_cold_finds_kkt"] += 1
                rel = (lc["cost"] - rec["oracle_cost"]) / abs(rec["oracle_cost"]) if np.isfinite(rec["oracle_cost"]) and rec["oracle_cost"] else np.nan
                if abs(rel) <= 1e-5:
                    b["lbfgsb_cold_same_cost_1e-5"] += 1
                elif rel < -1e-5:
                    b["lbfgsb_cold_lower_cost"] += 1
            sc = rec.get("slsqp_cold", {})
            if sc.get("kkt_point"):
                b["slsqp_cold_finds_kkt"] += 1
                if np.isfinite(rec["oracle_cost"]) and abs(sc["cost"] - rec["oracle_cost"]) <= 1e-5 * abs(rec["oracle_cost"]):
                    b["slsqp_same_cost_1e-5"] += 1
            lo = rec.get("lbfgsb_from_oracle_point", {})
            if lo.get("kkt_point"):
                b["lbfgsb_from_oracle_point_kkt"] += 1
                if np.isfinite(rec["oracle_cost"]) and rec["oracle_cost"]:
                    b["from_oracle_point_cost_drop_rel_median"].append((rec["oracle_cost"] - lo["cost"]) / abs(rec["oracle_cost"]))
        for s, b in by.items():
            v = b["from_oracle_point_cost_drop_rel_median"]
            b["from_oracle_point_cost_drop_rel_median"] = float(np.median(v)) if v else None
        S["by_oracle_status"] = {str(k): v for k, v in sorted(by.items())}
        f.write(json.dumps(S) + "\n")
        print(json.dumps(S, indent=1))


if __name__ == "__main__":
    main()
