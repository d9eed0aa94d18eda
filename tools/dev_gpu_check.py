"""Developer smoke script (run on the GPU box): eval parity, solve parity vs the oracle, rough timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ros2_mpc_b200 import _shim, synth, load_params, make_params
from oracle import oracle as O

y = load_params()
rng = np.random.default_rng(0)
nB = int(os.environ.get("NB", "256"))
w = synth.robots_on_map(B=nB, seed=0)
pxf, puf = synth.straight_reference(w["x0"], w["goal"], 30)

for var in "BCA":
    p = make_params(var, y); po = O.variant_params(var, y)
    S = _shim.Solver(p)
    N = p.N
    kw = {}; kwo = {}
    xr = w["goal"]
    if var == "A": kw = dict(obs_x=w["obs_x"], obs_y=w["obs_y"])
    if var == "C": xr = pxf; kw = dict(uref=puf)
    # ---- eval parity at random points
    X = rng.normal(0, 1, (nB, N + 1, 3)) + w["x0"][:, None, :]
    X[:, 0, :] = w["x0"]
    U = rng.uniform(-0.2, 0.2, (nB, N, 2)); lam = rng.normal(0, 1, (nB, N, 3))
    g = S.eval_batch(w["x0"], xr, X, U, lam=lam, **kw)
    md = {k: 0.0 for k in ("f", "c", "grad", "stages")}
    for b in range(min(nB, 64)):
        kwb = {k: v[b] for k, v in kw.items()}
        o = O.evaluate(po, w["x0"][b], xr[b], X[b], U[b], lam=lam[b], **kwb)
        for k in md:
            a, r = np.asarray(g[k][b]), np.asarray(o[k])
            fin = np.isfinite(r)
            md[k] = max(md[k], float(np.max(np.abs(a[fin] - r[fin]) / (1 + np.abs(r[fin])))) if fin.any() else 0.0)
    print(var, "eval rel diffs", md, flush=True)
    # ---- solve parity
    t = time.time(); out = S.solve_batch(w["x0"], xr, **kw); tg = time.time() - t
    t = time.time(); ref = O.solve_batch(po, w["x0"], xr, **kw); tc = time.time() - t
    same = out["status"] == ref["status"]
    conv = (ref["status"] == 0) & (out["status"] == 0)
    dc = np.abs(out["cost"] - ref["cost"]) / np.abs(ref["cost"])
    dU = np.abs(out["U"] - ref["U"]).reshape(nB, -1).max(1); dX = np.abs(out["X"] - ref["X"]).reshape(nB, -1).max(1)
    print(var, "status agree", same.mean(), "both conv", conv.mean(), "gpu status", np.unique(out["status"], return_counts=True),
          "cpu status", np.unique(ref["status"], return_counts=True))
    if conv.any():
        print(var, " on converged: max rel dcost", dc[conv].max(), "max dU", dU[conv].max(), "max dX", dX[conv].max(),
              "iters gpu/cpu", out["iters"][conv].mean(), ref["iters"][conv].mean(), "iter mismatch frac", (out["iters"] != ref["iters"])[conv].mean())
        bad = conv & ((dc > 1e-5) | (dU > 1e-4) | (dX > 1e-4))
        print(var, " parity violations", int(bad.sum()), "of", int(conv.sum()))
    print(var, "first solve wall gpu %.3fs cpu %.3fs kernel %.3f ms" % (tg, tc, S.last_kernel_ms()), flush=True)
    # ---- timing at larger batch (tile the problems)
    for rep in (16, 128):
        Bb = nB * rep
        x0b = np.tile(w["x0"], (rep, 1)); xrb = np.tile(xr, (rep, 1))
        kwb = {k: np.tile(v, (rep, 1)) for k, v in kw.items()}
        S.solve_batch(x0b, xrb, **kwb)
        t = time.time(); o2 = S.solve_batch(x0b, xrb, **kwb); tg = time.time() - t
        km = S.last_kernel_ms()
        print(var, "B=%d e2e %.1f ms -> %.0f solves/s ; kernel %.2f ms -> %.0f solves/s ; iters mean %.1f max %d" % (
            Bb, tg * 1e3, Bb / tg, km, Bb / km * 1e3, o2["iters"].mean(), o2["iters"].max()), flush=True)
    S.close()
