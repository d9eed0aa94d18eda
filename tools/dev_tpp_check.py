"""Developer script (GPU box): lane-per-problem kernel vs the oracle and vs the warp kernel, rough timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ros2_mpc_b200 import _shim, synth, load_params, make_params
from oracle import oracle as O

y = load_params()
nB = int(os.environ.get("NB", "512"))
w = synth.robots_on_map(B=nB, seed=0)
pxf, puf = synth.straight_reference(w["x0"], w["goal"], 30)
ui = synth.warm_start_seeds(4, 30, [-0.05, -0.2], [0.15, 0.2], first_seed=1)

for var in os.environ.get("VARS", "BC"):
    p = make_params(var, y); po = O.variant_params(var, y)
    S = _shim.Solver(p)
    N = p.N
    kw = {}
    xr = w["goal"]
    if var == "C": xr = pxf; kw = dict(uref=puf)
    for use_ui in (False, True):
        kk = dict(kw)
        if use_ui:
            kk["u_init"] = np.repeat(ui, nB // 4, axis=0).reshape(nB, N, 2)
        S.set_kernel(_shim.KERNEL_LANE)
        out = S.solve_batch(w["x0"], xr, **kk)
        assert S.last_kernel_kind == _shim.KERNEL_LANE
        S.set_kernel(_shim.KERNEL_WARP)
        wout = S.solve_batch(w["x0"], xr, **kk)
        ko = dict(kk)
        if use_ui: ko["u_init"] = kk["u_init"].reshape(nB, -1)
        ref = O.solve_batch(po, w["x0"], xr, **ko)
        for name, r in (("oracle", ref), ("warp", wout)):
            same = out["status"] == r["status"]
            conv = (r["status"] == 0) & (out["status"] == 0)
            dc = np.abs(out["cost"] - r["cost"]) / np.abs(r["cost"])
            dU = np.abs(out["U"] - r["U"]).reshape(nB, -1).max(1); dX = np.abs(out["X"] - r["X"]).reshape(nB, -1).max(1)
            print(var, "ui" if use_ui else "cold", "lane vs", name, "status agree", same.mean(), "conv", conv.mean(),
                  "max rel dcost %.2e max dU %.2e max dX %.2e" % (dc[conv].max(), dU[conv].max(), dX[conv].max()),
                  "iters lane/ref %.3f %.3f mismatch %.4f ls %.3f %.3f" % (out["iters"][conv].mean(), r["iters"][conv].mean(),
                   (out["iters"] != r["iters"])[conv].mean(), out["ls"].mean(), r["ls"].mean()),
                  "status", np.unique(out["status"], return_counts=True), flush=True)
    # timing
    for rep in (32, 512, 2048):
        Bb = nB * rep
        x0b = np.tile(w["x0"], (rep, 1)); xrb = np.tile(xr, (rep, 1))
        kwb = {k: np.tile(v, (rep, 1)) for k, v in kw.items()}
        for kind, nm in ((_shim.KERNEL_LANE, "lane"), (_shim.KERNEL_WARP, "warp")):
            S.set_kernel(kind)
            S.solve_batch(x0b, xrb, **kwb)
            t = time.time(); o2 = S.solve_batch(x0b, xrb, **kwb); tg = time.time() - t
            km = S.last_kernel_ms()
            print(var, nm, "B=%d e2e %.1f ms ; kernel %.2f ms -> %.0f solves/s ; iters mean %.2f max %d conv %.4f" % (
                Bb, tg * 1e3, km, Bb / km * 1e3, o2["iters"].mean(), o2["iters"].max(), np.isin(o2["status"], (0, 1)).mean()), flush=True)
    S.close()
