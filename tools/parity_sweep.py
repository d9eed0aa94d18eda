"""GPU box: wide randomised parity sweep of both solve kernels against the CPU oracle (variants B and C, fresh seeds,
cold starts and random warm-start controls).  Prints one JSON line per case; non-zero exit on a parity violation."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ros2_mpc_b200 import _shim, synth, load_params, make_params
from oracle import oracle as O

y = load_params()
N = y["N"]
bad_total = 0
for var in ("B", "C"):
    S = _shim.Solver(make_params(var, y))
    po = O.variant_params(var, y)
    for seed in range(10, 10 + int(os.environ.get("SEEDS", "6"))):
        w = synth.robots_on_map(B=4096, seed=seed)
        xr, kw = w["goal"], {}
        if var == "C":
            pxf, puf = synth.straight_reference(w["x0"], w["goal"], N)
            xr, kw = pxf, dict(uref=puf)
        rng = np.random.default_rng(seed)
        ui = np.clip(rng.normal(0, 0.08, (4096, N, 2)), [-0.05, -0.2], [0.15, 0.2]) if seed % 2 else None
        ref = O.solve_batch(po, w["x0"], xr, u_init=None if ui is None else ui.reshape(4096, -1), **kw)
        for kind, name in ((_shim.KERNEL_LANE, "lane"), (_shim.KERNEL_WARP, "warp")):
            S.set_kernel(kind)
            out = S.solve_batch(w["x0"], xr, u_init=ui, **kw)
            ok = np.isin(ref["status"], (0, 1))
            st = bool(np.array_equal(out["status"], ref["status"]))
            dc = np.abs(out["cost"] - ref["cost"]) / np.abs(ref["cost"])
            dU = np.abs(out["U"] - ref["U"]).reshape(4096, -1).max(1)
            dX = np.abs(out["X"] - ref["X"]).reshape(4096, -1).max(1)
            viol = int(((dc > 1e-5) | (dU > 1e-4) | (dX > 1e-4))[ok].sum()) + (0 if st else 1)
            bad_total += viol
            print(json.dumps({"variant": var, "seed": seed, "kernel": name, "u_init": ui is not None, "problems": 4096,
                              "converged": float(ok.mean()), "status_identical": st, "violations": viol,
                              "max_rel_dcost": float(dc[ok].max()), "max_dU": float(dU[ok].max()), "max_dX": float(dX[ok].max()),
                              "iters_identical_frac": float((out["iters"] == ref["iters"]).mean()),
                              "iters_mean": float(out["iters"].mean())}), flush=True)
    S.close()
print("TOTAL violations", bad_total)
sys.exit(1 if bad_total else 0)
