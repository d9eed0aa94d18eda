"""GPU box: variant A (obstacle cost active, non-convex) on 4096 map problems, warp kernel vs the CPU oracle."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ros2_mpc_b200 import _shim, synth, load_params, make_params
from oracle import oracle as O

y = load_params()
for seed in (0, 21):
    w = synth.robots_on_map(B=4096, seed=seed)
    S = _shim.Solver(make_params("A", y))
    out = S.solve_batch(w["x0"], w["goal"], obs_x=w["obs_x"], obs_y=w["obs_y"])
    t = time.time()
    ref = O.solve_batch(O.variant_params("A", y), w["x0"], w["goal"], obs_x=w["obs_x"], obs_y=w["obs_y"], nthreads=len(os.sched_getaffinity(0)))
    tc = time.time() - t
    S.close()
    same_status = out["status"] == ref["status"]
    both = np.isin(out["status"], (0, 1)) & np.isin(ref["status"], (0, 1))
    dc = np.abs(out["cost"] - ref["cost"]) / np.abs(ref["cost"])
    dU = np.abs(out["U"] - ref["U"]).reshape(4096, -1).max(1)
    dX = np.abs(out["X"] - ref["X"]).reshape(4096, -1).max(1)
    par = (dc <= 1e-5) & (dU <= 1e-4) & (dX <= 1e-4)
    print(json.dumps({"variant": "A", "seed": seed, "problems": 4096, "status_identical_frac": float(same_status.mean()),
                      "both_converged_frac": float(both.mean()), "parity_frac_of_both_converged": float(par[both].mean()),
                      "same_cost_1e-5_frac": float((dc <= 1e-5)[both].mean()),
                      "iters_identical_frac": float((out["iters"] == ref["iters"])[both].mean()),
                      "status_counts_gpu": {int(k): int(v) for k, v in zip(*np.unique(out["status"], return_counts=True))},
                      "oracle_seconds": tc}), flush=True)
