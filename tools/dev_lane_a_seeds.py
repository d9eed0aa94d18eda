"""GPU box: variant A, bench workload (4096 robots x SEEDS warm-start seeds), lane kernel with different hand-over thresholds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from ros2_mpc_b200 import _shim, load_params, make_params
y = load_params(); p = make_params("A", y)
wl = bench.build_workload("A", 4096, int(os.environ.get("SEEDS", "64")), 0, y)
for hi in [int(v) for v in os.environ.get("HANDS", "96,128,192,256").split(",")]:
    os.environ["B200MPC_HAND_ITER"] = str(hi)
    S = _shim.Solver(p)
    for rep in range(2):
        o = S.solve_batch(wl["x0"], wl["xref"], obs_x=wl["obs_x"], obs_y=wl["obs_y"], u_init=wl["u_init"])
    conv = np.isin(o["status"], (0, 1)).mean()
    print("hand_iter %d: kernel %.1f ms, %.0f converged/s, iters %.1f max %d, kind %d" % (hi, S.last_kernel_ms(), conv * wl["B"] / S.last_kernel_ms() * 1e3,
          o["iters"].mean(), o["iters"].max(), S.last_kernel_kind), flush=True)
    S.close()
