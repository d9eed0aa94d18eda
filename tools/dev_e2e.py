"""GPU box: host-buffer solve (b200mpc_solve_batch, pinned buffers) next to the device-buffer solve of the same batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ros2_mpc_b200 import _shim, load_params, make_params
y = load_params(); p = make_params("B", y)
wl = bench.build_workload("B", 4096, int(os.environ.get("SEEDS", "256")), 0, y)
B, N = wl["B"], y["N"]
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
h = {k: pin(wl[k]) for k in ("x0", "xref", "u_init")}
out = dict(X=pin(np.empty((B, N + 1, 3))), U=pin(np.empty((B, N, 2))), cost=pin(np.empty(B)), status=pin(np.empty(B, np.int32)),
           iters=pin(np.empty(B, np.int32)), ls=pin(np.empty(B, np.int32)))
S = _shim.Solver(p)
for rep in range(4):
    t = time.perf_counter()
    S.solve_batch(h["x0"], h["xref"], u_init=h["u_init"], out=out)
    dt = (time.perf_counter() - t) * 1e3
    print("host-buffer call %.1f ms, kernel %.1f ms, chunks %d" % (dt, S.last_kernel_ms(), S.last_solve_chunks), flush=True)
S.close()
