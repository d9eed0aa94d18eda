"""Profiling driver (GPU box): NREP solves of a tiled map batch with one kernel kind.  Used under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ros2_mpc_b200 import _shim, synth, load_params, make_params

y = load_params()
var = os.environ.get("VAR", "B")
kind = {"lane": _shim.KERNEL_LANE, "warp": _shim.KERNEL_WARP, "auto": _shim.KERNEL_AUTO}[os.environ.get("KIND", "lane")]
B = int(os.environ.get("B", "131072"))
nrep = int(os.environ.get("NREP", "2"))
w = synth.robots_on_map(B=4096, seed=0)
rep = (B + 4095) // 4096
x0 = np.tile(w["x0"], (rep, 1))[:B]; xr = np.tile(w["goal"], (rep, 1))[:B]
kw = {}
if var == "C":
    pxf, puf = synth.straight_reference(w["x0"], w["goal"], 30)
    xr = np.tile(pxf, (rep, 1))[:B]; kw = dict(uref=np.tile(puf, (rep, 1))[:B])
if var == "A":
    kw = dict(obs_x=np.tile(w["obs_x"], (rep, 1))[:B], obs_y=np.tile(w["obs_y"], (rep, 1))[:B])
S = _shim.Solver(make_params(var, y))
S.set_kernel(kind)
for _ in range(nrep):
    o = S.solve_batch(x0, xr, **kw)
    print("kernel ms %.3f -> %.0f solves/s, iters %.2f, conv %.4f" % (S.last_kernel_ms(), B / S.last_kernel_ms() * 1e3, o["iters"].mean(),
          np.isin(o["status"], (0, 1)).mean()), flush=True)
print("lane kernel stats", S.lane_kernel_stats())
S.close()
