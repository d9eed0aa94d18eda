"""Developer script (GPU box): lane kernel vs warp kernel on the config-4 style batch; oracle on the mismatches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ros2_mpc_b200 import _shim, synth, load_params, make_params
from oracle import oracle as O
y = load_params()
R, Sd, N = 4096, int(os.environ.get("SEEDS", "64")), 30
w = synth.robots_on_map(B=R, seed=0)
p = make_params("B", y)
S = _shim.Solver(p)
ui = synth.warm_start_seeds(Sd, N, list(p.u_lo), list(p.u_hi))
B = R * Sd
x0 = np.tile(w["x0"], (Sd, 1)); goal = np.tile(w["goal"], (Sd, 1))
u_init = np.repeat(ui, R, axis=0).reshape(B, N, 2)
S.set_kernel(_shim.KERNEL_LANE); a = S.solve_batch(x0, goal, u_init=u_init)
S.set_kernel(_shim.KERNEL_WARP); b = S.solve_batch(x0, goal, u_init=u_init)
dc = np.abs(a["cost"] - b["cost"]) / np.abs(b["cost"])
dU = np.abs(a["U"] - b["U"]).reshape(B, -1).max(1)
bad = np.where((dc > 1e-5) | (dU > 1e-4) | (a["status"] != b["status"]))[0]
print("B", B, "mismatches", len(bad), "iter mismatch frac", (a["iters"] != b["iters"]).mean(), "ls mismatch", (a["ls"] != b["ls"]).mean())
print("status lane", np.unique(a["status"], return_counts=True), "warp", np.unique(b["status"], return_counts=True))
for i in bad[:12]:
    r = O.solve(O.variant_params("B", y), x0[i], goal[i], u_init=u_init[i].reshape(-1))
    print(i, "lane: st %d it %d ls %d cost %.9f | warp: st %d it %d ls %d cost %.9f | oracle: st %d it %d cost %.9f" % (
        a["status"][i], a["iters"][i], a["ls"][i], a["cost"][i], b["status"][i], b["iters"][i], b["ls"][i], b["cost"][i],
        r["status"], r["iters"], r["cost"]))
# where do iteration counts differ
dm = np.where(a["iters"] != b["iters"])[0]
print("iter mismatches", len(dm), "examples", [(int(i), int(a["iters"][i]), int(b["iters"][i]), int(a["ls"][i]), int(b["ls"][i])) for i in dm[:10]])
S.close()
S = _shim.Solver(p)
cold = S.solve_batch(w["x0"], w["goal"])
c = a["cost"].reshape(Sd, R)
rel = np.abs(c - cold["cost"][None]) / cold["cost"][None]
print("warm vs cold: frac rel>1e-5:", (rel > 1e-5).mean(), "robots affected", (rel > 1e-5).any(0).sum(), "max", rel.max())
dUc = np.abs(a["U"].reshape(Sd, R, -1) - cold["U"].reshape(1, R, -1)).max(2)
print("frac dU>1e-4", (dUc > 1e-4).mean(), "robots", (dUc > 1e-4).any(0).sum())
j = np.argwhere(rel > 1e-5)[:5]
for s_, r_ in j: print("seed", s_, "robot", r_, "warm cost", c[s_, r_], "cold", cold["cost"][r_])
S.close()
