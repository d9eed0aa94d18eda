"""GPU box: the lane-per-problem kernel with the obstacle cost (variant A) against the warp kernel and the oracle, and its
throughput on config-4-style batches."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from ros2_mpc_b200 import _shim, load_params, make_params
from oracle import oracle as O

y = load_params()
p = make_params("A", y)
seeds = int(os.environ.get("SEEDS", "2"))
wl = bench.build_workload("A", 4096, seeds, 0, y)
B = wl["B"]
kw = dict(obs_x=wl["obs_x"], obs_y=wl["obs_y"], u_init=wl["u_init"])
S = _shim.Solver(p)
res = {}
for name, kind in (("warp", _shim.KERNEL_WARP), ("lane", _shim.KERNEL_LANE)):
    S.set_kernel(kind)
    o = S.solve_batch(wl["x0"], wl["xref"], **kw)
    o = S.solve_batch(wl["x0"], wl["xref"], **kw)
    res[name] = o
    print(name, "kernel ms %.2f" % S.last_kernel_ms(), "solves/s %.0f" % (B / S.last_kernel_ms() * 1e3),
          "status", dict(zip(*[a.tolist() for a in np.unique(o["status"], return_counts=True)])), "iters %.2f" % o["iters"].mean(), flush=True)
w, l = res["warp"], res["lane"]
same = w["status"] == l["status"]
both = np.isin(w["status"], (0, 1)) & np.isin(l["status"], (0, 1))
with np.errstate(invalid="ignore", divide="ignore"):
    dc = np.abs(w["cost"] - l["cost"]) / np.abs(w["cost"])
    dU = np.abs(w["U"] - l["U"]).reshape(B, -1).max(1)
    dX = np.abs(w["X"] - l["X"]).reshape(B, -1).max(1)
par = (dc <= 1e-5) & (dU <= 1e-4) & (dX <= 1e-4)
print(json.dumps({"lane_vs_warp": {"problems": int(B), "status_identical": float(same.mean()), "both_converged": float(both.mean()),
                                   "parity_of_both_converged": float(par[both].mean()),
                                   "iters_identical_of_both": float((w["iters"] == l["iters"])[both].mean()),
                                   "worst_dU_where_parity": float(dU[both & par].max())}}), flush=True)
n = min(B, 1024)
idx = np.arange(0, B, B // n)[:n]
ref = O.solve_batch(O.variant_params("A", y), wl["x0"][idx], wl["xref"][idx], obs_x=wl["obs_x"][idx], obs_y=wl["obs_y"][idx],
                    u_init=wl["u_init"][idx].reshape(len(idx), -1))
for name in ("warp", "lane"):
    o = res[name]
    same = o["status"][idx] == ref["status"]
    both = np.isin(o["status"][idx], (0, 1)) & np.isin(ref["status"], (0, 1))
    with np.errstate(invalid="ignore", divide="ignore"):
        dc = np.abs(o["cost"][idx] - ref["cost"]) / np.abs(ref["cost"])
        dU = np.abs(o["U"][idx] - ref["U"]).reshape(len(idx), -1).max(1)
        dX = np.abs(o["X"][idx] - ref["X"]).reshape(len(idx), -1).max(1)
    par = (dc <= 1e-5) & (dU <= 1e-4) & (dX <= 1e-4)
    print(json.dumps({name + "_vs_oracle": {"sample": int(len(idx)), "status_identical": float(same.mean()), "parity_of_both_converged": float(par[both].mean()),
                                            "iters_identical_of_both": float((o["iters"][idx] == ref["iters"])[both].mean())}}), flush=True)
# throughput on larger batches (device time of the kernel)
for sd in [int(v) for v in os.environ.get("BIG", "16,64").split(",") if v]:
    wb = bench.build_workload("A", 4096, sd, 0, y)
    S.set_kernel(_shim.KERNEL_LANE)
    o = S.solve_batch(wb["x0"], wb["xref"], obs_x=wb["obs_x"], obs_y=wb["obs_y"], u_init=wb["u_init"])
    conv = np.isin(o["status"], (0, 1)).mean()
    print(json.dumps({"lane_kernel": {"problems": int(wb["B"]), "kernel_ms": S.last_kernel_ms(), "solves_per_s": wb["B"] / S.last_kernel_ms() * 1e3,
                                      "converged_per_s": conv * wb["B"] / S.last_kernel_ms() * 1e3, "converged": float(conv),
                                      "stats": S.lane_kernel_stats()}}), flush=True)
S.close()
