"""GPU box: throughput of the configurations BASELINE.json names besides the bench line (config 4 = bench.py).

  config 1  single solve (latency: tools/dev_latency.py / bench.py "latency")
  config 3  4096 random poses / goals on map_carto, variants A (obstacle cost active) and B, warp kernel
  config 5  horizon sweep N = 10 / 25 / 50 / 100, halved control box, 160 distinct obstacle points, variant-A cost form,
            4096 problems each (max_iter 300 as in the parity test)
Prints one JSON line per case: kernel time from CUDA events (b200mpc_last_kernel_ms), converged fraction, iterations.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ros2_mpc_b200 import _shim, synth, load_params, make_params

y = load_params()
w = synth.robots_on_map(B=4096, seed=0)


def run(name, p, x0, xr, reps=3, **kw):
    S = _shim.Solver(p)
    S.solve_batch(x0, xr, **kw)
    ms, t = [], []
    for _ in range(reps):
        t0 = time.perf_counter()
        o = S.solve_batch(x0, xr, **kw)
        t.append(time.perf_counter() - t0)
        ms.append(S.last_kernel_ms())
    conv = np.isin(o["status"], (0, 1))
    B = x0.shape[0]
    print(json.dumps({"case": name, "problems": B, "kernel": "lane" if S.last_kernel_kind == _shim.KERNEL_LANE else "warp",
                      "kernel_ms": float(np.median(ms)), "e2e_ms": float(np.median(t) * 1e3),
                      "converged_fraction": float(conv.mean()),
                      "converged_solves_per_s": float(conv.sum() / (np.median(ms) * 1e-3)),
                      "iters_mean": float(o["iters"].mean()), "iters_max": int(o["iters"].max()),
                      "status_counts": {int(k): int(v) for k, v in zip(*np.unique(o["status"], return_counts=True))}}), flush=True)
    S.close()


run("config3 variant B, 4096 problems", make_params("B", y), w["x0"], w["goal"])
run("config3 variant A, 4096 problems", make_params("A", y), w["x0"], w["goal"], obs_x=w["obs_x"], obs_y=w["obs_y"])
pxf, puf = synth.straight_reference(w["x0"], w["goal"], y["N"])
run("config3 variant C (tracking), 4096 problems", make_params("C", y), w["x0"], pxf, uref=puf)
# config 5 as SURVEY 8d states it (annulus from 0.3 m, IPOPT's max_iter 3000) and round 1's easier field (0.6 m, max_iter 300)
for field, r_in, max_iter in (("stated", 0.3, 3000), ("easier", 0.6, 300)):
    ox, oy = synth.dense_obstacle_field(w["x0"], seed=2, r_in=r_in)
    for N in (10, 25, 50, 100):
        p = make_params("A", y, N=N, u_lo=[-0.025, -0.1], u_hi=[0.075, 0.1], max_iter=max_iter)
        run(f"config5 ({field} field) N={N}, dense obstacle field, tightened bounds, 4096 problems", p, w["x0"], w["goal"], reps=2, obs_x=ox, obs_y=oy)
# large batches of the three variants through the automatic kernel choice
for var, rep in (("B", 64), ("C", 64)):
    x0 = np.tile(w["x0"], (rep, 1))
    if var == "C":
        run(f"variant C, {4096 * rep} problems", make_params("C", y), x0, np.tile(pxf, (rep, 1)), uref=np.tile(puf, (rep, 1)), reps=2)
    else:
        run(f"variant B, {4096 * rep} problems", make_params("B", y), x0, np.tile(w["goal"], (rep, 1)), reps=2)

# closed loop on the device (fleet.py: goals -> solve -> limiter / goal logic / plant / next measurement), no host round trips
import torch
from ros2_mpc_b200 import references as rf
from ros2_mpc_b200.fleet import FleetPointStabilization, FleetObstacleAvoidance
for Bf, steps in ((4096, 20), (262144, 5)):
    rng = np.random.default_rng(1)
    tt = np.linspace(0, 1, 60)
    path = np.stack([1.6 * tt, 0.4 * np.sin(2.5 * tt)], axis=1)
    head, _, _ = rf.get_headings(path, y["dt"])
    start = np.c_[rng.normal(0, 0.3, Bf), rng.normal(0, 0.3, Bf), rng.uniform(-0.6, 0.6, Bf)]
    goal = np.tile(np.r_[path[-1], 0.0, 0.0, head[-1]], (Bf, 1))
    fl = FleetPointStabilization(start, goal, path, head, params=y, warm_start=False, quantise=True)
    fl.step(2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fl.step(steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    snap = fl.snapshot()
    print(json.dumps({"case": f"closed loop on the device, {Bf} robots x {steps} control steps (cold start each step, as the reference)",
                      "robots": Bf, "control_steps": steps, "ms_per_control_step": ms / steps,
                      "robot_steps_per_s": Bf * steps / (ms * 1e-3),
                      "solve_success_fraction_last_step": float(np.isin(snap["status"], (0, 1)).mean()),
                      "iters_mean_last_step": float(snap["iters"].mean())}), flush=True)
    fl.close()

# the same for the obstacle-active variant A on the shared map (ray-cast -> obstacle list -> goal -> solve -> control step)
m = synth.load_map()
for Bf, steps in ((4096, 10),):
    wr = synth.robots_on_map(B=Bf, seed=7, m=m)
    K = 24
    tt = np.linspace(0, 1, K)[None, :, None]
    path = wr["x0"][:, None, :2] * (1 - tt) + wr["goal"][:, None, :2] * tt
    seg = np.diff(path, axis=1)
    hd = np.arctan2(seg[..., 1], seg[..., 0]); hd = np.concatenate([hd, hd[:, -1:]], axis=1)
    goal = np.c_[wr["goal"][:, :2], np.zeros((Bf, 2)), wr["goal"][:, 2]]
    fl = FleetObstacleAvoidance(wr["x0"], goal, path, hd, m, params=y)
    fl.step(2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fl.step(steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    snap = fl.snapshot()
    print(json.dumps({"case": f"closed loop on the device with the obstacle cost (variant A), {Bf} robots x {steps} control steps",
                      "robots": Bf, "control_steps": steps, "ms_per_control_step": ms / steps,
                      "robot_steps_per_s": Bf * steps / (ms * 1e-3),
                      "solve_success_fraction_last_step": float(np.isin(snap["status"], (0, 1)).mean()),
                      "iters_mean_last_step": float(snap["iters"].mean())}), flush=True)
    fl.close()
