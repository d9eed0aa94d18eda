"""Profiling driver (GPU box): the obstacle-list kernel on device-resident scans (as bench.py's obstacle_builder line)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ros2_mpc_b200 import _shim, load_params, make_params
import bench

y = load_params()
S = _shim.Solver(make_params("B", y))
dev = torch.device("cuda", 0)
for _ in range(2):
    r = bench.obstacle_builder_line(S, torch, dev, y, reps=int(os.environ.get("REPS", "5")))
    print({k: r[k] for k in ("ms_per_launch", "robots_per_s", "achieved_gbs", "matches_cpu_mirror")}, flush=True)
S.close()
