"""GPU box: the costmap kernels' bench lines alone (bench.costmap_lines)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ros2_mpc_b200 import _shim, load_params, make_params
y = load_params()
S = _shim.Solver(make_params("B", y))
print(json.dumps(bench.costmap_lines(S, torch, torch.device("cuda", 0), y, 6548.2), indent=1))
S.close()
