import sys, os, json
sys.path.insert(0, "/root/repo")
import torch, bench
from ros2_mpc_b200 import load_params
y = load_params()
r = bench.variant_a_record(torch, torch.device("cuda", 0), y, 0, 36.5, 5, False, a_seeds=32)
c = r["config3"]
print(json.dumps({k: c[k] for k in ("one_launch_at_a_time", "double_buffered", "four_in_flight")}, indent=1))
