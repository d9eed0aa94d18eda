"""GPU box: one 1/8 shard of bench.py's config-4 batch (131 072 problems, device-resident) under different hand-over
thresholds of the lane kernel — what bounds the strong-scaling record at 8 GPUs."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ros2_mpc_b200 import _shim, load_params, make_params

y = load_params()
world = int(os.environ.get("SHARDS", "8"))
wl = bench.build_workload("B", 4096, 256, 0, y, seed_lo=0, seed_hi=256 // world)
p = make_params("B", y)
B, N = wl["B"], y["N"]
dev = torch.device("cuda", 0)
d = {k: (None if wl[k] is None else torch.from_numpy(np.ascontiguousarray(wl[k])).to(dev)) for k in ("x0", "xref", "uref", "u_init")}
dX = torch.empty((B, N + 1, 3), dtype=torch.float64, device=dev); dU = torch.empty((B, N, 2), dtype=torch.float64, device=dev)
dc = torch.empty(B, dtype=torch.float64, device=dev); ds = torch.empty(B, dtype=torch.int32, device=dev)
di = torch.empty(B, dtype=torch.int32, device=dev); dl = torch.empty(B, dtype=torch.int32, device=dev)
ptr = lambda t: 0 if t is None else t.data_ptr()
stream = torch.cuda.current_stream()
ref_status = None
for hi, ht in itertools.product(os.environ.get("HI", "60,40,28").split(","), os.environ.get("HT", "28,20,14,8").split(",")):
    if int(ht) > int(hi):
        continue
    os.environ["B200MPC_HAND_ITER"] = hi; os.environ["B200MPC_HAND_ITER_TAIL"] = ht
    S = _shim.Solver(p, device=0)
    def step():
        S.solve_batch_device(B, ptr(d["x0"]), ptr(d["xref"]), ptr(d["uref"]), 0, 0, 0, ptr(d["u_init"]), dX.data_ptr(), dU.data_ptr(),
                             dc.data_ptr(), ds.data_ptr(), di.data_ptr(), dl.data_ptr(), stream=stream.cuda_stream)
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5): step()
    e1.record(stream); torch.cuda.synchronize()
    st = ds.cpu().numpy(); it = di.cpu().numpy()
    if ref_status is None: ref_status, ref_it = st.copy(), it.copy()
    print(f"B={B} hand_iter={hi} tail={ht}: {e0.elapsed_time(e1)/5:.2f} ms/step, kernel {S.last_kernel_ms():.2f} ms, conv {np.isin(st,(0,1)).mean():.5f}, "
          f"status==first {np.array_equal(st, ref_status)}, iters==first {np.array_equal(it, ref_it)}, iters mean {it.mean():.2f} max {it.max()}, "
          f">28: {(it>28).mean():.4f} >60: {(it>60).mean():.5f}", flush=True)
    S.close()
