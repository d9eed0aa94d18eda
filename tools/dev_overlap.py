"""Developer script (GPU box): do consecutive batched solves overlap their end-of-batch tails when they alternate
between two handles / streams?  (A persistent lane-kernel CTA of the next solve starts on an SM as soon as the previous
solve's CTA there has exited.)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ros2_mpc_b200 import _shim, load_params

y = load_params()
wl = bench.build_workload("B", 4096, 256, 0, y)
p, B, N = wl["p"], wl["B"], y["N"]
dev = torch.device("cuda", 0)
t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d = {k: t(wl[k]) for k in ("x0", "xref", "u_init")}
ptr = lambda a: 0 if a is None else a.data_ptr()


def mk():
    S = _shim.Solver(p)
    o = dict(X=torch.empty((B, N + 1, 3), dtype=torch.float64, device=dev), U=torch.empty((B, N, 2), dtype=torch.float64, device=dev),
             c=torch.empty(B, dtype=torch.float64, device=dev), s=torch.empty(B, dtype=torch.int32, device=dev),
             i=torch.empty(B, dtype=torch.int32, device=dev), l=torch.empty(B, dtype=torch.int32, device=dev))
    return S, o, torch.cuda.Stream()


def step(S, o, st):
    S.solve_batch_device(B, ptr(d["x0"]), ptr(d["xref"]), 0, 0, 0, 0, ptr(d["u_init"]), ptr(o["X"]), ptr(o["U"]), ptr(o["c"]),
                         ptr(o["s"]), ptr(o["i"]), ptr(o["l"]), stream=st.cuda_stream)


hs = [mk(), mk()]
for K, nh in ((6, 1), (6, 2), (6, 1), (6, 2)):
    for h in hs:
        step(*h)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(K):
        step(*hs[k % nh])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("handles %d: %d steps in %.1f ms -> %.3f M solves/s" % (nh, K, dt * 1e3, K * B / dt / 1e6), flush=True)
