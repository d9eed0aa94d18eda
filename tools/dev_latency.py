"""Developer script (GPU box): single-solve latency split — wall time of the host-buffer call vs kernel time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ros2_mpc_b200 import _shim, synth, load_params, make_params

y = load_params()
w = synth.robots_on_map(B=256, seed=0)
for var in os.environ.get("VARS", "B"):
    S = _shim.Solver(make_params(var, y))
    kw = (lambda i: dict(obs_x=w["obs_x"][i], obs_y=w["obs_y"][i])) if var == "A" else (lambda i: {})
    S.solve_batch(w["x0"][:1], w["goal"][:1], **kw(0))
    tw, tk, it = [], [], []
    for i in range(256):
        t = time.perf_counter()
        o = S.solve_batch(w["x0"][i:i + 1], w["goal"][i:i + 1], **kw(i))
        tw.append(time.perf_counter() - t)
        tk.append(S.last_kernel_ms())
        it.append(int(o["iters"][0]))
    tw, tk, it = np.array(tw) * 1e3, np.array(tk), np.array(it)
    print(var, "wall p50 %.3f p99 %.3f ms | kernel p50 %.3f p99 %.3f ms | iters mean %.1f | kernel us/iter %.1f | overhead p50 %.3f ms" % (
        np.median(tw), np.quantile(tw, 0.99), np.median(tk), np.quantile(tk, 0.99), it.mean(), 1e3 * (tk / np.maximum(it, 1)).mean(),
        np.median(tw - tk)), flush=True)
    S.close()
