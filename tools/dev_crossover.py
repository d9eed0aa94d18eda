"""Developer script (GPU box): batch size at which the lane kernel overtakes the warp kernel (variants B, C)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ros2_mpc_b200 import _shim, synth, load_params, make_params
y = load_params()
w = synth.robots_on_map(B=4096, seed=0)
pxf, puf = synth.straight_reference(w["x0"], w["goal"], 30)
for var in "BC":
    S = _shim.Solver(make_params(var, y))
    for B in (8192, 12288, 16384, 20480, 24576, 32768, 49152, 65536):
        rep = (B + 4095) // 4096
        x0 = np.tile(w["x0"], (rep, 1))[:B]
        xr = np.tile(w["goal"] if var == "B" else pxf, (rep, 1))[:B]
        kw = dict(uref=np.tile(puf, (rep, 1))[:B]) if var == "C" else {}
        res = {}
        for kind, nm in ((_shim.KERNEL_WARP, "warp"), (_shim.KERNEL_LANE, "lane")):
            S.set_kernel(kind)
            S.solve_batch(x0, xr, **kw)
            ms = []
            for _ in range(3):
                S.solve_batch(x0, xr, **kw); ms.append(S.last_kernel_ms())
            res[nm] = float(np.median(ms))
        print(var, B, "warp %.2f ms  lane %.2f ms  -> %s" % (res["warp"], res["lane"], "lane" if res["lane"] < res["warp"] else "warp"), flush=True)
    S.close()
