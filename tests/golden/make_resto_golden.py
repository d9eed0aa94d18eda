"""Fixture for the restoration stand-in's second stage (deviation 3).

Four config-3 problems (map_carto, PCG64 seed 0, indices below) on which the first stage finds no acceptable point — the
roll-out of the slacks passes within centimetres of an obstacle point, exp(c/s) overflows — and the solve used to end
Restoration_Failed at a point with a cost of 2.5e8 ... 4.6e24.  For two of them scipy's SLSQP, run on the same
multiple-shooting NLP from the reference's cold start (tools/ipopt_fidelity_a.py), reaches a KKT point; its cost is stored
as the third-party answer the oracle (and the CUDA kernels) must now reproduce.
Run from the repository root:  python tests/golden/make_resto_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    from ros2_mpc_b200 import load_params, synth
    from ipopt_fidelity_a import work
    params = load_params()
    w = synth.robots_on_map(B=4096, seed=0, params=params)
    idx = np.array([1083, 482, 1750, 1264])
    slsqp = np.full(len(idx), np.nan)
    for n, i in enumerate(idx):
        rec = work((i, w["x0"][i], w["goal"][i], w["obs_x"][i], w["obs_y"][i], params))
        sc = rec.get("slsqp_cold", {})
        if sc.get("kkt_point"):
            slsqp[n] = sc["cost"]
        print(i, rec["oracle_status"], rec["oracle_cost"], sc)
    np.savez(os.path.join(ROOT, "tests", "golden", "resto_golden.npz"), index=idx, x0=w["x0"][idx], goal=w["goal"][idx],
             obs_x=w["obs_x"][idx], obs_y=w["obs_y"][idx], slsqp_cost=slsqp)


if __name__ == "__main__":
    main()
