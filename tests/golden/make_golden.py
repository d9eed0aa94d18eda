"""Generate the golden fixtures under tests/golden/ from the reference itself (build container only).

The GPU box has no /root/reference, so everything derived from it is committed as small .npz files together with
this script.  Three families:

 1. obstacles_golden.npz — the reference's own numba helpers (ros2_mpc/utils/utils.py, unmodified, loaded by file
    path) and the get_obstacles body of scripts/point_follower_local_planner.py:88-118 run on seeded laser scans
    (NaN / +-inf beams, truncation-toward-zero cases): occupancy grids and obstacle lists.
 2. nlp_golden.npz — the reference's three Mpc classes (unmodified sources) executed against the tracing casadi
    stand-in in tests/golden/fake_casadi: objective and constraint values of the NLP they pose at seeded random
    points.  This pins the oracle's / kernel's NLP restatement to the reference's code, independent of any solver.
 3. solve_golden.npz — Mpc.perform_mpc of the reference run end to end with scipy SLSQP substituted for IPOPT
    (CasADi/IPOPT are not installable offline): returned controls / trajectories / costs for config-1 problems.
    These pin the *optimum*; IPOPT's own iterates remain unpinned.
"""
import importlib.util
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------------------------
def make_obstacles_golden():
    ru = load_by_path("ref_utils", os.path.join(REF, "ros2_mpc/utils/utils.py"))

    def ref_get_obstacles(scan_data, angles, size, resolution, pos, ori, obstacles_x, obstacles_y):
        # body of scripts/point_follower_local_planner.py:88-118 driven through the reference's utils
        # (the script itself imports rclpy at module level and cannot be imported here)
        occ_grid = 1 - (ru.convert_laser_scan_to_occupancy_grid(scan_data, angles, resolution, size * 2) / 100)
        occ_grid = np.rot90(occ_grid, k=2)
        x, y = ru.convert_to_map_coordinates(occ_grid=occ_grid, map_resolution=resolution)
        idx = np.where(occ_grid == 0)
        obs = np.array([x[idx], y[idx]])
        rot = ru.rotate_coordinates(obs, ori[2])
        rot[0, :] += pos[0]
        rot[1, :] += pos[1]
        x_obs, y_obs = rot[0, :], rot[1, :]
        n = len(x_obs)
        try:
            xa = obstacles_x * x_obs[0]
            xa[:min(n, len(xa))] = x_obs[:len(xa)]
            ya = obstacles_y * y_obs[0]
            ya[:min(n, len(ya))] = y_obs[:len(ya)]
        except IndexError:
            xa = obstacles_x * 100
            ya = obstacles_y * 100
        return xa, ya, n

    rng = np.random.default_rng(20240)
    S = 48
    n = 360
    scans = np.empty((S, n)); angles = np.empty((S, 2)); pos = np.empty((S, 2)); yaw = np.empty(S)
    grids = np.empty((S, 80, 80), dtype=np.uint8)
    ox = np.empty((S, 160)); oy = np.empty((S, 160)); cnt = np.empty(S, dtype=np.int64)
    for i in range(S):
        kind = i % 6
        if kind == 0:    # room: walls at ~1.5 m
            sc = 1.5 / np.maximum(np.abs(np.cos(np.arange(n) * 6.28 / n)), np.abs(np.sin(np.arange(n) * 6.28 / n)))
            sc += rng.normal(0, 0.01, n)
        elif kind == 1:  # random ranges on a 1 cm raster (exact cell-boundary cases)
            sc = np.round(rng.uniform(0.12, 3.5, n), 2)
        elif kind == 2:  # far away: nothing inside the grid
            sc = rng.uniform(3.0, 3.5, n)
        elif kind == 3:  # sparse hits
            sc = np.full(n, 3.5); hit = rng.random(n) < 0.15; sc[hit] = rng.uniform(0.2, 1.9, hit.sum())
        elif kind == 4:  # close to the negative edge: coordinates in (-0.05, 0) truncate to cell 0
            sc = rng.uniform(1.9, 2.06, n)
        else:
            sc = rng.uniform(0.12, 2.8, n)
        if kind in (1, 5) and i % 12 >= 6:
            # thin the scan out so that most cases stay within the reference's 160-slot limit
            sc[rng.random(n) < 0.6] = 3.5
        m = rng.random(n)
        if i % 2 == 1:
            sc[m < 0.04] = np.inf
            sc[(m >= 0.04) & (m < 0.07)] = np.nan
        if i % 8 == 7:
            sc[m > 0.97] = -np.inf
        if i == 2:
            sc[:] = np.inf   # every beam out of range -> the robot's own cell
        a = np.array([0.0, 6.28]) if i % 3 else np.array([-3.14159, 3.14159])
        scans[i], angles[i] = sc, a
        pos[i] = np.round(rng.uniform(-3, 3, 2), 2); yaw[i] = np.round(rng.uniform(-3.14, 3.14), 2)
        g = ru.convert_laser_scan_to_occupancy_grid(sc.copy(), a, 0.05, 4.0)
        grids[i] = (g == 100).astype(np.uint8)
        xa, ya, c = ref_get_obstacles(sc.copy(), a, 2.0, 0.05, pos[i], np.array([0.0, 0.0, yaw[i]]), np.ones(160), np.ones(160))
        ox[i], oy[i], cnt[i] = xa, ya, c
    np.savez_compressed(os.path.join(HERE, "obstacles_golden.npz"), scans=scans, angles=angles, pos=pos, yaw=yaw,
                        grid_bits=np.packbits(grids), obs_x=ox, obs_y=oy, count=cnt)
    print("obstacles_golden: cells per scan", cnt.min(), cnt.max(), "overflow scans", int((cnt > 160).sum()))


# ------------------------------------------------------------------------------------------------------------
def build_reference_mpcs():
    sys.path.insert(0, os.path.join(HERE, "fake_casadi"))
    import casadi  # noqa: F401  (the stand-in)
    mods = {
        "A": load_by_path("ref_mpc_a", os.path.join(REF, "ros2_mpc/mpc_point_stabilization.py")),
        "B": load_by_path("ref_mpc_b", os.path.join(REF, "ros2_mpc/planner/local_planner_point_stabilization.py")),
        "C": load_by_path("ref_mpc_c", os.path.join(REF, "ros2_mpc/planner/local_planner_tracking.py")),
    }
    return {k: m.Mpc() for k, m in mods.items()}


def opti_vector(N, X, U):
    """(X (N+1,3), U (N,2)) -> the reference's Opti decision vector: 5 get_system_function symbols, X (3,N+1)
    column-major, U (2,N) column-major."""
    return np.concatenate([np.zeros(5), X.reshape(-1), U.reshape(-1)])


def make_nlp_golden(mpcs):
    rng = np.random.default_rng(777)
    out = {}
    N = 30
    npts = 6
    for var, mpc in mpcs.items():
        opti = mpc.opti
        assert len(opti.vars) == 5 + 3 * (N + 1) + 2 * N
        x0 = np.array([0.4, -0.3, 0.9]); goal = np.array([1.1, 0.2, 0.5])
        ang = rng.uniform(0, 2 * np.pi, 160); rr = rng.uniform(2.5, 4.0, 160)
        obs_x = x0[0] + rr * np.cos(ang); obs_y = x0[1] + rr * np.sin(ang)
        pf = rng.normal(0, 1, 3 * N); puf = rng.uniform(-0.1, 0.1, 2 * N)
        if var == "C":
            opti.set_value(mpc.P_X, np.concatenate([x0, pf])); opti.set_value(mpc.P_U, puf)
        else:
            opti.set_value(mpc.P, np.concatenate([x0, goal]))
        opti.set_value(mpc.obstacles_x, obs_x); opti.set_value(mpc.obstacles_y, obs_y)
        Xs = rng.normal(0, 0.4, (npts, N + 1, 3)) + x0
        Us = rng.uniform(-0.2, 0.2, (npts, N, 2))
        Xs[:, 0, :] = x0
        f = np.empty(npts); c = np.empty((npts, N, 3)); h = np.empty((npts, 2 * N))
        for i in range(npts):
            z = opti_vector(N, Xs[i], Us[i])
            # column 0 of X is overwritten by the parameter in the reference: perturb it to prove it is unused
            z[5:8] = rng.normal(0, 5, 3)
            f[i] = opti.eval_f(z)
            c[i] = opti.eval_g(z).reshape(N, 3)
            hv, lo, hi = opti.eval_h(z)
            h[i] = hv
        out.update({f"{var}_x0": x0, f"{var}_goal": goal, f"{var}_pf": pf, f"{var}_puf": puf, f"{var}_obs_x": obs_x,
                    f"{var}_obs_y": obs_y, f"{var}_X": Xs, f"{var}_U": Us, f"{var}_f": f, f"{var}_c": c, f"{var}_h": h,
                    f"{var}_h_lo": lo, f"{var}_h_hi": hi})
        print(var, "nlp golden f", f[:3])
    np.savez_compressed(os.path.join(HERE, "nlp_golden.npz"), **out)


def make_gauss_golden(mpcs):
    """Variant B's obstacle cost (local_planner_point_stabilization.py:60-67): built by the reference, then dropped from its
    objective (:127).  The reference's own method is called on the traced problem and its expression evaluated at seeded
    points with obstacle lists close to the trajectory; appended to nlp_golden.npz as Bobs_*."""
    from casadi import evaluate  # the stand-in's evaluator
    mpc = mpcs["B"]
    opti = mpc.opti
    rng = np.random.default_rng(4242)
    N = 30
    expr = mpc.define_obstacles_cost_function(cost_factor=5.0)   # params["reverse_factor"], as :43-45 passes it
    x0 = np.array([0.4, -0.3, 0.9])
    ang = rng.uniform(0, 2 * np.pi, 160); rr = rng.uniform(0.1, 0.8, 160)
    obs_x = x0[0] + rr * np.cos(ang); obs_y = x0[1] + rr * np.sin(ang)
    opti.set_value(mpc.P, np.concatenate([x0, np.array([1.1, 0.2, 0.5])]))
    opti.set_value(mpc.obstacles_x, obs_x); opti.set_value(mpc.obstacles_y, obs_y)
    npts = 6
    Xs = rng.normal(0, 0.3, (npts, N + 1, 3)) + x0
    Xs[:, 0, :] = x0
    Us = rng.uniform(-0.1, 0.1, (npts, N, 2))
    f = np.empty(npts)
    for i in range(npts):
        z = opti_vector(N, Xs[i], Us[i])
        f[i] = evaluate(expr.nodes(), opti._env(z))[0]
    g = dict(np.load(os.path.join(HERE, "nlp_golden.npz")))
    g.update(Bobs_x0=x0, Bobs_obs_x=obs_x, Bobs_obs_y=obs_y, Bobs_X=Xs, Bobs_U=Us, Bobs_f=f)
    np.savez_compressed(os.path.join(HERE, "nlp_golden.npz"), **g)
    print("gauss obstacle cost golden", f[:3])


def make_solve_golden(mpcs):
    out = {}
    N = 30
    u0 = np.zeros((2, N))
    sent = np.full(160, 100.0)
    # variant B: the defaults of perform_mpc and the BASELINE.md anchor problem
    for tag, x0, goal in (("B1", np.array([0.0, 0.0, 0.0]), np.array([1.0, 1.0, 0.0])),
                          ("B2", np.array([0.0, 0.0, 0.0]), np.array([10.0, 10.0, 0.0])),
                          ("B3", np.array([-2.96, 2.31, 0.5]), np.array([-2.4, 2.9, 1.2]))):
        t = time.time()
        mpc = mpcs["B"]
        u = mpc.perform_mpc(u0, x0, goal, sent, sent)
        # the full solution, re-solving through the same object the way perform_mpc does
        sol = mpc.opti.solve()
        out.update({f"{tag}_x0": x0, f"{tag}_goal": goal, f"{tag}_u0": np.asarray(u), f"{tag}_X": sol.value(mpc.X),
                    f"{tag}_U": sol.value(mpc.U), f"{tag}_cost": sol.value(mpc.opti.objective)})
        print(tag, "cost", out[f"{tag}_cost"], "u0", u, "%.1fs" % (time.time() - t))
    # variant A: config 1 (sentinel obstacles) and a wall 1 m to the side
    wall_x = np.linspace(-1.0, 2.0, 160); wall_y = np.full(160, 1.0)
    for tag, x0, goal, ox, oy in (("A1", np.array([0.0, 0.0, 0.0]), np.array([10.0, 10.0, 0.0]), sent, sent),
                                  ("A2", np.array([0.0, 0.0, 0.0]), np.array([1.5, 0.3, 0.0]), wall_x, wall_y)):
        t = time.time()
        mpc = mpcs["A"]
        x_opt, u_opt = mpc.perform_mpc(u0, x0, goal, ox, oy)
        z = np.concatenate([np.zeros(5), np.asarray(x_opt).reshape(-1, order="F"), np.asarray(u_opt).reshape(-1, order="F")])
        out.update({f"{tag}_x0": x0, f"{tag}_goal": goal, f"{tag}_obs_x": ox, f"{tag}_obs_y": oy, f"{tag}_X": x_opt,
                    f"{tag}_U": u_opt, f"{tag}_cost": mpc.opti.eval_f(z)})
        print(tag, "cost", out[f"{tag}_cost"], "u0", u_opt[:, 0], "%.1fs" % (time.time() - t))
    # variant C: straight reference
    mpc = mpcs["C"]
    x0 = np.array([0.1, -0.2, 0.3])
    t_ = (np.arange(1, N + 1) / N)[:, None]
    goal = np.array([1.0, 0.6, 0.0])
    head = np.arctan2(goal[1] - x0[1], goal[0] - x0[0])
    pf = np.concatenate([x0[:2] * (1 - t_) + goal[:2] * t_, np.full((N, 1), head)], axis=1).reshape(-1, 1)
    puf = np.tile([0.1, 0.0], N).reshape(-1, 1)
    x_opt, u0c = mpc.perform_mpc(u0, x0, pf, puf, sent, sent)
    sol = mpc.opti.solve()
    out.update({"C1_x0": x0, "C1_pf": pf.ravel(), "C1_puf": puf.ravel(), "C1_X": x_opt, "C1_u0": np.asarray(u0c),
                "C1_U": sol.value(mpc.U), "C1_cost": sol.value(mpc.opti.objective)})
    print("C1 cost", out["C1_cost"], "u0", u0c)
    np.savez_compressed(os.path.join(HERE, "solve_golden.npz"), **out)


if __name__ == "__main__":
    make_obstacles_golden()
    mpcs = build_reference_mpcs()
    make_nlp_golden(mpcs)
    make_gauss_golden(mpcs)
    make_solve_golden(mpcs)
