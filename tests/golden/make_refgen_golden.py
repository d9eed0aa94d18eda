"""Golden vectors for the reference producers, generated from the reference's own sources (build container only).

get_goal_for_mpc (scripts/point_follower_local_planner.py:16-30), get_headings and get_reference_trajectory
(scripts/path_follower_local_planner.py:14-73) live in scripts that import rclpy at module level, so the three
function definitions are cut out of the unmodified files with `ast` and executed as they are (numpy only).
Output: tests/golden/refgen_golden.npz."""
import ast
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/ros2_mpc/scripts"


def functions_of(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return [ns[n] for n in names]


class _Mpc:
    N = 30


def main():
    (get_goal_for_mpc,) = functions_of(os.path.join(REF, "point_follower_local_planner.py"), ["get_goal_for_mpc"])
    get_headings, get_reference_trajectory = functions_of(os.path.join(REF, "path_follower_local_planner.py"),
                                                          ["get_headings", "get_reference_trajectory"])
    rng = np.random.default_rng(777)
    paths, out = [], {}
    # paths: straight, curved, short (fewer points than the horizon), long
    t = np.linspace(0, 1, 120)
    paths.append(np.stack([4 * t - 2, 1.5 * np.sin(3 * t)], axis=1))
    paths.append(np.stack([np.linspace(-1, 2, 40), np.linspace(0.5, -1.0, 40)], axis=1))
    paths.append(np.round(np.cumsum(rng.normal(0.03, 0.04, (12, 2)), axis=0), 2))
    paths.append(np.round(np.cumsum(rng.normal(0.02, 0.05, (300, 2)), axis=0), 2))
    cases = []
    for pi, pxy in enumerate(paths):
        head, vel, om = get_headings(pxy, 0.2)
        out[f"path{pi}_xy"] = pxy; out[f"path{pi}_heading"] = head; out[f"path{pi}_velocity"] = vel; out[f"path{pi}_omega"] = om
        R = 64
        pos = np.empty((R, 3)); goal = np.empty((R, 5))
        gp = np.empty((R, 3)); pxf = np.empty((R, 90)); puf = np.empty((R, 60))
        for r in range(R):
            k = rng.integers(0, len(pxy))
            if r % 4 == 0:      # on a path point exactly (zero distance, ties)
                p = pxy[k].copy()
            elif r % 4 == 1:    # near the end of the path
                p = pxy[-1] + rng.normal(0, 0.3, 2)
            elif r % 4 == 2:    # on the 1 cm raster of the odometry subscriber
                p = np.round(pxy[k] + rng.normal(0, 0.4, 2), 2)
            else:               # far away
                p = pxy[k] + rng.normal(0, 3.0, 2)
            pos[r] = [p[0], p[1], np.round(rng.uniform(-3.14, 3.14), 2)]
            g = pxy[-1] + (rng.normal(0, 0.2, 2) if r % 3 else 0.0)
            goal[r] = [g[0], g[1], rng.normal(), rng.normal(), rng.uniform(-7, 7)]
            gp[r] = get_goal_for_mpc(pxy, head.reshape(-1, 1), goal[r], pos[r], 0.5)
            a, b = get_reference_trajectory(pos[r], goal[r], pxy, head, vel, om, _Mpc)
            pxf[r] = a.ravel(); puf[r] = b.ravel()
        out[f"path{pi}_pos"] = pos; out[f"path{pi}_goal"] = goal
        out[f"path{pi}_goal_pose"] = gp; out[f"path{pi}_pxf"] = pxf; out[f"path{pi}_puf"] = puf
        cases.append(pi)
    out["n_paths"] = np.array(len(paths))
    np.savez_compressed(os.path.join(HERE, "refgen_golden.npz"), **out)
    print("wrote refgen_golden.npz:", {k: v.shape for k, v in out.items() if k.startswith("path0")})


if __name__ == "__main__":
    main()
