"""Producers of the solve's reference arguments, batched over robots.

  get_goal_for_mpc          ros2_mpc/scripts/point_follower_local_planner.py:16-30   -> final_state (variants A / B)
  get_headings              ros2_mpc/scripts/path_follower_local_planner.py:14-24    -> per-path heading / velocity / omega
  get_reference_trajectory  ros2_mpc/scripts/path_follower_local_planner.py:27-73    -> pf, puf (variant C)

The per-control-step functions run on the GPU (libb200mpc.so: b200mpc_goals_batch, b200mpc_reftraj_batch; one warp per
robot, bit-exact with the reference — see csrc/refgen_kernel.cuh), and so does the per-path preprocessing get_headings
(b200mpc_headings_batch, csrc/sensor_kernel.cuh).  The quirks of the reference are kept:
headings are taken modulo 2 pi only in get_goal_for_mpc, the tracking reference tiles goal[:3] (x, y and whatever the
caller stores third) within 0.5 m of the path end, and every array is padded with its last element."""
import numpy as np

from .obstacles import _default_solver


def get_headings(path_xy, dt, solver=None):
    """Drop-in for get_headings(path_xy, dt) (path_follower_local_planner.py:14-23) on the GPU (b200mpc_headings_batch):
    path_xy (K,2) -> heading (K,), velocity (K,), omega (K-1,); a batch (P,K,2) returns (P,K), (P,K), (P,K-1).
    Velocities are bit-exact with the reference, headings to the last ulp of atan2."""
    path_xy = np.asarray(path_xy, dtype=np.float64)
    single = path_xy.ndim == 2
    h, v, w = (solver or _default_solver()).headings_batch(path_xy[None] if single else path_xy, dt)
    return (h[0], v[0], w[0]) if single else (h, v, w)


def get_goals_batch(path_xy, path_heading, goal, pos, lookahead_dist_=0.5, solver=None):
    """Batched get_goal_for_mpc: goal (B,5), pos (B,>=2); path shared (K,2)/(K,) or per robot (B,K,2)/(B,K).
    Returns goal_pose (B,3) and the chosen path index (B,), -1 where the final goal was taken."""
    S = solver or _default_solver()
    ph = np.asarray(path_heading, dtype=np.float64)
    if ph.ndim and ph.shape[-1] == 1 and np.ndim(path_xy) == ph.ndim:  # the node keeps headings as (K,1)
        ph = ph[..., 0]
    return S.goals_batch(path_xy, ph, np.atleast_2d(goal), np.atleast_2d(pos), lookahead_dist_)


def get_goal_for_mpc(path_xy, path_heading, goal, pos, lookahead_dist_=0.5, solver=None):
    """Drop-in for get_goal_for_mpc(path_xy, path_heading, goal, pos, lookahead_dist_) — one robot."""
    out, _ = get_goals_batch(path_xy, path_heading, np.asarray(goal, dtype=np.float64)[None, :5],
                             np.asarray(pos, dtype=np.float64)[None, :2], lookahead_dist_, solver)
    return out[0]


def get_reference_trajectories_batch(x0, goal, path_xy, path_heading, path_velocity, path_omega, solver):
    """Batched get_reference_trajectory: x0 (B,3), goal (B,>=3); `solver` carries the horizon N (a variant-C handle).
    Returns pxf (B,3N), puf (B,2N) and the nearest path index (B,)."""
    return solver.reftraj_batch(path_xy, path_heading, path_velocity, path_omega, np.atleast_2d(x0), np.atleast_2d(goal))


def get_reference_trajectory(x0, goal, path_xy, path_heading, path_velocity, path_omega, mpc):
    """Drop-in for get_reference_trajectory(x0, goal, path_xy, path_heading, path_velocity, path_omega, mpc) — one
    robot; `mpc` is the ros2_mpc_b200 Mpc object (it owns the device handle).  Returns pxf (3N,1), puf (2N,1)."""
    ph = np.asarray(path_heading, dtype=np.float64).reshape(-1)
    pxf, puf, _ = get_reference_trajectories_batch(np.asarray(x0)[None, :3], np.asarray(goal)[None, :3], path_xy, ph,
                                                   path_velocity, path_omega, mpc._solver)
    return pxf.reshape(-1, 1), puf.reshape(-1, 1)
