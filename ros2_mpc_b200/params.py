"""params.yaml loader with the reference's lookup order.

The reference's Mpc() takes no arguments and finds share/ros2_mpc/config/params.yaml through
ament_index_python (local_planner_point_stabilization.py:13-15).  Here: explicit path > ament index >
$ROS2_MPC_SHARE > the packaged default (same keys and values)."""
import os

import yaml

REQUIRED_KEYS = ("dt", "N", "Q", "R", "resolution", "cost_factor", "costmap_size", "inflation_radius",
                 "reverse_factor", "look_ahead_distance", "goal_threshold")

_PACKAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config", "params.yaml")


def find_params_file(params_path=None):
    if params_path is not None:
        return params_path
    try:
        from ament_index_python.packages import get_package_share_directory  # noqa: PLC0415
        cand = os.path.join(get_package_share_directory("ros2_mpc"), "config", "params.yaml")
        if os.path.exists(cand):
            return cand
    except Exception:  # ament not installed / package not registered
        pass
    share = os.environ.get("ROS2_MPC_SHARE")
    if share:
        cand = os.path.join(share, "config", "params.yaml")
        if os.path.exists(cand):
            return cand
    return _PACKAGED


def load_params(params_path=None):
    path = find_params_file(params_path)
    with open(path, "r") as f:
        params = yaml.safe_load(f)
    missing = [k for k in REQUIRED_KEYS if k not in params]
    if missing:
        raise KeyError(f"{path}: missing keys {missing}")
    return params


def obstacle_slots(params):
    """Length of the obstacle parameter vectors, int((costmap_size*2)/resolution)*2
    (local_planner_point_stabilization.py:155-156) — 160 for the shipped file."""
    return int((params["costmap_size"] * 2) / params["resolution"]) * 2
