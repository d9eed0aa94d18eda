"""`from ros2_mpc.planner.local_planner_tracking import Mpc` (scripts/path_follower_local_planner.py:5)."""
from ..mpc import MpcTracking as Mpc  # noqa: F401
