"""Same import path shape as the reference: `from ros2_mpc.planner.local_planner_point_stabilization import Mpc`
(scripts/point_follower_local_planner.py:7) becomes `from ros2_mpc_b200.planner.local_planner_point_stabilization
import Mpc`."""
from ..mpc import MpcPointStabilizationLocal as Mpc  # noqa: F401
