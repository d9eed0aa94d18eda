"""`from ros2_mpc.mpc_point_stabilization import Mpc` — the standalone obstacle-active variant."""
from .mpc import MpcPointStabilization as Mpc  # noqa: F401
