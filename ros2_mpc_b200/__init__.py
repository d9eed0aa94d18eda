"""ros2_mpc_b200 — B200-native batched nonlinear-MPC solve behind the ros2_mpc `Mpc.perform_mpc` interface.

The product path is hand-written sm_100a CUDA (csrc/b200mpc.cu -> libb200mpc.so) reached through a C ABI
(include/b200mpc.h) and this thin ctypes host layer.  No CPU fallback exists."""
from .mpc import MpcPointStabilization, MpcPointStabilizationLocal, MpcTracking, SolveError  # noqa: F401
from .params import load_params  # noqa: F401
from .variants import make_params  # noqa: F401
