"""The three NLP templates of the reference as b200mpc_params.

  'A'  ros2_mpc/mpc_point_stabilization.py:9-149                 RK4, fixed goal, exp(c/s) obstacle cost ACTIVE
  'B'  ros2_mpc/planner/local_planner_point_stabilization.py:11-178   RK4, fixed goal, obstacle cost built but dropped
  'C'  ros2_mpc/planner/local_planner_tracking.py:11-178         Euler, time-varying X/U reference, obstacle cost zeroed

The reference hard-codes several constants instead of reading params.yaml (R and the control box in A/B, Q in A)
and swaps cost_factor / reverse_factor at the call sites; those quirks are reproduced here, not "fixed".
"""
from . import _shim
from .params import obstacle_slots


def make_params(variant, y, N=None, obstacles=None, **overrides):
    """b200mpc_params for `variant` from the params.yaml dict `y`.

    obstacles=True on variant 'B' enables its (built but never minimised) gauss obstacle cost — the evident
    intent of the author, flagged non-reference.  `overrides` may set any struct field (u_lo, u_hi, tol, ...)."""
    p = _shim.default_params()
    p.N = int(y["N"] if N is None else N)
    p.M = obstacle_slots(y)                      # local_planner_point_stabilization.py:155-156
    p.dt = float(y["dt"])
    p.obs_r = float(y["inflation_radius"])
    p.obs_form, p.obs_c, p.obs_k0, p.obs_k1 = _shim.OBS_NONE, 0.0, 0, -1
    if variant == "A":
        p.integrator, p.ref_kind = _shim.RK4, _shim.REF_GOAL
        p.Q[:] = [0.00005, 0.05, 0.05]           # mpc_point_stabilization.py:87-90
        p.R[:] = [0.01, 0.01]                    # :92-93
        p.kappa = float(y["cost_factor"])        # :35 -> :85,99 (named reverse_factor there)
        p.obs_form, p.obs_c = _shim.OBS_EXPLOG, float(y["reverse_factor"])  # :33 -> :46-53
        p.obs_k0, p.obs_k1 = 0, p.N              # range(N+1) :48
        p.u_lo[:] = [-0.2, -0.1]
        p.u_hi[:] = [0.2, 0.1]                   # :82-83
    elif variant == "B":
        p.integrator, p.ref_kind = _shim.RK4, _shim.REF_GOAL
        p.Q[:] = [float(v) for v in y["Q"]]      # local_planner_point_stabilization.py:106-109
        p.R[:] = [0.5, 0.5]                      # :111-112 (params R ignored)
        p.kappa = float(y["cost_factor"])        # :47 -> :104,125
        p.u_lo[:] = [-0.05, -0.2]
        p.u_hi[:] = [0.15, 0.2]                  # :101-102
        if obstacles:
            p.obs_form, p.obs_c = _shim.OBS_GAUSS, float(y["reverse_factor"])  # :43-45 -> :60-67
            p.obs_k0, p.obs_k1 = 0, p.N - 1      # range(N) :62
    elif variant == "C":
        p.integrator, p.ref_kind = _shim.EULER, _shim.REF_TRAJ
        p.Q[:] = [float(v) for v in y["Q"]]      # local_planner_tracking.py:108-111
        p.R[:] = [float(v) for v in y["R"]]      # :113-115
        p.kappa = float(y["reverse_factor"])     # :124
        p.u_lo[:] = [-0.1, -0.2]
        p.u_hi[:] = [0.2, 0.2]                   # :94-95
    else:
        raise ValueError(f"unknown variant {variant!r} (expected 'A', 'B' or 'C')")
    for k, v in overrides.items():
        if k in ("u_lo", "u_hi", "Q", "R"):
            getattr(p, k)[:] = [float(t) for t in v]
        else:
            setattr(p, k, v)
    return p
