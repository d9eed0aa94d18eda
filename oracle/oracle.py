"""ctypes binding of oracle/libmpc_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product package (ros2_mpc_b200) never does.  Solver parity is unpinned (no CasADi/IPOPT offline);
the NLP restatement is pinned against the reference sources by tests/golden/make_golden.py.

Variant tables follow SURVEY.md App. A:
  A  /root/reference/ros2_mpc/mpc_point_stabilization.py:9-149
  B  /root/reference/ros2_mpc/planner/local_planner_point_stabilization.py:11-178
  C  /root/reference/ros2_mpc/planner/local_planner_tracking.py:11-178
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

RK4, EULER = 0, 1
OBS_NONE, OBS_GAUSS, OBS_EXPLOG = 0, 1, 2
REF_GOAL, REF_TRAJ = 0, 1

STATUS_NAMES = {
    0: "Solve_Succeeded", 1: "Solved_To_Acceptable_Level", 3: "Search_Direction_Becomes_Too_Small",
    -1: "Maximum_Iterations_Exceeded", -2: "Restoration_Failed", -3: "Error_In_Step_Computation",
    -13: "Invalid_Number_Detected",
}


class OrcParams(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("M", C.c_int), ("dt", C.c_double), ("integrator", C.c_int),
        ("Q", C.c_double * 3), ("R", C.c_double * 2), ("kappa", C.c_double), ("ref_kind", C.c_int),
        ("obs_form", C.c_int), ("obs_c", C.c_double), ("obs_r", C.c_double),
        ("obs_k0", C.c_int), ("obs_k1", C.c_int), ("u_lo", C.c_double * 2), ("u_hi", C.c_double * 2),
        ("tol", C.c_double), ("max_iter", C.c_int), ("acceptable_tol", C.c_double),
        ("acceptable_iter", C.c_int), ("mu_init", C.c_double), ("max_soc", C.c_int),
        ("linear_solver", C.c_int),
    ]


class OrcStats(C.Structure):
    _fields_ = [("iters", C.c_int), ("ls_extra", C.c_int), ("n_soc", C.c_int), ("n_reg", C.c_int),
                ("n_resto", C.c_int), ("mu", C.c_double), ("err", C.c_double), ("obj_scale", C.c_double)]


def build(force=False):
    so = os.path.join(_HERE, "libmpc_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("mpc_oracle.c", "mpc_oracle.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "clean", "all"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libmpc_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.orc_default_options.argtypes = [C.POINTER(OrcParams)]
        L.orc_solve.argtypes = [C.POINTER(OrcParams)] + [dp] * 10 + [C.POINTER(OrcStats)]
        L.orc_solve.restype = C.c_int
        L.orc_solve_batch.argtypes = [C.POINTER(OrcParams), C.c_int, dp, dp, dp, dp, dp, C.c_int, dp,
                                      dp, dp, dp, ip, ip, ip, C.c_int]
        L.orc_solve_batch.restype = C.c_int
        L.orc_eval.argtypes = [C.POINTER(OrcParams)] + [dp] * 8 + [C.c_double] + [dp] * 4
        L.orc_ldl_solve.argtypes = [C.c_int, dp, dp, ip]
        L.orc_ldl_solve.restype = C.c_int
        _LIB = L
    return _LIB


DEFAULT_YAML = dict(dt=0.2, N=30, Q=[1.0, 1.0, 0.005], R=[1.0, 1.0], resolution=0.05, cost_factor=0.5,
                    costmap_size=2.0, inflation_radius=0.2, reverse_factor=5.0, rotation_factor=2.0,
                    look_ahead_distance=0.5, goal_threshold=0.2)


def variant_params(variant, y=None, N=None, obstacles=None, **over):
    """orc_params for reference variant 'A' | 'B' | 'C' built from a params.yaml dict.

    obstacles=True on B enables the (built-but-dropped) gauss obstacle cost — non-reference."""
    y = dict(DEFAULT_YAML if y is None else y)
    p = OrcParams()
    lib().orc_default_options(C.byref(p))
    p.N = int(y["N"] if N is None else N)
    p.M = int((y["costmap_size"] * 2) / y["resolution"]) * 2
    p.dt = float(y["dt"])
    p.obs_r = float(y["inflation_radius"])
    p.obs_form = OBS_NONE
    p.obs_k0, p.obs_k1 = 0, -1
    if variant == "A":
        p.integrator, p.ref_kind = RK4, REF_GOAL
        p.Q[:] = [0.00005, 0.05, 0.05]          # mpc_point_stabilization.py:87-90
        p.R[:] = [0.01, 0.01]                   # :92-93
        p.kappa = float(y["cost_factor"])       # :35 (argument swap)
        p.obs_form, p.obs_c = OBS_EXPLOG, float(y["reverse_factor"])  # :33, :46-53
        p.obs_k0, p.obs_k1 = 0, p.N
        p.u_lo[:] = [-0.2, -0.1]; p.u_hi[:] = [0.2, 0.1]              # :82-83
    elif variant == "B":
        p.integrator, p.ref_kind = RK4, REF_GOAL
        p.Q[:] = [float(v) for v in y["Q"]]     # local_planner_point_stabilization.py:106-109
        p.R[:] = [0.5, 0.5]                     # :111-112
        p.kappa = float(y["cost_factor"])       # :47
        p.u_lo[:] = [-0.05, -0.2]; p.u_hi[:] = [0.15, 0.2]            # :101-102
        if obstacles:
            p.obs_form, p.obs_c = OBS_GAUSS, float(y["reverse_factor"])  # :43-45, :60-67
            p.obs_k0, p.obs_k1 = 0, p.N - 1
    elif variant == "C":
        p.integrator, p.ref_kind = EULER, REF_TRAJ
        p.Q[:] = [float(v) for v in y["Q"]]     # local_planner_tracking.py:108-111
        p.R[:] = [float(v) for v in y["R"]]     # :113-115
        p.kappa = float(y["reverse_factor"])    # :124
        p.u_lo[:] = [-0.1, -0.2]; p.u_hi[:] = [0.2, 0.2]              # :94-95
    else:
        raise ValueError(variant)
    for k, v in over.items():
        if k in ("u_lo", "u_hi", "Q", "R"):
            getattr(p, k)[:] = list(v)
        else:
            setattr(p, k, v)
    return p


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def solve(p, x0, xref, uref=None, obs_x=None, obs_y=None, u_init=None, x_init=None):
    """Single solve. Returns dict(X (3,N+1), U (2,N), cost, status, stats)."""
    N = p.N
    x0, xref, uref, obs_x, obs_y, u_init, x_init = map(_f64, (x0, xref, uref, obs_x, obs_y, u_init, x_init))
    if u_init is not None:
        u_init = np.ascontiguousarray(np.asarray(u_init).reshape(2, N).T)  # (2,N) -> stage-major
    if x_init is not None:
        x_init = np.ascontiguousarray(np.asarray(x_init).reshape(3, N + 1).T)
    X = np.zeros((N + 1, 3)); U = np.zeros((N, 2)); cost = np.zeros(1)
    st = OrcStats()
    status = lib().orc_solve(C.byref(p), _dp(x0), _dp(xref.ravel()), _dp(None if uref is None else uref.ravel()),
                             _dp(obs_x), _dp(obs_y), _dp(u_init), _dp(x_init), _dp(X), _dp(U), _dp(cost),
                             C.byref(st))
    return dict(X=X.T.copy(), U=U.T.copy(), cost=float(cost[0]), status=status,
                stats={f[0]: getattr(st, f[0]) for f in OrcStats._fields_})


def solve_batch(p, x0, xref, uref=None, obs_x=None, obs_y=None, u_init=None, nthreads=None):
    """Batch solve. x0 (B,3); xref (B,3) or (B,3N); uref (B,2N); obs_x/obs_y (B,M) or (M,) shared;
    u_init (B,2N) stage-major [v0,w0,v1,w1..]. Returns X (B,N+1,3), U (B,N,2), cost, status, iters, ls."""
    N = p.N
    x0 = _f64(x0); B = x0.shape[0]
    xref, uref, obs_x, obs_y, u_init = map(_f64, (xref, uref, obs_x, obs_y, u_init))
    stride = 0
    if obs_x is not None:
        stride = 0 if obs_x.ndim == 1 else obs_x.shape[1]
    X = np.zeros((B, N + 1, 3)); U = np.zeros((B, N, 2)); cost = np.zeros(B)
    status = np.zeros(B, np.int32); iters = np.zeros(B, np.int32); ls = np.zeros(B, np.int32)
    if nthreads is None:
        nthreads = len(os.sched_getaffinity(0))
    ip = C.POINTER(C.c_int)
    lib().orc_solve_batch(C.byref(p), B, _dp(x0), _dp(xref), _dp(uref), _dp(obs_x), _dp(obs_y), stride,
                          _dp(u_init), _dp(X), _dp(U), _dp(cost), status.ctypes.data_as(ip),
                          iters.ctypes.data_as(ip), ls.ctypes.data_as(ip), nthreads)
    return dict(X=X, U=U, cost=cost, status=status, iters=iters, ls=ls)


def evaluate(p, x0, xref, X, U, uref=None, obs_x=None, obs_y=None, lam=None, obj_scale=1.0):
    """NLP functions at a point. X (N+1,3) stage-major with X[0]==x0, U (N,2), lam (N,3)."""
    N = p.N
    x0, xref, uref, obs_x, obs_y, X, U, lam = map(_f64, (x0, xref, uref, obs_x, obs_y, X, U, lam))
    f = np.zeros(1); c = np.zeros((N, 3)); g = np.zeros(5 * N); st = np.zeros((N + 1, 36))
    lib().orc_eval(C.byref(p), _dp(x0), _dp(xref.ravel()), _dp(None if uref is None else uref.ravel()),
                   _dp(obs_x), _dp(obs_y), _dp(X), _dp(U), _dp(lam), float(obj_scale),
                   _dp(f), _dp(c), _dp(g), _dp(st))
    return dict(f=float(f[0]), c=c, grad=g, stages=st)


def kkt_certificate(p, x0, xref, X, U, uref=None, obs_x=None, obs_y=None, u_init=None, bound_tol=1e-6):
    """Independent first-order optimality certificate of a returned point (X (N+1,3), U (N,2)), computed from orc_eval
    only (no solver state): the multipliers of the shooting defects follow from the adjoint recursion
    lam_N = -g_N, lam_k = A_k' lam_{k+1} - g_k, the reduced gradient w.r.t. U_k is g_u - B_k' lam_{k+1}; it must vanish
    off the control bounds and point outwards on them.
    IPOPT's convergence test is on the SCALED problem: objective scaling df = min(1, 100 / |grad f|_inf) at the starting
    point (X = 0, U = u_init), and the dual infeasibility is divided by s_d = max(100, (|lam|_1 + |z|_1) / n) / 100.
    Returns dict(defect, stationarity (unscaled, controls within bound_tol of a bound count as active), df,
    scaled (= df * stationarity / s_d), complementarity (= df * max z_i * distance to the bound z_i acts on: the scaled
    complementarity IPOPT drives to mu; it covers interior controls, active bounds and everything between))."""
    N = p.N
    x0, X, U = _f64(x0), _f64(X), _f64(U)
    Xs = np.zeros((N + 1, 3)); Xs[0] = x0
    Us = np.zeros((N, 2)) if u_init is None else _f64(u_init).reshape(N, 2)
    gs = evaluate(p, x0, xref, Xs, Us, uref=uref, obs_x=obs_x, obs_y=obs_y)["grad"]
    gmax = float(np.abs(gs).max())
    df = max(min(1.0, 100.0 / gmax), 1e-8) if gmax > 100.0 else 1.0
    e = evaluate(p, x0, xref, X, U, uref=uref, obs_x=obs_x, obs_y=obs_y)
    g, st = e["grad"], e["stages"]
    lam = np.zeros((N + 2, 3))
    for k in range(N, 0, -1):
        gk = g[3 * (k - 1):3 * (k - 1) + 3]
        if k == N:
            lam[k] = -gk
        else:
            A = np.array([[1, 0, st[k, 0]], [0, 1, st[k, 1]], [0, 0, 1]])
            lam[k] = A.T @ lam[k + 1] - gk
    # stationarity in u with bound multipliers z_L, z_U >= 0:  rg - z_L + z_U = 0  ->  z_L = max(rg, 0), z_U = max(-rg, 0);
    # what remains to be checked is complementarity, z_L (u - lo) and z_U (hi - u) (an interior control needs rg = 0, a
    # control on a bound needs the right sign: both are "the product is small")
    viol, zsum, compl = 0.0, 0.0, 0.0
    for k in range(N):
        b11, b12, b21, b22 = st[k, 2:6]
        Bm = np.array([[b11, b12], [b21, b22], [0, p.dt]])
        rg = g[3 * N + 2 * k:3 * N + 2 * k + 2] - Bm.T @ lam[k + 1]
        zsum += 2.0 * float(np.abs(rg).sum())   # the multiplier of U - S = 0 and the active bound multiplier
        for i in range(2):
            dl, dh = max(U[k, i] - p.u_lo[i], 0.0), max(p.u_hi[i] - U[k, i], 0.0)
            compl = max(compl, rg[i] * dl if rg[i] > 0 else -rg[i] * dh)
            if dl <= bound_tol:
                viol = max(viol, max(0.0, -rg[i]))
            elif dh <= bound_tol:
                viol = max(viol, max(0.0, rg[i]))
            else:
                viol = max(viol, abs(rg[i]))
    sd = max(100.0, df * (float(np.abs(lam).sum()) + zsum) / (9 * N)) / 100.0
    return dict(defect=float(np.abs(e["c"]).max()), stationarity=float(viol), df=float(df), scaled=float(df * viol / sd),
                complementarity=float(df * compl))
