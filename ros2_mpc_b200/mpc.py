"""Drop-in `Mpc` classes: the reference's per-control-step solve surface on top of libb200mpc.so.

Reference surface mirrored (names, argument meaning, return values, error behaviour):
  Mpc()                         reads config/params.yaml; attributes dt, N, n_states, n_controls,
                                inflation_radius, Q, R                    (local_planner_point_stabilization.py:12-58)
  Mpc.perform_mpc(u0, initial_state, final_state, obstacles_x=None, obstacles_y=None)
        variant B -> u_opt[:, 0]                                          (local_planner_point_stabilization.py:69-87)
        variant A -> (x_opt (3,N+1), u_opt (2,N))                         (mpc_point_stabilization.py:55-68)
  Mpc.perform_mpc(u0, x0, pf, puf, obstacles_x=None, obstacles_y=None)
        variant C -> (x_opt (3,N+1), u_opt[:, 0])                         (local_planner_tracking.py:65-80)
  A failed solve raises RuntimeError carrying IPOPT's return_status name, as opti.solve() does.
Extension: perform_mpc_batch(...) solves B independent problems in one kernel launch; Mpc(devices=[0, 1, ...]) shards that
batch over several GPUs of the node (contiguous slices, no collective, results written into one set of host arrays).
"""
import numpy as np

from . import _shim
from .params import load_params
from .variants import make_params


class SolveError(RuntimeError):
    """opti.solve() raises RuntimeError on any status other than success; this subclass carries the status."""

    def __init__(self, status):
        self.status = int(status)
        self.status_name = _shim.STATUS_NAMES.get(self.status, str(self.status))
        super().__init__(f"Error in Opti::solve: Solver failed. return_status is '{self.status_name}'")


class _MpcBase:
    variant = None

    def __init__(self, params_path=None, device=0, N=None, obstacles=None, devices=None, **overrides):
        params = load_params(params_path)
        self.params = params
        self.dt = params["dt"]
        self.N = params["N"] if N is None else int(N)
        self.inflation_radius = params["inflation_radius"]
        self.Q = params["Q"]
        self.R = params["R"]
        self.n_states = 3
        self.n_controls = 2
        self._p = make_params(self.variant, params, N=self.N, obstacles=obstacles, **overrides)
        self.n_obstacles = self._p.M
        # devices=[0, 1, ...]: perform_mpc_batch shards its batch over these GPUs (one handle + host thread per device)
        self._solver = _shim.MultiSolver(self._p, devices) if devices else _shim.Solver(self._p, device=device)
        # opti.parameter values persist between calls until set again
        self._obs_x = None
        self._obs_y = None
        self.last_status = None
        self.last_iterations = None
        self.last_cost = None

    # -- helpers ------------------------------------------------------------------------------------------
    def _u0(self, u0, B=None):
        """casadi (2,N) initial guess -> stage-major (N,2)."""
        if u0 is None:
            return None
        u0 = np.asarray(u0, dtype=np.float64)
        if B is None:
            if u0.shape != (self.n_controls, self.N):
                raise ValueError(f"u0 must have shape ({self.n_controls}, {self.N})")
            return np.ascontiguousarray(u0.T)[None]
        if u0.shape == (B, self.n_controls, self.N):
            return np.ascontiguousarray(np.transpose(u0, (0, 2, 1)))
        if u0.shape == (self.n_controls, self.N):
            return np.ascontiguousarray(np.broadcast_to(u0.T, (B, self.N, 2)))
        raise ValueError(f"u0 must have shape (B, {self.n_controls}, {self.N})")

    def _obstacles(self, obstacles_x, obstacles_y):
        if obstacles_x is not None and obstacles_y is not None:
            ox = np.asarray(obstacles_x, dtype=np.float64).reshape(-1)
            oy = np.asarray(obstacles_y, dtype=np.float64).reshape(-1)
            if ox.shape[0] != self.n_obstacles or oy.shape[0] != self.n_obstacles:
                raise ValueError(f"obstacles_x / obstacles_y must have {self.n_obstacles} entries")
            self._obs_x, self._obs_y = ox, oy
        if self._p.obs_form != _shim.OBS_NONE and self._obs_x is None:
            # the reference fails inside CasADi here: the obstacle parameters are active but were never set
            raise RuntimeError("obstacles_x / obstacles_y have no value: the obstacle cost of this variant is active")
        if self._p.obs_form == _shim.OBS_NONE:
            return None, None
        return self._obs_x, self._obs_y

    def _finish_single(self, out):
        self.last_status = int(out["status"][0])
        self.last_iterations = int(out["iters"][0])
        self.last_cost = float(out["cost"][0])
        if self.last_status not in _shim.SUCCESS_STATUSES:
            raise SolveError(self.last_status)
        x_opt = np.ascontiguousarray(out["X"][0].T)  # (3, N+1) like sol.value(X)
        u_opt = np.ascontiguousarray(out["U"][0].T)  # (2, N)   like sol.value(U)
        return x_opt, u_opt

    def close(self):
        self._solver.close()


class _PointStabilization(_MpcBase):
    def _solve(self, u0, initial_state, final_state, obstacles_x, obstacles_y):
        ox, oy = self._obstacles(obstacles_x, obstacles_y)
        x0 = np.asarray(initial_state, dtype=np.float64).reshape(1, 3)
        goal = np.asarray(final_state, dtype=np.float64).reshape(1, 3)
        out = self._solver.solve_batch(x0, goal, obs_x=ox, obs_y=oy, u_init=self._u0(u0))
        return self._finish_single(out)

    def perform_mpc_batch(self, u0, initial_state, final_state, obstacles_x=None, obstacles_y=None):
        """B independent problems: initial_state (B,3), final_state (B,3), u0 (B,2,N) | (2,N) | None,
        obstacles (B,M) | (M,).  Returns dict(x_opt (B,3,N+1), u_opt (B,2,N), u0 (B,2), cost, status, iterations);
        per-problem failures are reported in `status`, not raised."""
        x0 = np.asarray(initial_state, dtype=np.float64).reshape(-1, 3)
        B = x0.shape[0]
        goal = np.asarray(final_state, dtype=np.float64).reshape(B, 3)
        ox = oy = None
        if self._p.obs_form != _shim.OBS_NONE:
            if obstacles_x is None or obstacles_y is None:
                raise RuntimeError("obstacles_x / obstacles_y are required: the obstacle cost of this variant is active")
            ox, oy = np.asarray(obstacles_x, dtype=np.float64), np.asarray(obstacles_y, dtype=np.float64)
        out = self._solver.solve_batch(x0, goal, obs_x=ox, obs_y=oy, u_init=self._u0(u0, B))
        return _batch_result(out)


def _batch_result(out):
    return dict(x_opt=np.transpose(out["X"], (0, 2, 1)), u_opt=np.transpose(out["U"], (0, 2, 1)),
                u0=out["U"][:, 0, :].copy(), cost=out["cost"], status=out["status"], iterations=out["iters"],
                ls_trials=out["ls"])


class MpcPointStabilizationLocal(_PointStabilization):
    """Variant B — ros2_mpc.planner.local_planner_point_stabilization.Mpc (what the launched node uses)."""
    variant = "B"

    def perform_mpc(self, u0, initial_state=np.array([0, 0, 0]), final_state=np.array([10, 10, 0]),
                    obstacles_x=None, obstacles_y=None):
        _, u_opt = self._solve(u0, initial_state, final_state, obstacles_x, obstacles_y)
        return u_opt[:, 0]


class MpcPointStabilization(_PointStabilization):
    """Variant A — ros2_mpc.mpc_point_stabilization.Mpc (obstacle cost active)."""
    variant = "A"

    def perform_mpc(self, u0, initial_state=np.array([0, 0, 0]), final_state=np.array([10, 10, 0]),
                    obstacles_x=None, obstacles_y=None):
        x_opt, u_opt = self._solve(u0, initial_state, final_state, obstacles_x, obstacles_y)
        return x_opt, u_opt


class MpcTracking(_MpcBase):
    """Variant C — ros2_mpc.planner.local_planner_tracking.Mpc."""
    variant = "C"

    def perform_mpc(self, u0, x0, pf, puf, obstacles_x=None, obstacles_y=None):
        self._obstacles(obstacles_x, obstacles_y)  # stored like opti.set_value; the cost term is zeroed (:41)
        N = self.N
        x0 = np.asarray(x0, dtype=np.float64).reshape(1, 3)
        pf = np.asarray(pf, dtype=np.float64).reshape(1, 3 * N)
        puf = np.asarray(puf, dtype=np.float64).reshape(1, 2 * N)
        out = self._solver.solve_batch(x0, pf, uref=puf, u_init=self._u0(u0))
        x_opt, u_opt = self._finish_single(out)
        return x_opt, u_opt[:, 0]

    def perform_mpc_batch(self, u0, x0, pf, puf):
        x0 = np.asarray(x0, dtype=np.float64).reshape(-1, 3)
        B = x0.shape[0]
        out = self._solver.solve_batch(x0, np.asarray(pf, dtype=np.float64).reshape(B, 3 * self.N),
                                       uref=np.asarray(puf, dtype=np.float64).reshape(B, 2 * self.N),
                                       u_init=self._u0(u0, B))
        return _batch_result(out)
