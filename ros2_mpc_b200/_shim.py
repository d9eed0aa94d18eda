"""ctypes binding of libb200mpc.so (include/b200mpc.h).  Thin: argument marshalling only.

There is no CPU fallback: if the shared library is missing, or no CUDA device is present, construction fails with
a RuntimeError naming the cause."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200MPC_LIB selects an alternative build of the same library (kernel-tuning experiments)
LIB_PATH = os.environ.get("B200MPC_LIB") or os.path.join(_HERE, "libb200mpc.so")
_LIB = None

RK4, EULER = 0, 1
OBS_NONE, OBS_GAUSS, OBS_EXPLOG = 0, 1, 2
REF_GOAL, REF_TRAJ = 0, 1
KERNEL_AUTO, KERNEL_WARP, KERNEL_LANE = 0, 1, 2

STATUS_NAMES = {
    0: "Solve_Succeeded", 1: "Solved_To_Acceptable_Level", 3: "Search_Direction_Becomes_Too_Small",
    -1: "Maximum_Iterations_Exceeded",
    -2: "Restoration_Failed", -3: "Error_In_Step_Computation", -13: "Invalid_Number_Detected",
}
SUCCESS_STATUSES = (0, 1)  # CasADi's return_success(): Solve_Succeeded, Solved_To_Acceptable_Level


class Params(C.Structure):
    """struct b200mpc_params"""
    _fields_ = [
        ("N", C.c_int32), ("M", C.c_int32), ("dt", C.c_double), ("integrator", C.c_int32), ("ref_kind", C.c_int32),
        ("Q", C.c_double * 3), ("R", C.c_double * 2), ("kappa", C.c_double),
        ("obs_form", C.c_int32), ("obs_k0", C.c_int32), ("obs_k1", C.c_int32), ("max_iter", C.c_int32),
        ("obs_c", C.c_double), ("obs_r", C.c_double), ("u_lo", C.c_double * 2), ("u_hi", C.c_double * 2),
        ("tol", C.c_double), ("acceptable_tol", C.c_double), ("mu_init", C.c_double),
        ("acceptable_iter", C.c_int32), ("max_soc", C.c_int32),
    ]


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  ros2_mpc_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p
        L.b200mpc_abi_version.restype = C.c_int
        L.b200mpc_default_options.argtypes = [C.POINTER(Params)]
        L.b200mpc_default_options.restype = None
        L.b200mpc_create.argtypes = [C.POINTER(Params), C.c_int]
        L.b200mpc_create.restype = vp
        L.b200mpc_destroy.argtypes = [vp]
        L.b200mpc_destroy.restype = None
        L.b200mpc_last_error.argtypes = [vp]
        L.b200mpc_last_error.restype = C.c_char_p
        L.b200mpc_solve_batch.argtypes = [vp, C.c_int, dp, dp, dp, dp, dp, C.c_int, dp, dp, dp, dp, ip, ip, ip]
        L.b200mpc_solve_batch.restype = C.c_int
        L.b200mpc_solve_batch_multi.argtypes = [C.POINTER(vp), C.c_int, C.c_int, dp, dp, dp, dp, dp, C.c_int, dp, dp, dp, dp, ip,
                                                ip, ip]
        L.b200mpc_solve_batch_multi.restype = C.c_int
        # device entry point: raw addresses
        L.b200mpc_solve_batch_device.argtypes = [vp, C.c_int] + [vp] * 5 + [C.c_int] + [vp] * 7 + [vp]
        L.b200mpc_solve_batch_device.restype = C.c_int
        L.b200mpc_eval_batch.argtypes = [vp, C.c_int, dp, dp, dp, dp, dp, C.c_int, dp, dp, dp, C.c_double,
                                         dp, dp, dp, dp]
        L.b200mpc_eval_batch.restype = C.c_int
        L.b200mpc_launch_count.argtypes = [vp]
        L.b200mpc_launch_count.restype = C.c_longlong
        L.b200mpc_last_kernel_ms.argtypes = [vp]
        L.b200mpc_last_kernel_ms.restype = C.c_float
        L.b200mpc_measure_fp64_peak.argtypes = [vp, C.POINTER(C.c_double)]
        L.b200mpc_measure_fp64_peak.restype = C.c_int
        L.b200mpc_set_kernel.argtypes = [vp, C.c_int]
        L.b200mpc_set_kernel.restype = C.c_int
        L.b200mpc_last_kernel_kind.argtypes = [vp]
        L.b200mpc_last_kernel_kind.restype = C.c_int
        L.b200mpc_last_solve_chunks.argtypes = [vp]
        L.b200mpc_last_solve_chunks.restype = C.c_int
        L.b200mpc_lane_kernel_stats.argtypes = [vp, C.POINTER(C.c_ulonglong * 8)]
        L.b200mpc_lane_kernel_stats.restype = C.c_int
        L.b200mpc_obstacles_batch.argtypes = [vp, C.c_int, C.c_int, dp, dp, dp, dp, dp, C.c_double, C.c_double, C.c_int,
                                              dp, dp, ip]
        L.b200mpc_obstacles_batch.restype = C.c_int
        L.b200mpc_obstacles_batch_device.argtypes = [vp, C.c_int, C.c_int] + [vp] * 5 + [C.c_double, C.c_double, C.c_int] + \
                                                     [vp] * 3 + [vp]
        L.b200mpc_obstacles_batch_device.restype = C.c_int
        L.b200mpc_goals_batch.argtypes = [vp, C.c_int, C.c_int, dp, dp, C.c_int, dp, dp, C.c_double, dp, ip]
        L.b200mpc_goals_batch.restype = C.c_int
        L.b200mpc_goals_batch_device.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int, C.c_double, vp, vp, vp]
        L.b200mpc_goals_batch_device.restype = C.c_int
        L.b200mpc_reftraj_batch.argtypes = [vp, C.c_int, C.c_int, dp, dp, dp, dp, C.c_int, C.c_int, dp, dp, dp, dp, ip]
        L.b200mpc_reftraj_batch.restype = C.c_int
        L.b200mpc_reftraj_batch_device.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]
        L.b200mpc_reftraj_batch_device.restype = C.c_int
        L.b200mpc_control_step_device.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, vp, C.c_double, C.c_double,
                                                  C.c_int, vp, vp, vp]
        L.b200mpc_control_step_device.restype = C.c_int
        u8p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
        L.b200mpc_dilate_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int, u8p]
        L.b200mpc_dilate_batch_device.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, vp]
        L.b200mpc_inflate_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, dp]
        L.b200mpc_inflate_batch_device.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp]
        L.b200mpc_local_costmap_batch.argtypes = [vp, C.c_int, C.c_int, dp, dp, dp, dp, C.c_double, C.c_double, C.c_int,
                                                  C.c_int, u8p]
        L.b200mpc_local_costmap_batch_device.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_double, C.c_double,
                                                         C.c_int, C.c_int, vp, vp]
        L.b200mpc_raycast_batch.argtypes = [vp, C.c_int, C.c_int, u32p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                            dp] + [C.c_double] * 5 + [dp]
        L.b200mpc_raycast_batch_device.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_double, C.c_double,
                                                   C.c_double, vp, C.c_int] + [C.c_double] * 5 + [vp, vp]
        L.b200mpc_headings_batch.argtypes = [vp, C.c_int, C.c_int, dp, C.c_double, dp, dp, dp]
        L.b200mpc_headings_batch_device.argtypes = [vp, C.c_int, C.c_int, vp, C.c_double, vp, vp, vp, vp]
        for fn in ("dilate", "inflate", "local_costmap", "raycast", "headings"):
            getattr(L, f"b200mpc_{fn}_batch").restype = C.c_int
            getattr(L, f"b200mpc_{fn}_batch_device").restype = C.c_int
        L.b200mpc_sizeof_params.restype = C.c_int
        if L.b200mpc_sizeof_params() != C.sizeof(Params):
            raise RuntimeError("b200mpc_params layout mismatch between _shim.Params and libb200mpc.so")
        _LIB = L
    return _LIB


def default_params():
    p = Params()
    lib().b200mpc_default_options(C.byref(p))
    return p


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class Solver:
    """Owns one b200mpc_handle (one CUDA device)."""

    def __init__(self, params, device=0):
        self._L = lib()
        self.params = params
        self._h = self._L.b200mpc_create(C.byref(params), int(device))
        if not self._h:
            raise RuntimeError("b200mpc_create failed: " + self._L.b200mpc_last_error(None).decode())
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.b200mpc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"b200mpc error {rc}: " + self._L.b200mpc_last_error(self._h).decode())

    @property
    def launch_count(self):
        return int(self._L.b200mpc_launch_count(self._h))

    def last_kernel_ms(self):
        return float(self._L.b200mpc_last_kernel_ms(self._h))

    def set_kernel(self, kind):
        """KERNEL_AUTO (by batch size), KERNEL_WARP (warp per problem) or KERNEL_LANE (lane per problem)."""
        self._check(self._L.b200mpc_set_kernel(self._h, int(kind)))

    @property
    def last_kernel_kind(self):
        return int(self._L.b200mpc_last_kernel_kind(self._h))

    @property
    def last_solve_chunks(self):
        """Chunks the most recent host-buffer solve was streamed in (0 = plain copy-in / solve / copy-out)."""
        return int(self._L.b200mpc_last_solve_chunks(self._h))

    def lane_kernel_stats(self):
        """Cumulative sweep statistics of the lane kernel (TPP_STATS builds): dict sweep -> (executions, mean lanes)."""
        a = (C.c_ulonglong * 8)()
        self._check(self._L.b200mpc_lane_kernel_stats(self._h, C.byref(a)))
        return {n: (int(a[2 * i]), (a[2 * i + 1] / a[2 * i]) if a[2 * i] else 0.0)
                for i, n in enumerate(("backward", "forward", "trial", "trip"))}

    def measure_fp64_peak(self):
        """FP64 FMA peak of the device [TFLOP/s], measured with a DFMA-saturating micro-kernel."""
        v = C.c_double(0.0)
        self._check(self._L.b200mpc_measure_fp64_peak(self._h, C.byref(v)))
        return float(v.value)

    def solve_batch(self, x0, xref, uref=None, obs_x=None, obs_y=None, u_init=None, out=None):
        """Host-buffer solve.  x0 (B,3); xref (B,3)|(B,3N); uref (B,2N); obs_x/obs_y (B,M) or (M,) shared;
        u_init (B,N,2).  Returns dict X (B,N+1,3), U (B,N,2), cost, status, iters, ls."""
        p = self.params
        N = p.N
        x0 = _f64(x0)
        B = x0.shape[0] if x0.ndim == 2 else 1
        x0 = x0.reshape(B, 3)
        nref = 3 if p.ref_kind == REF_GOAL else 3 * N
        xref = _f64(xref, (B, nref))
        uref = _f64(uref, (B, 2 * N)) if uref is not None else None
        stride = 0
        if obs_x is not None:
            obs_x, obs_y = _f64(obs_x), _f64(obs_y)
            if obs_x.shape[-1] != p.M or obs_y.shape != obs_x.shape:
                raise ValueError(f"obstacle lists must have {p.M} slots")
            stride = 0 if obs_x.ndim == 1 else p.M
            if obs_x.ndim == 2 and obs_x.shape[0] != B:
                raise ValueError("obstacle lists: batch size mismatch")
        u_init = _f64(u_init, (B, N, 2)) if u_init is not None else None
        if out is None:
            out = dict(X=np.empty((B, N + 1, 3)), U=np.empty((B, N, 2)), cost=np.empty(B),
                       status=np.empty(B, np.int32), iters=np.empty(B, np.int32), ls=np.empty(B, np.int32))
        rc = self._solve_call(B, _dp(x0), _dp(xref), _dp(uref), _dp(obs_x), _dp(obs_y), stride,
                              _dp(u_init), _dp(out["X"]), _dp(out["U"]), _dp(out["cost"]),
                              _ip(out["status"]), _ip(out["iters"]), _ip(out["ls"]))
        self._check(rc)
        return out

    def _solve_call(self, B, *ptrs):
        return self._L.b200mpc_solve_batch(self._h, B, *ptrs)

    def obstacles_batch(self, scan, beam_cos, beam_sin, pos, yaw, size, resolution, slots):
        """Obstacle lists for B robots (host buffers).  scan (B,n); beam_cos/beam_sin (n,); pos (B,2); yaw (B,).
        Returns obs_x, obs_y (B,slots) float64 and the raw occupied-cell count (B,) int32."""
        scan = _f64(scan)
        B, n = scan.shape
        bc, bs = _f64(beam_cos, (n,)), _f64(beam_sin, (n,))
        pos, yaw = _f64(pos, (B, 2)), _f64(yaw, (B,))
        ox, oy, cnt = np.empty((B, slots)), np.empty((B, slots)), np.empty(B, np.int32)
        rc = self._L.b200mpc_obstacles_batch(self._h, B, n, _dp(scan), _dp(bc), _dp(bs), _dp(pos), _dp(yaw), float(size),
                                             float(resolution), int(slots), _dp(ox), _dp(oy), _ip(cnt))
        self._check(rc)
        return ox, oy, cnt

    def obstacles_batch_device(self, B, n_beams, scan, beam_cos, beam_sin, pos, yaw, size, resolution, slots, obs_x, obs_y,
                               count, stream=0):
        """Device-buffer variant: raw device addresses (int, 0 = NULL for count); asynchronous on `stream`."""
        v = lambda a: C.c_void_p(int(a) if a else None)  # noqa: E731
        rc = self._L.b200mpc_obstacles_batch_device(self._h, int(B), int(n_beams), v(scan), v(beam_cos), v(beam_sin), v(pos),
                                                    v(yaw), float(size), float(resolution), int(slots), v(obs_x), v(obs_y),
                                                    v(count), v(stream))
        self._check(rc)

    def goals_batch(self, path_xy, path_heading, goal, pos, lookahead):
        """Look-ahead goals (get_goal_for_mpc) for B robots.  path_xy (K,2)|(B,K,2); path_heading (K,)|(B,K);
        goal (B,5); pos (B,2).  Returns goal_pose (B,3) and the chosen path index (B,) (-1 = final goal)."""
        path_xy, goal, pos = _f64(path_xy), _f64(goal), _f64(pos)
        B = goal.shape[0]
        per = 1 if path_xy.ndim == 3 else 0
        K = path_xy.shape[-2]
        path_heading = _f64(path_heading, (B, K) if per else (K,))
        pos = np.ascontiguousarray(pos[:, :2])
        out, idx = np.empty((B, 3)), np.empty(B, np.int32)
        self._check(self._L.b200mpc_goals_batch(self._h, B, K, _dp(path_xy), _dp(path_heading), per, _dp(goal), _dp(pos),
                                                float(lookahead), _dp(out), _ip(idx)))
        return out, idx

    def reftraj_batch(self, path_xy, path_heading, path_velocity, path_omega, x0, goal):
        """Tracking references (get_reference_trajectory) for B robots.  Returns pxf (B,3N), puf (B,2N), nearest (B,)."""
        path_xy, x0 = _f64(path_xy), _f64(x0)
        B = x0.shape[0]
        per = 1 if path_xy.ndim == 3 else 0
        K = path_xy.shape[-2]
        path_omega = _f64(path_omega)
        n_om = path_omega.shape[-1]
        path_heading = _f64(path_heading, (B, K) if per else (K,))
        path_velocity = _f64(path_velocity, (B, K) if per else (K,))
        goal = np.ascontiguousarray(_f64(goal)[:, :3])
        N = self.params.N
        pxf, puf, idx = np.empty((B, 3 * N)), np.empty((B, 2 * N)), np.empty(B, np.int32)
        self._check(self._L.b200mpc_reftraj_batch(self._h, B, K, _dp(path_xy), _dp(path_heading), _dp(path_velocity),
                                                  _dp(path_omega), n_om, per, _dp(x0), _dp(goal), _dp(pxf), _dp(puf), _ip(idx)))
        return pxf, puf, idx

    def goals_batch_device(self, B, K, path_xy, path_heading, per_robot_paths, goal, pos, pos_stride, lookahead, goal_out,
                           index_out=0, stream=0):
        v = lambda a: C.c_void_p(int(a) if a else None)  # noqa: E731
        self._check(self._L.b200mpc_goals_batch_device(self._h, int(B), int(K), v(path_xy), v(path_heading),
                                                       int(per_robot_paths), v(goal), v(pos), int(pos_stride),
                                                       float(lookahead), v(goal_out), v(index_out), v(stream)))

    def control_step_device(self, B, U_sol, status, state, x0, u_last, goal, goal_stride, goal_flag, goal_threshold,
                            accel_limit, quantise, cmd_out, u_next=0, stream=0):
        """Limiter / goal logic / plant step / next measurement for a fleet (raw device addresses)."""
        v = lambda a: C.c_void_p(int(a) if a else None)  # noqa: E731
        self._check(self._L.b200mpc_control_step_device(self._h, int(B), v(U_sol), v(status), v(state), v(x0), v(u_last),
                                                        v(goal), int(goal_stride), v(goal_flag), float(goal_threshold),
                                                        float(accel_limit), int(bool(quantise)), v(cmd_out), v(u_next),
                                                        v(stream)))

    def solve_batch_device(self, B, x0, xref, uref, obs_x, obs_y, obs_stride, u_init, X, U, cost, status, iters, ls,
                           stream=0):
        """Device-buffer solve: every argument is a raw device address (int, 0 = NULL); asynchronous on `stream`."""
        v = lambda a: C.c_void_p(int(a) if a else None)  # noqa: E731
        rc = self._L.b200mpc_solve_batch_device(self._h, int(B), v(x0), v(xref), v(uref), v(obs_x), v(obs_y),
                                                int(obs_stride), v(u_init), v(X), v(U), v(cost), v(status),
                                                v(iters), v(ls), v(stream))
        self._check(rc)

    def eval_batch(self, x0, xref, X, U, uref=None, obs_x=None, obs_y=None, lam=None, obj_scale=1.0):
        """NLP functions at given points.  X (B,N+1,3), U (B,N,2), lam (B,N,3)."""
        p = self.params
        N = p.N
        x0 = _f64(x0)
        B = x0.shape[0]
        nref = 3 if p.ref_kind == REF_GOAL else 3 * N
        xref = _f64(xref, (B, nref))
        uref = _f64(uref, (B, 2 * N)) if uref is not None else None
        stride = 0
        if obs_x is not None:
            obs_x, obs_y = _f64(obs_x), _f64(obs_y)
            stride = 0 if obs_x.ndim == 1 else p.M
        X, U = _f64(X, (B, N + 1, 3)), _f64(U, (B, N, 2))
        lam = _f64(lam, (B, N, 3)) if lam is not None else None
        f = np.empty(B); c = np.empty((B, N, 3)); g = np.empty((B, 5 * N)); st = np.empty((B, N + 1, 36))
        rc = self._L.b200mpc_eval_batch(self._h, B, _dp(x0), _dp(xref), _dp(uref), _dp(obs_x), _dp(obs_y), stride,
                                        _dp(X), _dp(U), _dp(lam), float(obj_scale), _dp(f), _dp(c), _dp(g), _dp(st))
        self._check(rc)
        return dict(f=f, c=c, grad=g, stages=st)

    # ---- costmap inflation / dilation (csrc/costmap_kernel.cuh) ------------------------------------------------------
    def dilate_batch(self, grids, kh=10, kw=10):
        """cv2.dilate(grid, np.ones((kh, kw)), iterations=1).astype(np.uint8) for grids (B,H,W) float64 -> (B,H,W) uint8."""
        g = _f64(grids)
        B, H, W = g.shape
        out = np.empty((B, H, W), np.uint8)
        self._check(self._L.b200mpc_dilate_batch(self._h, B, H, W, _dp(g), int(kh), int(kw),
                                                 out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def inflate_batch(self, grids, inflation_matrix, cells_inflation):
        """inflate_global (utils/costmap.py:5-20) for grids (B,H,W) float64 -> (B,H,W) float64."""
        g, m = _f64(grids), _f64(inflation_matrix)
        B, H, W = g.shape
        n = 2 * int(cells_inflation) + 1
        if m.shape != (n, n):
            raise ValueError(f"inflation_matrix must be ({n}, {n})")
        out = np.empty((B, H, W))
        self._check(self._L.b200mpc_inflate_batch(self._h, B, H, W, _dp(g), _dp(m), int(cells_inflation), _dp(out)))
        return out

    def local_costmap_batch(self, scan, beam_cos, beam_sin, yaw, size, resolution, kh=10, kw=10):
        """scan -> occupancy grid (rotation = yaw) -> dilate -> uint8 image (local_costmap_publisher.py:29-35), (B,nc,nc)."""
        scan = _f64(scan)
        B, n = scan.shape
        bc, bs, yaw = _f64(beam_cos, (n,)), _f64(beam_sin, (n,)), _f64(yaw, (B,))
        nc = int(float(size) * 2 / float(resolution))
        out = np.empty((B, nc, nc), np.uint8)
        self._check(self._L.b200mpc_local_costmap_batch(self._h, B, n, _dp(scan), _dp(bc), _dp(bs), _dp(yaw), float(size),
                                                        float(resolution), int(kh), int(kw),
                                                        out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def device_call(self, name, *args):
        """Generic caller of a *_device entry point: ints / floats are passed as such, device addresses as c_void_p
        (wrap them in `DevPtr`)."""
        conv = [C.c_void_p(int(a.addr) if a.addr else None) if isinstance(a, DevPtr) else a for a in args]
        self._check(getattr(self._L, name)(self._h, *conv))

    # ---- sensor model / path preprocessing (csrc/sensor_kernel.cuh) ---------------------------------------------------
    def raycast_batch(self, occ_bits, H, W, origin, resolution, pose, angle_min=0.0, angle_max=6.28, range_min=0.12,
                      range_max=3.5, step=0.01, n_beams=360):
        """Laser scans (B,n_beams) of the shared occupancy map for poses (B,3).  occ_bits: (H, ceil(W/32)) uint32."""
        pose = _f64(pose)
        B = pose.shape[0]
        occ_bits = np.ascontiguousarray(occ_bits, dtype=np.uint32)
        scan = np.empty((B, n_beams))
        self._check(self._L.b200mpc_raycast_batch(self._h, B, int(n_beams), occ_bits.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                  int(H), int(W), float(origin[0]), float(origin[1]), float(resolution),
                                                  _dp(np.ascontiguousarray(pose[:, :3])), float(angle_min), float(angle_max),
                                                  float(range_min), float(range_max), float(step), _dp(scan)))
        return scan

    def headings_batch(self, path_xy, dt):
        """get_headings for P paths: path_xy (P,K,2) -> heading (P,K), velocity (P,K), omega (P,K-1)."""
        path_xy = _f64(path_xy)
        P, K, _ = path_xy.shape
        h, v, w = np.empty((P, K)), np.empty((P, K)), np.empty((P, K - 1))
        self._check(self._L.b200mpc_headings_batch(self._h, P, K, _dp(path_xy), float(dt), _dp(h), _dp(v), _dp(w)))
        return h, v, w


class MultiSolver(Solver):
    """One batch sharded over several GPUs of the node (b200mpc_solve_batch_multi): a handle, a host thread and streams per
    device, contiguous slices of the batch index, results written straight into the caller's arrays.  solve_batch has the
    signature and the results of Solver.solve_batch; the per-device entry points act on the first device."""

    def __init__(self, params, devices):
        devices = [int(d) for d in devices]
        if len(devices) < 1 or len(set(devices)) != len(devices):
            raise ValueError("devices must be a non-empty list of distinct device indices")
        self._solvers = [Solver(params, device=d) for d in devices]
        first = self._solvers[0]
        self._L, self.params, self._h, self.device, self.devices = first._L, params, first._h, first.device, devices
        self._handles = (C.c_void_p * len(devices))(*[s._h for s in self._solvers])

    def _solve_call(self, B, *ptrs):
        return self._L.b200mpc_solve_batch_multi(self._handles, len(self._solvers), B, *ptrs)

    def shard_kernel_ms(self):
        """Device time of each shard's solve kernel of the most recent call."""
        return [s.last_kernel_ms() for s in self._solvers]

    def close(self):
        for s in getattr(self, "_solvers", []):
            s.close()
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DevPtr:
    """A raw device address for Solver.device_call."""

    def __init__(self, addr):
        self.addr = int(addr) if addr else 0
