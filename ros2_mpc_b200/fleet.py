"""Closed loop of a fleet of point-stabilisation planners, resident on one GPU.

Per control step and robot the node loop of the reference (scripts/point_follower_local_planner.py:150-231) does
    goal_mpc = get_goal_for_mpc(path, headings, goal, pos, look_ahead_distance)          :179
    u = mpc.perform_mpc(zeros, x0, goal_mpc, obstacles_x, obstacles_y)                     :194
    acceleration limiter, goal-reached logic, publish the command                          :196-231
and the robot / simulator moves on.  Here the three stages are three kernel launches on one CUDA stream for all B
robots (b200mpc_goals_batch_device -> b200mpc_solve_batch_device -> b200mpc_control_step_device); nothing returns to
the host until `snapshot()`.  Variant B does not minimise the obstacle cost (:127 of the planner class), so the scan /
obstacle-list stage is not part of this loop; `FleetObstacleAvoidance` is the same loop for the obstacle-active variant A
(mpc_point_stabilization.py) with the sensor and get_obstacles in front: ray-cast of the shared map -> obstacle list ->
look-ahead goal -> solve -> control step, five launches per control step, all on the device.  torch only owns the
device buffers."""
import numpy as np

from . import _shim, load_params, make_params


class FleetPointStabilization:
    def __init__(self, start_state, goal, path_xy, path_heading, params=None, device=0, warm_start=False, quantise=True,
                 accel_limit=0.03):
        import torch  # noqa: PLC0415
        self._torch = torch
        y = params or load_params()
        self.params = y
        self.solver = _shim.Solver(make_params("B", y), device=device)
        self.dev = torch.device("cuda", device)
        self.N = y["N"]
        self.lookahead = float(y["look_ahead_distance"])
        self.goal_threshold = float(y["goal_threshold"])
        self.accel_limit = float(accel_limit)
        self.quantise = bool(quantise)
        f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.dev)  # noqa: E731
        start_state = np.atleast_2d(np.asarray(start_state, dtype=np.float64))
        self.B = B = start_state.shape[0]
        self.state = f(start_state)
        meas = start_state.copy()
        if quantise:
            meas = np.round(meas, 2)
        meas[:, 2] = meas[:, 2] % (2 * np.pi)
        self.x0 = f(meas)
        self.goal = f(np.atleast_2d(goal))                       # (B,5): x, y, -, -, yaw
        path_xy = np.asarray(path_xy, dtype=np.float64)
        self.per_robot = 1 if path_xy.ndim == 3 else 0
        self.K = path_xy.shape[-2]
        self.path_xy = f(path_xy)
        self.path_heading = f(np.asarray(path_heading, dtype=np.float64).reshape(path_xy.shape[:-1]))
        z = lambda *s, dt=torch.float64: torch.zeros(s, dtype=dt, device=self.dev)  # noqa: E731
        self.goal_mpc = z(B, 3)
        self.u_last = z(B, 2)
        self.goal_flag = z(B, dt=torch.int32)
        self.cmd = z(B, 2)
        self.X, self.U = z(B, self.N + 1, 3), z(B, self.N, 2)
        self.cost, self.status = z(B), z(B, dt=torch.int32)
        self.iters, self.ls = z(B, dt=torch.int32), z(B, dt=torch.int32)
        self.u_next = z(B, self.N, 2) if warm_start else None
        self.steps = 0

    def step(self, n=1):
        """n control steps for every robot, enqueued on torch's current stream; does not synchronise."""
        S, p = self.solver, (lambda t: 0 if t is None else t.data_ptr())
        stream = self._torch.cuda.current_stream().cuda_stream
        for _ in range(n):
            S.goals_batch_device(self.B, self.K, p(self.path_xy), p(self.path_heading), self.per_robot, p(self.goal),
                                 p(self.x0), 3, self.lookahead, p(self.goal_mpc), 0, stream=stream)
            S.solve_batch_device(self.B, p(self.x0), p(self.goal_mpc), 0, 0, 0, 0,
                                 p(self.u_next) if (self.u_next is not None and self.steps > 0) else 0,
                                 p(self.X), p(self.U), p(self.cost), p(self.status), p(self.iters), p(self.ls), stream=stream)
            S.control_step_device(self.B, p(self.U), p(self.status), p(self.state), p(self.x0), p(self.u_last), p(self.goal), 5,
                                  p(self.goal_flag), self.goal_threshold, self.accel_limit, self.quantise, p(self.cmd),
                                  p(self.u_next), stream=stream)
            self.steps += 1

    def snapshot(self):
        self._torch.cuda.synchronize()
        g = lambda t: t.cpu().numpy()  # noqa: E731
        return dict(state=g(self.state), x0=g(self.x0), cmd=g(self.cmd), goal_flag=g(self.goal_flag), u_last=g(self.u_last),
                    goal_mpc=g(self.goal_mpc), status=g(self.status), iters=g(self.iters), U=g(self.U))

    def close(self):
        self.solver.close()


class FleetObstacleAvoidance(FleetPointStabilization):
    """Closed loop of a fleet of the obstacle-active variant A on one shared static map, resident on the GPU.

    Per control step (all robots, one stream, no host round trip):
        scan      = lidar of the TRUE pose on the occupancy map            b200mpc_raycast_batch_device
                    (the simulator's job in the reference: a LaserScan message, core/ros_topics.py)
        obstacles = get_obstacles(scan, ..., pos, yaw of the MEASURED pose)  b200mpc_obstacles_batch_device
                    (scripts/point_follower_local_planner.py:88-118; more than 160 cells: the first 160, where the node raises)
        goal_mpc  = get_goal_for_mpc(...)                                    b200mpc_goals_batch_device   (:16-30)
        plan      = Mpc.perform_mpc(zeros, x0, goal_mpc, obstacles_x, obstacles_y), mpc_point_stabilization.py:55-68
        command, goal logic, plant step, next measurement                    b200mpc_control_step_device  (:196-231)
    m: map dict of synth.load_map() (occupancy, origin, resolution)."""

    def __init__(self, start_state, goal, path_xy, path_heading, m, params=None, device=0, warm_start=False, quantise=True,
                 accel_limit=0.03, n_beams=360, angle_min=0.0, angle_max=6.28, range_min=0.12, range_max=3.5, step=0.01,
                 slots=160):
        from . import obstacles as ob, sensors  # noqa: PLC0415
        super().__init__(start_state, goal, path_xy, path_heading, params=params, device=device, warm_start=warm_start,
                         quantise=quantise, accel_limit=accel_limit)
        torch, y = self._torch, self.params
        self.solver.close()
        self.solver = _shim.Solver(make_params("A", y), device=device)
        self.lidar = dict(n=int(n_beams), amin=float(angle_min), amax=float(angle_max), rmin=float(range_min),
                          rmax=float(range_max), step=float(step))
        self.slots = int(slots)
        self.size, self.res = float(y["costmap_size"]), float(y["resolution"])
        H, W = m["occ"].shape
        self.map_hw = (int(H), int(W))
        self.map_origin = (float(m["origin"][0]), float(m["origin"][1]))
        self.map_res = float(m["resolution"])
        self.occ_bits = torch.from_numpy(sensors.map_bits(m).view(np.int32).copy()).to(self.dev)
        bc, bs = ob.beam_table(self.lidar["n"], (angle_min, angle_max))
        f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.dev)  # noqa: E731
        self.beam_cos, self.beam_sin = f(bc), f(bs)
        B = self.B
        self.scan = torch.zeros((B, self.lidar["n"]), dtype=torch.float64, device=self.dev)
        self.obs_x = torch.zeros((B, self.slots), dtype=torch.float64, device=self.dev)
        self.obs_y = torch.zeros((B, self.slots), dtype=torch.float64, device=self.dev)
        self.obs_count = torch.zeros(B, dtype=torch.int32, device=self.dev)
        self.yaw = torch.zeros(B, dtype=torch.float64, device=self.dev)

    def step(self, n=1):
        S, p = self.solver, (lambda t: 0 if t is None else t.data_ptr())
        D = _shim.DevPtr
        stream = self._torch.cuda.current_stream().cuda_stream
        L = self.lidar
        H, W = self.map_hw
        for _ in range(n):
            S.device_call("b200mpc_raycast_batch_device", self.B, L["n"], D(p(self.occ_bits)), H, W, self.map_origin[0],
                          self.map_origin[1], self.map_res, D(p(self.state)), 3, L["amin"], L["amax"], L["rmin"], L["rmax"],
                          L["step"], D(p(self.scan)), D(stream))
            # get_obstacles takes pos and yaw of the measured pose; x0 rows are (x, y, yaw): pos = x0 with stride 3 is not an
            # argument of the C call (pos is [B][2]), so the two columns are gathered on the device
            pos = self.x0[:, :2].contiguous()
            self.yaw.copy_(self.x0[:, 2])
            S.obstacles_batch_device(self.B, L["n"], p(self.scan), p(self.beam_cos), p(self.beam_sin), p(pos), p(self.yaw),
                                     self.size, self.res, self.slots, p(self.obs_x), p(self.obs_y), p(self.obs_count),
                                     stream=stream)
            S.goals_batch_device(self.B, self.K, p(self.path_xy), p(self.path_heading), self.per_robot, p(self.goal),
                                 p(self.x0), 3, self.lookahead, p(self.goal_mpc), 0, stream=stream)
            S.solve_batch_device(self.B, p(self.x0), p(self.goal_mpc), 0, p(self.obs_x), p(self.obs_y), self.slots,
                                 p(self.u_next) if (self.u_next is not None and self.steps > 0) else 0,
                                 p(self.X), p(self.U), p(self.cost), p(self.status), p(self.iters), p(self.ls), stream=stream)
            S.control_step_device(self.B, p(self.U), p(self.status), p(self.state), p(self.x0), p(self.u_last), p(self.goal), 5,
                                  p(self.goal_flag), self.goal_threshold, self.accel_limit, self.quantise, p(self.cmd),
                                  p(self.u_next), stream=stream)
            self.steps += 1

    def snapshot(self):
        d = super().snapshot()
        g = lambda t: t.cpu().numpy()  # noqa: E731
        d.update(scan=g(self.scan), obs_x=g(self.obs_x), obs_y=g(self.obs_y), obs_count=g(self.obs_count), cost=g(self.cost))
        return d
