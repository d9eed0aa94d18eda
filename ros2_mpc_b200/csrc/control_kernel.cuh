// control_kernel.cuh — what the planner node does with the solution of a control step, for a fleet of robots, on the
// device, so that closed loops run many control steps without host round trips:
//   * acceleration limiter and goal-reached logic        ros2_mpc/scripts/point_follower_local_planner.py:196-231
//   * the measured state the next solve starts from       :172 (x0 = [pos, yaw % 2 pi]) with the odometry
//     subscriber's rounding to two decimals               ros2_mpc/core/ros_topics.py:66-80
//   * a plant for simulation: the RK4 unicycle of the model (the reference drives Gazebo / a TurtleBot3 instead)
//   * optional warm start: the previous plan shifted by one stage (the reference always passes zeros, :174)
// One thread per robot; every quantity a decision depends on is formed with separately rounded IEEE operations in the
// order numpy uses, so flags and commands are bit-exact with the reference's expressions.
#pragma once

struct ControlArgs {
    int B, N;
    const double *U;        // [B][N][2] solution of this control step
    const int *status;      // [B] solver status (0 / 1 = success); NULL = all succeeded
    double *state;          // [B][3] true pose of the simulated robot (x, y, yaw), advanced by dt
    double *x0;             // [B][3] measured state: in = what the solve started from, out = next measurement
    double *u_last;         // [B][2]
    const double *goal;     // [B][goal_stride]: (x, y, ...)
    int goal_stride;
    int *goal_flag;         // [B] GOAL_FLAG of the node loop
    double goal_threshold, accel_limit, dt;
    int quantise;           // round the odometry to two decimals like OdomSubscriber.odom_callback
    double *cmd;            // [B][2] the command the robot executes during this step
    double *u_next;         // [B][N][2] or NULL: shifted plan for a warm start of the next solve
};

// np.round(v, 2): rint(v * 100) / 100
__device__ __forceinline__ double ctl_round2(double v) { return __ddiv_rn(rint(__dmul_rn(v, 100.0)), 100.0); }
__device__ __forceinline__ double ctl_pymod(double a, double m) {
    double r = fmod(a, m);
    if (r != 0.0) { if ((m < 0.0) != (r < 0.0)) r += m; }
    else r = copysign(0.0, m);
    return r;
}
// np.linalg.norm of one 2-vector = sqrt(x.dot(x)); the BLAS dot product fuses the second product into the sum
// (tests/golden/control_golden.npz has cases exactly on both thresholds)
__device__ __forceinline__ double ctl_norm2(double dx, double dy) {
    return sqrt(__fma_rn(dy, dy, __dmul_rn(dx, dx)));
}

__global__ void __launch_bounds__(128) control_step_kernel(const ControlArgs a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const double *Ub = a.U + (size_t)b * a.N * 2;
    const bool solved = !a.status || a.status[b] == 0 || a.status[b] == 1;
    double u0 = solved ? Ub[0] : 0.0, u1 = solved ? Ub[1] : 0.0; // a failed solve stops the robot (the node would raise)
    double ul0 = a.u_last[2 * (size_t)b], ul1 = a.u_last[2 * (size_t)b + 1];
    int flag = a.goal_flag[b];
    double c0, c1;
    // :196-205
    if (flag) { c0 = 0.0; c1 = 0.0; }
    else if (ctl_norm2(u0 - ul0, u1 - ul1) > a.accel_limit) {
        c0 = __dadd_rn(ul0, a.accel_limit); c1 = __dadd_rn(ul1, a.accel_limit); // (the reference adds the limit to both)
        ul0 = u0; ul1 = u1;
    } else { c0 = u0; c1 = u1; ul0 = u0; ul1 = u1; }
    // :207-231, with the measured state the solve started from
    const double mx = a.x0[3 * (size_t)b], my = a.x0[3 * (size_t)b + 1];
    const double *g = a.goal + (size_t)b * a.goal_stride;
    if (ctl_norm2(mx - g[0], my - g[1]) > a.goal_threshold) flag = 0;
    else if (!flag) { c0 = 0.0; c1 = 0.0; flag = 1; }
    a.u_last[2 * (size_t)b] = ul0; a.u_last[2 * (size_t)b + 1] = ul1;
    a.goal_flag[b] = flag;
    a.cmd[2 * (size_t)b] = c0; a.cmd[2 * (size_t)b + 1] = c1;
    // plant: one RK4 step of the unicycle under the zero-order-hold command
    double x = a.state[3 * (size_t)b], y = a.state[3 * (size_t)b + 1], th = a.state[3 * (size_t)b + 2];
    {
        double s0, k0, sm, km, se, ke;
        sincos(th, &s0, &k0);
        sincos(th + 0.5 * a.dt * c1, &sm, &km);
        sincos(th + a.dt * c1, &se, &ke);
        const double h = a.dt / 6.0;
        x += h * c0 * (k0 + 4.0 * km + ke);
        y += h * c0 * (s0 + 4.0 * sm + se);
        th += a.dt * c1;
    }
    a.state[3 * (size_t)b] = x; a.state[3 * (size_t)b + 1] = y; a.state[3 * (size_t)b + 2] = th;
    // next measurement (:172): pos, ori rounded by the odometry subscriber, yaw % 2 pi
    const double two_pi = 2.0 * 3.141592653589793;
    const double qx = a.quantise ? ctl_round2(x) : x, qy = a.quantise ? ctl_round2(y) : y;
    const double qt = a.quantise ? ctl_round2(th) : th;
    a.x0[3 * (size_t)b] = qx; a.x0[3 * (size_t)b + 1] = qy; a.x0[3 * (size_t)b + 2] = ctl_pymod(qt, two_pi);
    // warm start: plan shifted by one stage, last stage repeated
    if (a.u_next) {
        double *un = a.u_next + (size_t)b * a.N * 2;
        for (int k = 0; k < a.N; k++) {
            const int kk = (k + 1 < a.N) ? k + 1 : a.N - 1;
            un[2 * k] = solved ? Ub[2 * kk] : 0.0;
            un[2 * k + 1] = solved ? Ub[2 * kk + 1] : 0.0;
        }
    }
}
