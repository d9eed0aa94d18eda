/*
 * b200mpc.h — C ABI of libb200mpc.so: batched nonlinear-MPC solve on B200 (sm_100a), FP64.
 *
 * Drop-in boundary for the per-control-step solve of nitesh-subedi/ros2_mpc.  Each entry point names the
 * reference interface it replaces (paths under the reference repository):
 *
 *   b200mpc_create        <- Mpc.__init__            ros2_mpc/planner/local_planner_point_stabilization.py:12-57
 *                                                    ros2_mpc/mpc_point_stabilization.py:10-44
 *                                                    ros2_mpc/planner/local_planner_tracking.py:12-53
 *                            (casadi.Opti problem construction + opti.solver("ipopt", opts))
 *   b200mpc_solve_batch   <- Mpc.perform_mpc         local_planner_point_stabilization.py:69-87,
 *                                                    mpc_point_stabilization.py:55-68, local_planner_tracking.py:65-80
 *                            (opti.set_initial / set_value / solve / sol.value), batched over independent problems
 *   b200mpc_solve_batch_device  same, caller-owned device buffers, asynchronous on a CUDA stream
 *   b200mpc_solve_batch_multi   same, one batch sharded over the GPUs of a node (one host thread + handle per device)
 *   b200mpc_eval_batch    <- the NLP functions CasADi evaluates inside opti.solve():
 *                            rk4 :136-148, euler_integration (tracking) :132-137, define_cost_function :104-127,
 *                            define_obstacles_cost_function (mpc_point_stabilization.py:46-53)
 *   b200mpc_obstacles_batch[_device]  <- get_obstacles   ros2_mpc/scripts/point_follower_local_planner.py:88-118
 *                            with utils.convert_laser_scan_to_occupancy_grid (ros2_mpc/utils/utils.py:5-43),
 *                            convert_to_map_coordinates (:114-124), rotate_coordinates (:145-152): the producer of the
 *                            obstacles_x / obstacles_y arguments of perform_mpc, batched over robots
 *   b200mpc_goals_batch[_device]    <- get_goal_for_mpc   ros2_mpc/scripts/point_follower_local_planner.py:16-30
 *                            (the final_state argument of perform_mpc, variants A / B), batched over robots
 *   b200mpc_reftraj_batch[_device]  <- get_reference_trajectory   ros2_mpc/scripts/path_follower_local_planner.py:27-73
 *                            (the pf / puf arguments of perform_mpc, variant C), batched over robots
 *   b200mpc_control_step_device     <- the node loop after the solve   ros2_mpc/scripts/point_follower_local_planner.py:196-231
 *                            (acceleration limiter, goal-reached logic), the next measured state (:172 with the rounding
 *                            of ros2_mpc/core/ros_topics.py:66-80) and a simulated plant step, for a fleet on the device
 *   b200mpc_dilate_batch[_device]        <- cv2.dilate(grid, np.ones((10, 10)), iterations=1).astype(np.uint8)
 *                            ros2_mpc/core/local_costmap_publisher.py:34-35, ros2_mpc/core/global_costmap_publisher.py (same call)
 *   b200mpc_inflate_batch[_device]       <- inflate_global / inflate_local   ros2_mpc/utils/costmap.py:5-41
 *   b200mpc_local_costmap_batch[_device] <- the loop body of the local costmap publisher   ros2_mpc/core/local_costmap_publisher.py:29-35
 *                            (convert_laser_scan_to_occupancy_grid with rotation = yaw, utils.py:5-43, then the dilation), fused
 *   b200mpc_raycast_batch[_device]       the laser scanner of the simulated robots: scans of the shared static map
 *                            (maps/map_carto.pgm with the pixel convention of ros2_mpc/core/map_server.py:14-20), the input
 *                            LaserSubscriber.get_scan() delivers to get_obstacles (ros2_mpc/core/ros_topics.py:103)
 *   b200mpc_headings_batch[_device]      <- get_headings   ros2_mpc/scripts/path_follower_local_planner.py:14-23
 *   b200mpc_destroy       <- garbage collection of the Mpc / Opti object
 *
 * Conventions: plain pointers and sizes only; no C++ exceptions cross the boundary; functions return 0 on
 * success and a negative B200MPC_E_* code otherwise (text via b200mpc_last_error).  Per-problem solver outcomes
 * are reported in status_out with IPOPT's ApplicationReturnStatus values (the reference raises RuntimeError from
 * opti.solve() for any status other than Solve_Succeeded / Solved_To_Acceptable_Level).
 * A handle is bound to one device and is not thread-safe (one handle per thread and device), matching the
 * reference's one-Mpc-per-process usage.  ONE solve may be in flight per handle: the work-queue counter, the lane
 * kernel's workspace, the hand-over records and the timing events belong to the handle, so a second
 * b200mpc_solve_batch_device on another stream has to wait for the first (use one handle per stream; two handles
 * alternating overlap the tail of one batch with the start of the next).  There is no CPU fallback: every entry point
 * fails if no CUDA device is available.
 *
 * Layouts (all FP64, problem-major, "stage-major" inside a problem):
 *   x0      [B][3]                 initial state (x, y, theta)                       = P[0:3]
 *   xref    [B][3]   (ref_kind 0)  goal state                                        = P[3:6]
 *           [B][3N]  (ref_kind 1)  state reference ref_1..ref_N                      = P_X[3:]
 *   uref    [B][2N]  (ref_kind 1)  control reference, NULL otherwise                 = P_U
 *   obs_x/y [B][M] with obs_stride = M, or one shared list [M] with obs_stride = 0   = obstacles_x / obstacles_y
 *   u_init  [B][N][2] or NULL (= zeros)      the u0 argument of perform_mpc (casadi shape (2,N), column-major)
 *   X_out   [B][N+1][3]   predicted states, X_out[b][0] == x0[b]      (sol.value(X) is its transpose, (3,N+1))
 *   U_out   [B][N][2]     optimal controls                            (sol.value(U) is its transpose, (2,N))
 *   cost_out[B]           objective value at the returned point
 *   status_out[B], iters_out[B] (accepted interior-point iterations), ls_out[B] (extra line-search trial evaluations)
 */
#ifndef B200MPC_H
#define B200MPC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MPC_ABI_VERSION 1
#define B200MPC_MAX_N 127 /* horizon limit of the warp-per-problem kernel (4 stages per lane) */
#define B200MPC_MAX_M 1024 /* obstacle slots staged in shared memory per problem */

/* integrator */
#define B200MPC_RK4 0
#define B200MPC_EULER 1
/* obstacle cost form */
#define B200MPC_OBS_NONE 0
#define B200MPC_OBS_GAUSS 1  /* c*exp(-s),  s = ((x-ox)^2+(y-oy)^2)/r^2 */
#define B200MPC_OBS_EXPLOG 2 /* exp(c/s) = exp(c*exp(-log(s))) */
/* reference kind */
#define B200MPC_REF_GOAL 0
#define B200MPC_REF_TRAJ 1

/* per-problem status: IPOPT ApplicationReturnStatus */
#define B200MPC_SOLVE_SUCCEEDED 0
#define B200MPC_SOLVED_TO_ACCEPTABLE_LEVEL 1
#define B200MPC_SEARCH_DIRECTION_TOO_SMALL 3 /* tiny steps at the smallest barrier parameter: "solved to best possible numerical accuracy" (not a success for Opti) */
#define B200MPC_MAXITER_EXCEEDED (-1)
#define B200MPC_RESTORATION_FAILED (-2)
#define B200MPC_ERROR_IN_STEP_COMPUTATION (-3)
#define B200MPC_INVALID_NUMBER_DETECTED (-13)

/* solve kernels (b200mpc_set_kernel) */
#define B200MPC_KERNEL_AUTO 0 /* by batch size: lane-per-problem from B200MPC_LANE_KERNEL_MIN_BATCH problems on */
#define B200MPC_KERNEL_WARP 1 /* one warp per problem, horizon in registers: small batches, single-solve latency */
#define B200MPC_KERNEL_LANE 2 /* one lane per problem, horizon streamed through an HBM workspace: large batches */
#define B200MPC_LANE_KERNEL_MIN_BATCH 30720 /* measured crossover: a lane-kernel launch lasts >= ~20 ms (60-70 trips), the warp kernel solves 1.4 M problems/s */

/* library error codes */
#define B200MPC_E_ARG (-1)
#define B200MPC_E_CUDA (-2)
#define B200MPC_E_NODEVICE (-3)
#define B200MPC_E_NOMEM (-4)

typedef struct b200mpc_params {
    int32_t N;          /* horizon, params.yaml "N" */
    int32_t M;          /* obstacle slots: int(costmap_size*2/resolution)*2 */
    double dt;          /* params.yaml "dt" */
    int32_t integrator; /* B200MPC_RK4 | B200MPC_EULER */
    int32_t ref_kind;   /* B200MPC_REF_GOAL | B200MPC_REF_TRAJ */
    double Q[3];        /* diagonal state weights */
    double R[2];        /* diagonal control weights */
    double kappa;       /* exponent of the reverse penalty (1/exp(v))**kappa */
    int32_t obs_form;   /* B200MPC_OBS_* */
    int32_t obs_k0;     /* first stage carrying the obstacle sum */
    int32_t obs_k1;     /* last stage carrying the obstacle sum (inclusive) */
    int32_t max_iter;   /* ipopt max_iter (3000) */
    double obs_c;       /* obstacle cost factor */
    double obs_r;       /* inflation radius */
    double u_lo[2];     /* lower bounds on (v, omega) */
    double u_hi[2];     /* upper bounds on (v, omega) */
    double tol;             /* ipopt tol (1e-8) */
    double acceptable_tol;  /* 1e-6 */
    double mu_init;         /* 0.1 */
    int32_t acceptable_iter; /* 15 */
    int32_t max_soc;         /* 4 */
} b200mpc_params;

typedef struct b200mpc_handle b200mpc_handle;

/* Fills the solver options with IPOPT's defaults (tol, max_iter, acceptable_*, mu_init, max_soc). */
void b200mpc_default_options(b200mpc_params *p);

/* Creates a solver bound to CUDA device `device`.  Returns NULL on failure (see b200mpc_last_error(NULL)). */
b200mpc_handle *b200mpc_create(const b200mpc_params *p, int device);
void b200mpc_destroy(b200mpc_handle *h);

/* Last error text of the handle (or of the last failed b200mpc_create when h == NULL). */
const char *b200mpc_last_error(const b200mpc_handle *h);

/* Host-buffer solve: copies inputs host->device, runs the solve kernel, copies results back, blocks until done.
 * Any of cost_out / iters_out / ls_out may be NULL.
 * Large batches (>= 131072 problems on the lane-per-problem kernel) whose buffers are page-locked are STREAMED: the
 * inputs are copied in chunks on a copy stream while the persistent kernel already solves the first chunks, and each
 * chunk of results is copied out as soon as its last problem has finished (the kernel raises a per-chunk flag in
 * host memory).  Pageable buffers take the plain copy-in / solve / copy-out sequence.  Results are identical. */
int b200mpc_solve_batch(b200mpc_handle *h, int B, const double *x0, const double *xref, const double *uref,
                        const double *obs_x, const double *obs_y, int obs_stride, const double *u_init,
                        double *X_out, double *U_out, double *cost_out, int32_t *status_out, int32_t *iters_out,
                        int32_t *ls_out);

/* Device-buffer solve: every pointer is a device pointer on the handle's device; the kernel is enqueued on
 * `stream` (a cudaStream_t passed as void*, NULL = the legacy default stream) and the call returns without
 * synchronising.  Same NULL rules as above. */
int b200mpc_solve_batch_device(b200mpc_handle *h, int B, const double *x0, const double *xref, const double *uref,
                               const double *obs_x, const double *obs_y, int obs_stride, const double *u_init,
                               double *X_out, double *U_out, double *cost_out, int32_t *status_out,
                               int32_t *iters_out, int32_t *ls_out, void *stream);

/* Host-buffer solve of ONE batch on G devices of the node: `handles` are G handles created from the same
 * parameters on G different devices; device g solves the contiguous slice [lo_g, hi_g) of the batch (sizes differ by at most
 * one) on its own host thread and streams, reads its inputs from and writes its results straight into that slice of the
 * caller's arrays (no gather copy, no collective).  Arguments and NULL rules as b200mpc_solve_batch; a shared obstacle list
 * (obs_stride = 0) goes to every device.  Page-locked buffers are streamed per device as in b200mpc_solve_batch.
 * Blocks until every device has finished; returns the first device's error code if one fails (text on handles[0]). */
int b200mpc_solve_batch_multi(b200mpc_handle **handles, int G, int B, const double *x0, const double *xref,
                              const double *uref, const double *obs_x, const double *obs_y, int obs_stride,
                              const double *u_init, double *X_out, double *U_out, double *cost_out, int32_t *status_out,
                              int32_t *iters_out, int32_t *ls_out);

/* NLP function evaluation at given points (host buffers, blocking): for each problem b, X[b][N+1][3] (X[b][0]
 * must equal x0[b]), U[b][N][2], lam[b][N][3] (multipliers of c_k = X_k - F(X_{k-1},U_{k-1}), k = 1..N; NULL = 0):
 *   f_out[B]              objective
 *   c_out[B][N][3]        shooting defects
 *   grad_out[B][5N]       objective gradient w.r.t. (X_1..X_N, U_0..U_{N-1})
 *   stage_out[B][N+1][36] per stage: a13,a23,b11,b12,b21,b22, Lagrangian Hessian 5x5 row-major on
 *                         (x,y,theta,v,omega), c_{k+1}[3], 2 pad.
 * Any output pointer may be NULL. */
int b200mpc_eval_batch(b200mpc_handle *h, int B, const double *x0, const double *xref, const double *uref,
                       const double *obs_x, const double *obs_y, int obs_stride, const double *X,
                       const double *U, const double *lam, double obj_scale, double *f_out, double *c_out,
                       double *grad_out, double *stage_out);

/* Obstacle-list construction for B robots (get_obstacles of the reference, see the header comment).
 *   scan      [B][n_beams]  laser ranges (NaN / +-inf allowed, handled as the reference does)
 *   beam_cos, beam_sin [n_beams]  cos / sin of the beam angles i*(angle_max-angle_min)/n_beams + angle_min — a property
 *             of the lidar, computed once by the caller exactly as utils.py:18-20 does (this keeps the cell indices
 *             bit-exact with the reference; the device only multiplies, adds and divides)
 *   pos [B][2], yaw [B]     robot pose (pos, ori[2])
 *   size, resolution        params.yaml costmap_size and resolution (the local grid spans 2*size metres)
 *   obs_x, obs_y [B][slots] obstacle points in world coordinates: the occupied cells in np.where order of the
 *             180-degree-rotated grid, padded with the first one; all 100.0 when the scan marks no cell
 *   count [B] (may be NULL) number of occupied cells; count > slots means the list was truncated (the reference
 *             raises ValueError in that case — the Python mirror offers both behaviours)
 * The _device variant takes device pointers and enqueues on `stream` without synchronising, so that scan -> obstacle
 * list -> solve can run back to back on the device. */
int b200mpc_obstacles_batch(b200mpc_handle *h, int B, int n_beams, const double *scan, const double *beam_cos,
                            const double *beam_sin, const double *pos, const double *yaw, double size, double resolution,
                            int slots, double *obs_x, double *obs_y, int32_t *count);
int b200mpc_obstacles_batch_device(b200mpc_handle *h, int B, int n_beams, const double *scan, const double *beam_cos,
                                   const double *beam_sin, const double *pos, const double *yaw, double size,
                                   double resolution, int slots, double *obs_x, double *obs_y, int32_t *count,
                                   void *stream);

/* Look-ahead goals for B robots (get_goal_for_mpc).  path_xy [K][2] and path_heading [K] shared by the batch, or
 * [B][K][2] / [B][K] with per_robot_paths != 0; goal [B][5] (x, y, -, -, yaw: the reference reads goal[0], goal[1],
 * goal[4]); pos [B][2].  goal_out [B][3] = the final goal (yaw mod 2 pi) when it is nearer than `lookahead`, else the
 * first path point farther than `lookahead` (the nearest one if there is none) with its heading mod 2 pi.
 * index_out [B] (may be NULL): the chosen path index, -1 for the final goal.  Bit-exact with the reference. */
int b200mpc_goals_batch(b200mpc_handle *h, int B, int K, const double *path_xy, const double *path_heading,
                        int per_robot_paths, const double *goal, const double *pos, double lookahead, double *goal_out,
                        int32_t *index_out);
/* (device variant: pos has `pos_stride` doubles per robot, so the measured states x0 [B][3] can be passed directly) */
int b200mpc_goals_batch_device(b200mpc_handle *h, int B, int K, const double *path_xy, const double *path_heading,
                               int per_robot_paths, const double *goal, const double *pos, int pos_stride,
                               double lookahead, double *goal_out, int32_t *index_out, void *stream);

/* Tracking references for B robots (get_reference_trajectory; N = the handle's horizon).  path_xy [K][2],
 * path_heading [K], path_velocity [K], path_omega [n_omega] (n_omega = K-1 as get_headings returns it, or K) shared or
 * per robot; x0 [B][3]; goal [B][3] (the reference tiles goal[:3] when the robot is within 0.5 m of the path end).
 * pxf_out [B][3N] and puf_out [B][2N] are the P_X[3:] / P_U parameters (xref / uref of b200mpc_solve_batch);
 * index_out [B] (may be NULL) the nearest path index.  Every array is padded with its last element, as the reference
 * does.  Bit-exact with the reference. */
int b200mpc_reftraj_batch(b200mpc_handle *h, int B, int K, const double *path_xy, const double *path_heading,
                          const double *path_velocity, const double *path_omega, int n_omega, int per_robot_paths,
                          const double *x0, const double *goal, double *pxf_out, double *puf_out, int32_t *index_out);
int b200mpc_reftraj_batch_device(b200mpc_handle *h, int B, int K, const double *path_xy, const double *path_heading,
                                 const double *path_velocity, const double *path_omega, int n_omega, int per_robot_paths,
                                 const double *x0, const double *goal, double *pxf_out, double *puf_out,
                                 int32_t *index_out, void *stream);

/* One control step of a fleet after its solve, entirely on the device (all pointers are device pointers; asynchronous
 * on `stream`).  Per robot, as the node loop does (scripts/point_follower_local_planner.py:196-231):
 *   u = U_sol[b][0];  GOAL_FLAG set -> command (0,0);  |u - u_last| > accel_limit -> command u_last + accel_limit (both
 *   components, as the reference) else u;  u_last = u;  then with the measured position x0[b] the solve started from:
 *   farther than goal_threshold from goal[b][0:2] -> GOAL_FLAG = 0, else (first time) command (0,0), GOAL_FLAG = 1.
 * A failed solve (status not 0 / 1; status may be NULL) commands (0,0) — the node would raise.
 * Then state[b] (true pose) advances by one RK4 step of the unicycle under the command, x0[b] becomes the next
 * measurement (pose rounded to two decimals when `quantise`, yaw % 2 pi), cmd_out[b] the executed command, and u_next
 * (may be NULL) the plan shifted by one stage for a warm start.  N and dt are the handle's. */
int b200mpc_control_step_device(b200mpc_handle *h, int B, const double *U_sol, const int32_t *status, double *state,
                                double *x0, double *u_last, const double *goal, int goal_stride, int32_t *goal_flag,
                                double goal_threshold, double accel_limit, int quantise, double *cmd_out, double *u_next,
                                void *stream);

/* ---- costmap inflation / dilation (the costmap publishers' image operations; the MPC itself does not read them) ----
 * Dilation with a kh x kw box of ones, OpenCV's conventions for cv2.dilate(grid, np.ones((kh, kw))): anchor (kh/2, kw/2),
 * the border never wins:  out[y][x] = max grid[y + i - kh/2][x + j - kw/2], 0 <= i < kh, 0 <= j < kw, inside the image,
 * then the cast to uint8 (truncation; the grids hold 0 / 100).  grid [B][H][W] float64 -> out [B][H][W] uint8. */
int b200mpc_dilate_batch(b200mpc_handle *h, int B, int H, int W, const double *grid, int kh, int kw, uint8_t *out);
int b200mpc_dilate_batch_device(b200mpc_handle *h, int B, int H, int W, const double *grid, int kh, int kw, uint8_t *out,
                                void *stream);
/* inflate_global (utils/costmap.py:5-20): every cell whose value is exactly 0 and whose (2c+1)^2 window lies completely
 * inside the grid stamps np.minimum(window, inflation_matrix).  grid, out [B][H][W] float64; inflation_matrix [2c+1][2c+1].
 * inflate_local is the same on a cropped grid (the Python mirror computes the crop as the reference's slices do). */
int b200mpc_inflate_batch(b200mpc_handle *h, int B, int H, int W, const double *grid, const double *inflation_matrix,
                          int cells_inflation, double *out);
int b200mpc_inflate_batch_device(b200mpc_handle *h, int B, int H, int W, const double *grid, const double *inflation_matrix,
                                 int cells_inflation, double *out, void *stream);
/* Local costmap images for B robots: scan -> occupancy grid rotated by yaw (utils.py:5-43 with rotation = orientation[2])
 * -> dilation -> uint8 (0 / 100), out [B][nc][nc] with nc = int(2*size/resolution).  beam_cos / beam_sin as for
 * b200mpc_obstacles_batch.  The float64 grid never exists in memory. */
int b200mpc_local_costmap_batch(b200mpc_handle *h, int B, int n_beams, const double *scan, const double *beam_cos,
                                const double *beam_sin, const double *yaw, double size, double resolution, int kh, int kw,
                                uint8_t *out);
int b200mpc_local_costmap_batch_device(b200mpc_handle *h, int B, int n_beams, const double *scan, const double *beam_cos,
                                       const double *beam_sin, const double *yaw, double size, double resolution, int kh,
                                       int kw, uint8_t *out, void *stream);

/* ---- the simulated lidar: scans of a shared static occupancy map ----
 * occ_bits [H][ceil(W/32)] uint32: bit (c & 31) of word c >> 5 in row r = cell (r, c) occupied; row 0 is the lowest y
 * (core/map_server.py:20 flips the image), cell (r, c) covers [origin + c*res, origin + (c+1)*res) x [.. r ..].
 * pose [B][pose_stride] = (x, y, yaw, ...).  Beam i points along yaw + (i*(angle_max-angle_min)/n_beams + angle_min) and is
 * sampled at range_min + t*step, t = 0 .. round((range_max-range_min)/step); the first sample in an occupied cell is the
 * range, range_max otherwise.  scan [B][n_beams].  The map bits are staged in shared memory once per thread block. */
int b200mpc_raycast_batch(b200mpc_handle *h, int B, int n_beams, const uint32_t *occ_bits, int H, int W, double origin_x,
                          double origin_y, double resolution, const double *pose, double angle_min, double angle_max,
                          double range_min, double range_max, double step, double *scan);
int b200mpc_raycast_batch_device(b200mpc_handle *h, int B, int n_beams, const uint32_t *occ_bits, int H, int W,
                                 double origin_x, double origin_y, double resolution, const double *pose, int pose_stride,
                                 double angle_min, double angle_max, double range_min, double range_max, double step,
                                 double *scan, void *stream);

/* get_headings for P paths of K >= 2 points (path_xy [P][K][2]): heading [P][K] = arctan2 of the segments, the last one
 * repeated; velocity [P][K] = 2 * segment length / dt, the last one repeated; omega [P][K-1] = heading differences / 2. */
int b200mpc_headings_batch(b200mpc_handle *h, int P, int K, const double *path_xy, double dt, double *heading,
                           double *velocity, double *omega);
int b200mpc_headings_batch_device(b200mpc_handle *h, int P, int K, const double *path_xy, double dt, double *heading,
                                  double *velocity, double *omega, void *stream);

/* Forces one of the two solve kernels (default B200MPC_KERNEL_AUTO; the environment variable B200MPC_KERNEL=warp|lane
 * sets the default of new handles).  Both kernels run the same algorithm; results agree to rounding (with the obstacle
 * cost active the problem is non-convex and rounding can tip a solve to another local optimum).  AUTO takes the
 * lane-per-problem kernel from 30 720 problems on (131 072 with the obstacle cost). */
int b200mpc_set_kernel(b200mpc_handle *h, int kind);
/* B200MPC_KERNEL_WARP / _LANE: the kernel the most recent solve used. */
int b200mpc_last_kernel_kind(const b200mpc_handle *h);
/* Number of chunks the most recent host-buffer solve was streamed in (0 = plain copy-in / solve / copy-out). */
int b200mpc_last_solve_chunks(const b200mpc_handle *h);

/* Diagnostics of the lane-per-problem kernel, cumulative over the handle's life: out[2i], out[2i+1] = warp-level
 * executions and active lanes of sweep i (0 backward, 1 forward, 2 trial) and of the trips (i = 3).  Counted only
 * in library builds with -DTPP_STATS=1 (zeros otherwise). */
int b200mpc_lane_kernel_stats(b200mpc_handle *h, unsigned long long out[8]);

/* Number of kernel launches issued by this handle so far (bench.py reports it as gpu_launches). */
long long b200mpc_launch_count(const b200mpc_handle *h);
/* Device time [ms] of the most recent solve kernel measured with CUDA events on its own stream
 * (valid after the stream has been synchronised; host-buffer solves synchronise themselves). */
float b200mpc_last_kernel_ms(b200mpc_handle *h);

/* Measures the device's FP64 FMA throughput [TFLOP/s] with a register-resident DFMA loop on every SM.
 * bench.py uses it as the roofline denominator of the solve kernel (which is FP64-pipe / latency bound). */
int b200mpc_measure_fp64_peak(b200mpc_handle *h, double *tflops_out);

int b200mpc_abi_version(void);
/* sizeof(struct b200mpc_params) as compiled, so a binding can verify its struct layout. */
int b200mpc_sizeof_params(void);

#ifdef __cplusplus
}
#endif
#endif
