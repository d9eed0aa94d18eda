// sensor_kernel.cuh — producers that sit in front of the obstacle-list kernel and the reference kernels.
//
//   raycast_kernel    laser scans of a SHARED static occupancy map for a batch of robot poses (SURVEY.md section 8 row f1,
//                     second half; section 8e "shared read-only map replicated per GPU").  The map is the reference's
//                     maps/map_carto.pgm with the pixel convention of ros2_mpc/core/map_server.py:14-20 (occupied = pixel 0,
//                     flipped so that row 0 is the lowest y; resolution / origin from maps/map_carto.yaml:1-7).  Beam i of
//                     a robot points along yaw + i*(angle_max-angle_min)/n + angle_min (the beam convention of
//                     ros2_mpc/utils/utils.py:18) and is sampled every `step` metres from range_min on; the first sample
//                     inside an occupied cell is the range, range_max if there is none.  This is the sensor model of
//                     ros2_mpc_b200/synth.py:raycast (host numpy, used to build the test workloads), operation for
//                     operation: separately rounded products and sums, IEEE division, floor.
//                     The occupancy bits of the whole map (314 x 224 cells = 8.8 KB) are staged ONCE per CTA in shared
//                     memory; a warp takes one robot, a lane every 32nd beam.  The closed loop of an obstacle-active fleet
//                     then runs scan -> obstacle list -> solve -> control step without leaving the device.
//   headings_kernel   <- get_headings   ros2_mpc/scripts/path_follower_local_planner.py:14-23: per path, heading =
//                     arctan2 of the segment, repeated at the end; omega = half the heading difference; velocity =
//                     2 * segment length / dt, repeated at the end.  One thread per path point.
#pragma once

struct RaycastArgs {
    int B, n, H, W, nsteps, wpr;   // wpr = 32-bit words per map row
    const unsigned *occ_bits;      // [H][wpr] occupancy bits of the shared map (bit c of row r = cell (r, c) occupied)
    const double *pose;            // [B][pose_stride]: x, y, yaw
    int pose_stride;
    double angle_min, angle_inc;   // beam i: yaw + (i*angle_inc/n + angle_min), angle_inc = angle_max - angle_min
    double range_min, range_max, step;
    double ox, oy, res;            // map origin and resolution
    double *scan;                  // [B][n]
};

#define RAY_WARPS 8

// floor(v / res) exactly as numpy computes np.floor(v / res), without an IEEE division per sample: q = v * (1/res) is
// within a few ulp of the quotient, so the two floors can only differ when q is that close to an integer.
__device__ __forceinline__ int ray_cell(double v, double res, double inv_res) {
    const double q = v * inv_res;
    const double fq = floor(q);
    const double f = q - fq, tol = 1e-14 * fabs(q) + 1e-300;
    if (f <= tol || f >= 1.0 - tol) return (int)floor(__ddiv_rn(v, res));
    return (int)fq;
}

__global__ void __launch_bounds__(RAY_WARPS * 32) raycast_kernel(const RaycastArgs a) {
    extern __shared__ __align__(16) unsigned char ray_smem[];
    unsigned *occ = reinterpret_cast<unsigned *>(ray_smem);
    for (int i = threadIdx.x; i < a.H * a.wpr; i += blockDim.x) occ[i] = a.occ_bits[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double inv_res = 1.0 / a.res;
    for (int b = blockIdx.x * RAY_WARPS + wid; b < a.B; b += gridDim.x * RAY_WARPS) {
        const double *ps = a.pose + (size_t)b * a.pose_stride;
        const double px = ps[0], py = ps[1], yaw = ps[2];
        double *out = a.scan + (size_t)b * a.n;
        for (int i = lane; i < a.n; i += 32) {
            // ang = yaw + (i * (angle_max - angle_min) / n + angle_min), evaluated left to right as numpy does
            const double ang = __dadd_rn(yaw, __dadd_rn(__ddiv_rn(__dmul_rn((double)i, a.angle_inc), (double)a.n), a.angle_min));
            double sa, ca;
            sincos(ang, &sa, &ca);
            double rng = a.range_max;
            for (int t = 0; t < a.nsteps; t++) {
                const double r = __dadd_rn(a.range_min, __dmul_rn((double)t, a.step));
                const double vx = __dadd_rn(__dadd_rn(px, __dmul_rn(r, ca)), -a.ox);
                const double vy = __dadd_rn(__dadd_rn(py, __dmul_rn(r, sa)), -a.oy);
                const int col = ray_cell(vx, a.res, inv_res), row = ray_cell(vy, a.res, inv_res);
                if ((unsigned)row < (unsigned)a.H && (unsigned)col < (unsigned)a.W) {
                    if ((occ[row * a.wpr + (col >> 5)] >> (col & 31)) & 1u) { rng = r; break; }
                }
            }
            out[i] = rng;
        }
    }
}

struct HeadingsArgs {
    int P, K;               // paths, points per path
    const double *path_xy;  // [P][K][2]
    double dt;
    double *heading;        // [P][K]
    double *velocity;       // [P][K]
    double *omega;          // [P][K-1]
};

__global__ void __launch_bounds__(256) headings_kernel(const HeadingsArgs a) {
    const long long total = (long long)a.P * a.K;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(idx / a.K), k = (int)(idx - (long long)p * a.K);
        const double *xy = a.path_xy + (size_t)p * a.K * 2;
        // segment s = min(k, K-2): the last point repeats the last segment's values (np.append(..., [-1]))
        auto seg_heading = [&](int s) { return atan2(xy[2 * (s + 1) + 1] - xy[2 * s + 1], xy[2 * (s + 1)] - xy[2 * s]); };
        const int s = (k < a.K - 1) ? k : a.K - 2;
        const double h = seg_heading(s);
        a.heading[(size_t)p * a.K + k] = h;
        const double dx = xy[2 * (s + 1)] - xy[2 * s], dy = xy[2 * (s + 1) + 1] - xy[2 * s + 1];
        // (np.linalg.norm(diff, axis=1) / dt) * 2: products rounded separately, then sqrt
        const double nrm = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        a.velocity[(size_t)p * a.K + k] = __dmul_rn(__ddiv_rn(nrm, a.dt), 2.0);
        if (k < a.K - 1) {
            const int s1 = (k + 1 < a.K - 1) ? k + 1 : a.K - 2;
            a.omega[(size_t)p * (a.K - 1) + k] = __ddiv_rn(__dadd_rn(seg_heading(s1), -h), 2.0);
        }
    }
}
