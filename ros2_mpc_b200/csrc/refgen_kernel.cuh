// refgen_kernel.cuh — batched producers of the solve's reference arguments.
//
//   goals_kernel     <- get_goal_for_mpc          ros2_mpc/scripts/point_follower_local_planner.py:16-30
//                       (look-ahead goal of the point-stabilisation planner: final_state of perform_mpc)
//   reftraj_kernel   <- get_reference_trajectory  ros2_mpc/scripts/path_follower_local_planner.py:27-73
//                       (pxf / puf of the tracking planner: pf, puf of perform_mpc)
//
// One warp per robot: the distances to the K path points are evaluated lane-strided, "first index beyond the
// look-ahead distance" and "nearest index" are warp reductions with numpy's tie-breaking (lowest index).  Distances
// are formed exactly as numpy forms them — (dx*dx + dy*dy) with separately rounded products, then sqrt — and the
// heading wrap is Python's float modulo, so indices and outputs are bit-exact with the reference.
// Paths may be shared by the whole batch (path_stride 0) or per robot (path_stride = K).
#pragma once

struct RefGenArgs {
    int B, K, N, Ko;          // robots, path points, horizon, len(path_omega)
    long long path_stride;    // 0: one path for all robots; K: per-robot paths
    const double *path_xy;    // [.][K][2]
    const double *heading;    // [.][K]
    const double *velocity;   // [.][K]   (reftraj only)
    const double *omega;      // [.][Ko]  (reftraj only)
    const double *goal;       // [B][goal_stride]
    int goal_stride;
    const double *pos;        // [B][pos_stride]: (x, y, ...)
    int pos_stride;
    double lookahead;
    double *out_goal;         // [B][3]            (goals)
    double *pxf, *puf;        // [B][3N], [B][2N]  (reftraj)
    int *nearest;             // [B] (may be NULL): chosen path index
};

#define REFGEN_WARPS 8

// Python's a % m for floats (numpy remainder): fmod, then shifted into the sign of the divisor
__device__ __forceinline__ double refgen_pymod(double a, double m) {
    double r = fmod(a, m);
    if (r != 0.0) { if ((m < 0.0) != (r < 0.0)) r += m; }
    else r = copysign(0.0, m);
    return r;
}
// np.linalg.norm of a 2-vector: sqrt(dx*dx + dy*dy), products rounded separately
__device__ __forceinline__ double refgen_norm2(double dx, double dy) {
    return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}
// np.linalg.norm of ONE 2-vector (no axis argument): sqrt(x.dot(x)), and the BLAS dot product fuses the second product
// into the sum: sqrt(fma(dy, dy, dx*dx)) — one rounding less than the axis=1 form above (measured against numpy on 20 000
// vectors; it decides comparisons that land within an ulp of a threshold)
__device__ __forceinline__ double refgen_norm2_dot(double dx, double dy) {
    return sqrt(__fma_rn(dy, dy, __dmul_rn(dx, dx)));
}
// np.argmin ordering: smaller value wins, a NaN beats everything, ties go to the lower index
__device__ __forceinline__ bool refgen_better(double a, int ia, double b, int ib) {
    const bool an = (a != a), bn = (b != b);
    if (an != bn) return an;
    if (!an && a != b) return a < b;
    return ia < ib;
}
__device__ __forceinline__ void refgen_argmin(double &v, int &i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(FULL, v, o);
        const int oi = __shfl_xor_sync(FULL, i, o);
        if (refgen_better(ov, oi, v, i)) { v = ov; i = oi; }
    }
}

__global__ void __launch_bounds__(REFGEN_WARPS * 32) goals_kernel(const RefGenArgs a) {
    const int lane = threadIdx.x & 31;
    const double two_pi = 2.0 * 3.141592653589793;
    for (int b = blockIdx.x * REFGEN_WARPS + (threadIdx.x >> 5); b < a.B; b += gridDim.x * REFGEN_WARPS) {
        const double *g = a.goal + (size_t)b * a.goal_stride;
        const double px = a.pos[(size_t)b * a.pos_stride], py = a.pos[(size_t)b * a.pos_stride + 1];
        double *out = a.out_goal + (size_t)b * 3;
        if (refgen_norm2_dot(g[0] - px, g[1] - py) < a.lookahead) {
            if (lane == 0) { out[0] = g[0]; out[1] = g[1]; out[2] = refgen_pymod(g[4], two_pi); }
            if (lane == 0 && a.nearest) a.nearest[b] = -1;
            continue;
        }
        const double *pxy = a.path_xy + (size_t)b * a.path_stride * 2;
        const double *ph = a.heading + (size_t)b * a.path_stride;
        int first = 0x7fffffff, imin = 0x7fffffff;
        double vmin = INFINITY;
        for (int k = lane; k < a.K; k += 32) {
            const double d = refgen_norm2(pxy[2 * k] - px, pxy[2 * k + 1] - py);
            if (d > a.lookahead && k < first) first = k;
            if (imin == 0x7fffffff || refgen_better(d, k, vmin, imin)) { vmin = d; imin = k; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(FULL, first, o));
        refgen_argmin(vmin, imin);
        const int idx = (first != 0x7fffffff) ? first : imin;
        if (lane == 0) {
            out[0] = pxy[2 * idx]; out[1] = pxy[2 * idx + 1]; out[2] = refgen_pymod(ph[idx], two_pi);
            if (a.nearest) a.nearest[b] = idx;
        }
    }
}

__global__ void __launch_bounds__(REFGEN_WARPS * 32) reftraj_kernel(const RefGenArgs a) {
    const int lane = threadIdx.x & 31;
    for (int b = blockIdx.x * REFGEN_WARPS + (threadIdx.x >> 5); b < a.B; b += gridDim.x * REFGEN_WARPS) {
        const double *pxy = a.path_xy + (size_t)b * a.path_stride * 2;
        const double *ph = a.heading + (size_t)b * a.path_stride;
        const double *pv = a.velocity + (size_t)b * a.path_stride;
        const double *pw = a.omega + (size_t)b * (a.path_stride ? a.Ko : 0);
        const double x = a.pos[(size_t)b * a.pos_stride], y = a.pos[(size_t)b * a.pos_stride + 1];
        int imin = 0x7fffffff;
        double vmin = INFINITY;
        for (int k = lane; k < a.K; k += 32) {
            const double d = refgen_norm2(x - pxy[2 * k], y - pxy[2 * k + 1]);
            if (imin == 0x7fffffff || refgen_better(d, k, vmin, imin)) { vmin = d; imin = k; }
        }
        refgen_argmin(vmin, imin);
        const bool at_end = refgen_norm2_dot(x - pxy[2 * (a.K - 1)], y - pxy[2 * (a.K - 1) + 1]) < 0.5;
        const double *g = a.goal + (size_t)b * a.goal_stride;
        double *pxf = a.pxf + (size_t)b * 3 * a.N, *puf = a.puf + (size_t)b * 2 * a.N;
        for (int i = lane; i < a.N; i += 32) {
            const int k = min(imin + i, a.K - 1); // the reference pads every array with its last element
            if (at_end) { pxf[3 * i] = g[0]; pxf[3 * i + 1] = g[1]; pxf[3 * i + 2] = g[2]; }
            else { pxf[3 * i] = pxy[2 * k]; pxf[3 * i + 1] = pxy[2 * k + 1]; pxf[3 * i + 2] = ph[k]; }
            puf[2 * i] = pv[k];
            puf[2 * i + 1] = pw[min(k, a.Ko - 1)];
        }
        if (lane == 0 && a.nearest) a.nearest[b] = imin;
    }
}
