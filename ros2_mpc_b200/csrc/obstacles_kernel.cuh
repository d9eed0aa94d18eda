// obstacles_kernel.cuh — batched obstacle-list construction: laser scan -> local occupancy grid -> fixed-length list
// of obstacle points in world coordinates (the obstacles_x / obstacles_y parameters of the MPC).
//
// Replaces, for a batch of robots, the producer that runs right before every solve in the reference:
//   get_obstacles                          ros2_mpc/scripts/point_follower_local_planner.py:88-118
//   convert_laser_scan_to_occupancy_grid   ros2_mpc/utils/utils.py:5-43
//   convert_to_map_coordinates             ros2_mpc/utils/utils.py:114-124
//   rotate_coordinates                     ros2_mpc/utils/utils.py:145-152
//
// One warp per robot.  The occupancy grid (num_cells x num_cells, 80 x 80 for params.yaml) is a bit set in shared
// memory, stored directly in the order np.where() walks the 180-degree-rotated grid (np.rot90(k=2): rotated
// row-major position q = num_cells^2 - 1 - (y*num_cells + x)), so "the first `slots` obstacle cells" are the first
// `slots` set bits.  Ranks come from per-lane popcounts and a warp prefix sum; the list is assembled in shared
// memory and written out with coalesced stores.  HBM-bound: 8*n_beams bytes in, 16*slots + 4 bytes out per robot.
//
// Cell indexing is bit-exact with the reference: the beam direction table cos/sin(i*(max-min)/n + min) is computed
// once on the host (it is a property of the lidar, not of the robot), and every device operation that feeds an index
// is a single correctly-rounded IEEE operation (__dmul_rn / __dadd_rn / __ddiv_rn: no FMA contraction).  The quirks
// of the reference are kept: the rotation-by-0.0 matrix product turns +-inf coordinates into NaN, NaN becomes 0,
// int() truncates toward zero, more than `slots` cells are truncated (the raw count is returned so that a caller
// can raise like the reference does), no cell at all yields the sentinel 100.0.
#pragma once

struct ObsBuildArgs {
    int B, n, nc, slots, nwords;
    const double *scan;      // [B][n]
    const double *bcos, *bsin; // [n] beam direction table
    const double *pos;       // [B][2]
    const double *yaw;       // [B]
    double half, res, origin; // map_size/2, resolution, (nc/2)*resolution
    double *ox, *oy;         // [B][slots]
    int *count;              // [B] raw number of occupied cells
};

#define OBS_WARPS 8

__device__ __forceinline__ double obs_fix(double v) { return (v != v) ? 0.0 : v; }

// trunc(v / res), the reference's int(x_indices[i] / cell_size) (utils.py:39), without paying for an IEEE division per
// beam: q = v * (1/res) is within 4.5e-16 |q| of the correctly rounded quotient, so the two can only truncate
// differently when q lies that close to an integer — only then is the division carried out.  Bit-exact.
__device__ __forceinline__ double obs_cell(double v, double res, double inv_res) {
    const double q = v * inv_res;
    if (fabs(q - rint(q)) <= 1e-15 * fabs(q)) return trunc(__ddiv_rn(v, res));
    return trunc(q);
}
// the same as a cell number: truncation toward zero is the conversion's rounding mode, the conversion saturates, and
// one unsigned compare then tests 0 <= idx < num_cells ((-0.05, 0) -> cell 0, as int() does)
__device__ __forceinline__ int obs_cell_i(double v, double res, double inv_res) {
    const double q = v * inv_res;
    int iq = __double2int_rz(q);
    // distance of q from the integers on either side of it (rint() costs ~20 instructions on this architecture)
    const double f = fabs(q - (double)iq), tol = 1e-15 * fabs(q);
    if (f <= tol || f >= 1.0 - tol) iq = __double2int_rz(__ddiv_rn(v, res));
    return iq;
}

__global__ void __launch_bounds__(OBS_WARPS * 32) obstacles_kernel(const ObsBuildArgs a) {
    extern __shared__ __align__(16) unsigned char obs_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t per_warp = ((size_t)a.nwords * 8 + (size_t)a.slots * 4 + 15) & ~(size_t)15;
    unsigned *bits = reinterpret_cast<unsigned *>(obs_smem + wid * per_warp);
    int *wpre = reinterpret_cast<int *>(bits + a.nwords);  // set bits in front of each word
    int *qlist = wpre + a.nwords;                          // positions of the first `slots` set bits, in order
    const int wpl = (a.nwords + 31) / 32; // words per lane (contiguous, so that lane order = bit order)
    const int nbits = a.nc * a.nc;
    const double inv_res = 1.0 / a.res, fnc = (double)a.nc;
    const unsigned magic = (unsigned)((0x100000000ull + (unsigned)a.nc - 1) / (unsigned)a.nc); // q / nc = umulhi(q, magic), q < 2^16
    for (int b = blockIdx.x * OBS_WARPS + wid; b < a.B; b += gridDim.x * OBS_WARPS) {
        for (int w = lane; w < a.nwords; w += 32) bits[w] = 0u;
        __syncwarp();
        const double *sc = a.scan + (size_t)b * a.n;
        {   // the warp's next scan on its way into L2 while this one is processed (one 128-byte line per lane)
            const long long bn = (long long)b + (long long)gridDim.x * OBS_WARPS;
            if (bn < a.B && lane * 16 < a.n) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.scan + (size_t)bn * a.n + lane * 16));
        }
        // ---- cell indices, occupancy bits in rotated np.where order ----
        // (an infinite coordinate would have to be replaced by the scan's largest finite one, utils.py:30-31: that
        //  needs a second pass, taken only if one shows up — the rotation-by-0.0 product turns the infinities of
        //  infinite ranges into NaN, so with the reference's call it never does)
        int anyinf = 0;
        // four beams per lane in flight (the ranges are the only HBM reads of the kernel)
        for (int i0 = lane; i0 < a.n; i0 += 128) {
            double r4[4];
#pragma unroll
            for (int u = 0; u < 4; u++) r4[u] = (i0 + 32 * u < a.n) ? __ldcs(sc + i0 + 32 * u) : 0.0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + 32 * u;
                if (i >= a.n) break;
                const double r = r4[u];
                double xr, yr;
                if (fabs(r) < 1.0e300) {
                    // finite range: the rotation by 0.0 ([[1, -0], [0, 1]] @ [x; y]) returns x and y unchanged
                    xr = __dmul_rn(r, a.bcos[i]); yr = __dmul_rn(r, a.bsin[i]);
                } else {
                    const double x = __dmul_rn(r, a.bcos[i]), y = __dmul_rn(r, a.bsin[i]);
                    xr = obs_fix(__dadd_rn(x, __dmul_rn(-0.0, y)));
                    yr = obs_fix(__dadd_rn(__dmul_rn(0.0, x), y));
                    if (isinf(xr) || isinf(yr)) { anyinf = 1; continue; }
                }
                const unsigned ix = (unsigned)obs_cell_i(__dadd_rn(xr, a.half), a.res, inv_res);
                const unsigned iy = (unsigned)obs_cell_i(__dadd_rn(yr, a.half), a.res, inv_res);
                if (ix < (unsigned)a.nc && iy < (unsigned)a.nc) {
                    const int q = nbits - 1 - (int)(iy * a.nc + ix);
                    atomicOr(bits + (q >> 5), 1u << (q & 31));
                }
            }
        }
        if (__any_sync(FULL, anyinf)) {
            double mx = -INFINITY, my = -INFINITY;
            for (int i = lane; i < a.n; i += 32) {
                const double r = sc[i];
                const double x = __dmul_rn(r, a.bcos[i]), y = __dmul_rn(r, a.bsin[i]);
                const double xr = obs_fix(__dadd_rn(x, __dmul_rn(-0.0, y)));
                const double yr = obs_fix(__dadd_rn(__dmul_rn(0.0, x), y));
                if (!isinf(xr)) mx = fmax(mx, xr);
                if (!isinf(yr)) my = fmax(my, yr);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
                my = fmax(my, __shfl_xor_sync(FULL, my, o));
            }
            for (int i = lane; i < a.n; i += 32) {
                const double r = sc[i];
                const double x = __dmul_rn(r, a.bcos[i]), y = __dmul_rn(r, a.bsin[i]);
                double xr = obs_fix(__dadd_rn(x, __dmul_rn(-0.0, y)));
                double yr = obs_fix(__dadd_rn(__dmul_rn(0.0, x), y));
                if (!isinf(xr) && !isinf(yr)) continue; // already marked
                if (isinf(xr)) xr = mx;
                if (isinf(yr)) yr = my;
                const double tx = trunc(__ddiv_rn(__dadd_rn(xr, a.half), a.res));
                const double ty = trunc(__ddiv_rn(__dadd_rn(yr, a.half), a.res));
                if (tx >= 0.0 && tx < fnc && ty >= 0.0 && ty < fnc) {
                    const int q = nbits - 1 - ((int)ty * a.nc + (int)tx);
                    atomicOr(bits + (q >> 5), 1u << (q & 31));
                }
            }
        }
        __syncwarp();
        // ---- ranks: per-lane popcount over a contiguous word range, exclusive warp prefix sum; per-word prefix to
        //      shared memory ----
        const int w0 = lane * wpl;
        int mine = 0;
        for (int t = 0; t < wpl; t++)
            if (w0 + t < a.nwords) mine += __popc(bits[w0 + t]);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        {
            int run = incl - mine;
            for (int t = 0; t < wpl; t++)
                if (w0 + t < a.nwords) { wpre[w0 + t] = run; run += __popc(bits[w0 + t]); }
        }
        __syncwarp();
        // ---- positions of the first `slots` set bits, in order; the words are dealt round-robin so that the bits of
        //      a wall (consecutive words) spread over the lanes ----
        for (int w = lane; w < a.nwords; w += 32) {
            int rank = wpre[w];
            unsigned m = bits[w];
            while (m && rank < a.slots) {
                qlist[rank++] = (w << 5) + (__ffs(m) - 1);
                m &= m - 1;
            }
        }
        __syncwarp();
        // ---- world coordinates, one slot per lane (balanced), padding, coalesced streaming stores ----
        const int nout = min(total, a.slots);
        double sn = 0.0, cs = 1.0, px = 0.0, py = 0.0;
        if (nout > 0) {
            sincos(a.yaw[b], &sn, &cs);
            px = a.pos[2 * (size_t)b]; py = a.pos[2 * (size_t)b + 1];
        }
        double *gx = a.ox + (size_t)b * a.slots, *gy = a.oy + (size_t)b * a.slots;
        double fx = 100.0, fy = 100.0; // sentinel of the "No obstacles" branch
        for (int s0 = 0; s0 < a.slots; s0 += 32) {
            const int sl = s0 + lane;
            double wx = 0.0, wy = 0.0;
            if (sl < nout) {
                const int q = qlist[sl];
                const int i = (int)__umulhi((unsigned)q, magic), j = q - i * a.nc;
                // convert_to_map_coordinates: x = -i*res + origin, y = -j*res + origin
                const double cx = __dadd_rn(__dmul_rn(-(double)i, a.res), a.origin);
                const double cy = __dadd_rn(__dmul_rn(-(double)j, a.res), a.origin);
                wx = __dadd_rn(__dadd_rn(__dmul_rn(cs, cx), __dmul_rn(-sn, cy)), px);
                wy = __dadd_rn(__dadd_rn(__dmul_rn(sn, cx), __dmul_rn(cs, cy)), py);
            }
            if (s0 == 0 && nout > 0) { fx = __shfl_sync(FULL, wx, 0); fy = __shfl_sync(FULL, wy, 0); } // padding: first obstacle
            if (sl < a.slots) {
                __stcs(gx + sl, (sl < nout) ? wx : fx);
                __stcs(gy + sl, (sl < nout) ? wy : fy);
            }
        }
        if (lane == 0 && a.count) a.count[b] = total;
        __syncwarp();
    }
}
