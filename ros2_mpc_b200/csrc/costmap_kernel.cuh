// costmap_kernel.cuh — costmap inflation and dilation for batches of grids.
//
// Replaces, for a batch of grids:
//   cv2.dilate(grid, np.ones((kh, kw)), iterations=1).astype(np.uint8)      ros2_mpc/core/local_costmap_publisher.py:34-35,
//                                                                            ros2_mpc/core/global_costmap_publisher.py (same call)
//   inflate_global / inflate_local                                           ros2_mpc/utils/costmap.py:5-41
//   and the whole loop body of the local costmap publisher, scan -> grid -> dilate -> uint8 image
//                                                                            ros2_mpc/core/local_costmap_publisher.py:29-35
//                                                                            with ros2_mpc/utils/utils.py:5-43 (rotation = yaw)
//
// dilate_kernel.  OpenCV's conventions for this call: anchor = kernel centre (kh/2, kw/2), constant border that never
// wins the maximum, so  out[y][x] = max src[y + i - kh/2][x + j - kw/2]  over 0 <= i < kh, 0 <= j < kw  inside the
// image; float64 in, uint8 out (the cast truncates toward zero; the grids hold 0 / 100).  One CTA per tile: the tile
// and its halo are staged in shared memory with coalesced loads, the maximum is separable (rows, then columns).
// HBM-bound: 8 bytes read and 1 byte written per cell.
//
// inflate_kernel.  The reference stamps np.minimum(window, inflation_matrix) around every cell whose ORIGINAL value is
// exactly 0 and whose (2c+1)^2 window lies completely inside the grid (a clipped window has another shape and is
// skipped).  min is commutative, so the result does not depend on the stamping order and each output cell can gather:
//   out[p] = min(in[p], min over d in [-c,c]^2 with q = p - d inside the border of c cells and in[q] == 0 of M[d + c]).
//
// local_costmap_kernel.  One warp per robot: the 80 x 80 occupancy grid is a bit set in shared memory (cell indices exactly
// as utils.py:33-41 computes them, see obstacles_kernel.cuh; here the scan is rotated by the robot's yaw first, with the
// arithmetic of the reference's np.dot: x' = fma(-sin, y, cos*x), y' = fma(cos, y, sin*x)), the dilation is done on the
// bits (OR of shifted rows, then OR of rows), and the uint8 image (0 / 100) leaves in 16-byte stores.
// HBM-bound: 8*n_beams bytes in, num_cells^2 bytes out per robot — the float64 grid never exists in memory.
#pragma once

struct DilateArgs {
    int B, H, W, kh, kw, TH, TW;
    const double *in;    // [B][H][W]
    unsigned char *out;  // [B][H][W]
};

#define COSTMAP_THREADS 256

__global__ void __launch_bounds__(COSTMAP_THREADS) dilate_kernel(const DilateArgs a) {
    extern __shared__ __align__(16) unsigned char cm_smem[];
    const int ah = a.kh / 2, aw = a.kw / 2;
    const int SH = a.TH + a.kh - 1, SW = a.TW + a.kw - 1;
    double *src = reinterpret_cast<double *>(cm_smem);        // [SH][SW]
    double *tmp = src + (size_t)SH * SW;                      // [SH][TW]  row maxima
    const int tiles_x = (a.W + a.TW - 1) / a.TW, tiles_y = (a.H + a.TH - 1) / a.TH;
    const int per_grid = tiles_x * tiles_y;
    const double lowest = -1.7976931348623157e308; // never wins (OpenCV's default border value for a dilation)
    for (long long t = blockIdx.x; t < (long long)a.B * per_grid; t += gridDim.x) {
        const int b = (int)(t / per_grid), tt = (int)(t % per_grid);
        const int y0 = (tt / tiles_x) * a.TH, x0 = (tt % tiles_x) * a.TW;
        const double *g = a.in + (size_t)b * a.H * a.W;
        for (int i = threadIdx.x; i < SH * SW; i += COSTMAP_THREADS) {
            const int r = i / SW, c = i - r * SW;
            const int y = y0 + r - ah, x = x0 + c - aw;
            src[i] = (y >= 0 && y < a.H && x >= 0 && x < a.W) ? __ldcs(g + (size_t)y * a.W + x) : lowest;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < SH * a.TW; i += COSTMAP_THREADS) {
            const int r = i / a.TW, c = i - r * a.TW;
            const double *s = src + (size_t)r * SW + c;
            double m = s[0];
            for (int j = 1; j < a.kw; j++) m = fmax(m, s[j]);
            tmp[i] = m;
        }
        __syncthreads();
        unsigned char *o = a.out + (size_t)b * a.H * a.W;
        for (int i = threadIdx.x; i < a.TH * a.TW; i += COSTMAP_THREADS) {
            const int r = i / a.TW, c = i - r * a.TW;
            const int y = y0 + r, x = x0 + c;
            if (y >= a.H || x >= a.W) continue;
            const double *s = tmp + (size_t)r * a.TW + c;
            double m = s[0];
            for (int k = 1; k < a.kh; k++) m = fmax(m, s[(size_t)k * a.TW]);
            o[(size_t)y * a.W + x] = (unsigned char)__double2int_rz(m);
        }
        __syncthreads();
    }
}

// Grids small enough for shared memory to hold the grid and its row maxima (the costmap publishers' 80 x 80 local grid:
// 2 x 51 KB, two CTAs per SM): one CTA per grid, the grid is read ONCE with 16-byte loads (no halo re-reads, no index
// arithmetic in the load), windows are clipped at the image border instead of padded.  Same results as dilate_kernel.
__global__ void __launch_bounds__(COSTMAP_THREADS) dilate_whole_kernel(const DilateArgs a) {
    extern __shared__ __align__(16) unsigned char cm_smem[];
    const int ah = a.kh / 2, aw = a.kw / 2, HW = a.H * a.W;
    double *src = reinterpret_cast<double *>(cm_smem); // [H][W]
    double *tmp = src + HW;                            // [H][W] row maxima
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        const double *g = a.in + (size_t)b * HW;
        if ((HW & 1) == 0) {
            const double2 *g2 = reinterpret_cast<const double2 *>(g);
            double2 *s2 = reinterpret_cast<double2 *>(src);
            for (int i = threadIdx.x; i < HW / 2; i += COSTMAP_THREADS) s2[i] = __ldcs(g2 + i);
        } else {
            for (int i = threadIdx.x; i < HW; i += COSTMAP_THREADS) src[i] = __ldcs(g + i);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < HW; i += COSTMAP_THREADS) {
            const int y = i / a.W, x = i - y * a.W;
            const int lo = max(0, x - aw), hi = min(a.W - 1, x - aw + a.kw - 1);
            const double *s = src + y * a.W;
            double m = s[lo];
            for (int j = lo + 1; j <= hi; j++) m = fmax(m, s[j]);
            tmp[i] = m;
        }
        __syncthreads();
        unsigned char *o = a.out + (size_t)b * HW;
        for (int i = threadIdx.x; i < HW; i += COSTMAP_THREADS) {
            const int y = i / a.W, x = i - y * a.W;
            const int lo = max(0, y - ah), hi = min(a.H - 1, y - ah + a.kh - 1);
            double m = tmp[lo * a.W + x];
            for (int k = lo + 1; k <= hi; k++) m = fmax(m, tmp[k * a.W + x]);
            o[i] = (unsigned char)__double2int_rz(m);
        }
        __syncthreads();
    }
}

struct InflateArgs {
    int B, H, W, c, TH, TW;
    const double *in;   // [B][H][W]
    const double *M;    // [(2c+1)][(2c+1)] inflation matrix
    double *out;        // [B][H][W]
};

__global__ void __launch_bounds__(COSTMAP_THREADS) inflate_kernel(const InflateArgs a) {
    extern __shared__ __align__(16) unsigned char cm_smem[];
    const int c = a.c, n = 2 * c + 1;
    const int SH = a.TH + 2 * c, SW = a.TW + 2 * c;
    double *src = reinterpret_cast<double *>(cm_smem);         // [SH][SW] original values (halo: 1.0 = not a source)
    double *Ms = src + (size_t)SH * SW;                        // [n][n]
    unsigned char *occ = reinterpret_cast<unsigned char *>(Ms + (size_t)n * n); // [SH][SW] 1 = stamping source
    const int tiles_x = (a.W + a.TW - 1) / a.TW, tiles_y = (a.H + a.TH - 1) / a.TH;
    const int per_grid = tiles_x * tiles_y;
    for (int i = threadIdx.x; i < n * n; i += COSTMAP_THREADS) Ms[i] = a.M[i];
    for (long long t = blockIdx.x; t < (long long)a.B * per_grid; t += gridDim.x) {
        const int b = (int)(t / per_grid), tt = (int)(t % per_grid);
        const int y0 = (tt / tiles_x) * a.TH, x0 = (tt % tiles_x) * a.TW;
        const double *g = a.in + (size_t)b * a.H * a.W;
        for (int i = threadIdx.x; i < SH * SW; i += COSTMAP_THREADS) {
            const int r = i / SW, cc = i - r * SW;
            const int y = y0 + r - c, x = x0 + cc - c;
            const bool inside = (y >= 0 && y < a.H && x >= 0 && x < a.W);
            const double v = inside ? __ldcs(g + (size_t)y * a.W + x) : 1.0;
            src[i] = v;
            // a source: original value exactly 0 and the whole window inside the grid (costmap.py:10-14)
            occ[i] = (inside && v == 0.0 && y >= c && y + c < a.H && x >= c && x + c < a.W) ? 1 : 0;
        }
        __syncthreads();
        double *o = a.out + (size_t)b * a.H * a.W;
        for (int i = threadIdx.x; i < a.TH * a.TW; i += COSTMAP_THREADS) {
            const int r = i / a.TW, cc = i - r * a.TW;
            const int y = y0 + r, x = x0 + cc;
            if (y >= a.H || x >= a.W) continue;
            double m = src[(size_t)(r + c) * SW + cc + c];
            // the source q = p - d sits at staged position (r + c - dy, cc + c - dx) and contributes M[dy + c][dx + c]
            for (int dy = -c; dy <= c; dy++) {
                const unsigned char *orow = occ + (size_t)(r + c - dy) * SW + cc + c;
                const double *mrow = Ms + (size_t)(dy + c) * n + c;
                for (int dx = -c; dx <= c; dx++)
                    if (orow[-dx]) m = fmin(m, mrow[dx]);
            }
            o[(size_t)y * a.W + x] = m;
        }
        __syncthreads();
    }
}

// ---- scan -> occupancy bits -> dilated uint8 image, one warp per robot ------------------------------------------------
struct LocalCostmapArgs {
    int B, n, nc, kh, kw, wpr;   // wpr = 32-bit words per grid row
    const double *scan;          // [B][n]
    const double *bcos, *bsin;   // [n] beam direction table (utils.py:18-20, computed once on the host)
    const double *yaw;           // [B] rotation argument (orientation[2])
    double half, res;            // map_size / 2, resolution
    unsigned char value;         // 100
    unsigned char *out;          // [B][nc][nc]
};

#define LCM_WARPS 8

// OR of the row shifted by -lo .. +hi bit positions (bit x of the result = any bit in [x - hi, x + lo] ... see caller)
__device__ __forceinline__ unsigned lcm_word(const unsigned *row, int wpr, int w) { return (w >= 0 && w < wpr) ? row[w] : 0u; }

__global__ void __launch_bounds__(LCM_WARPS * 32) local_costmap_kernel(const LocalCostmapArgs a) {
    extern __shared__ __align__(16) unsigned char cm_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int wpr = a.wpr, nc = a.nc;
    const size_t per_warp = (size_t)2 * nc * wpr * 4;
    unsigned *bits = reinterpret_cast<unsigned *>(cm_smem + wid * per_warp); // [nc][wpr] occupancy
    unsigned *hor = bits + (size_t)nc * wpr;                                  // [nc][wpr] horizontally dilated
    const double inv_res = 1.0 / a.res, fnc = (double)nc;
    const int ah = a.kh / 2, aw = a.kw / 2;
    // a set bit (ys, xs) lights the outputs y in [ys - (kh-1-ah), ys + ah], x in [xs - (kw-1-aw), xs + aw]
    const int left = a.kw - 1 - aw, right = aw, up = a.kh - 1 - ah, down = ah;
    const unsigned v4 = 0x01010101u * a.value;
    for (int b = blockIdx.x * LCM_WARPS + wid; b < a.B; b += gridDim.x * LCM_WARPS) {
        for (int w = lane; w < nc * wpr; w += 32) bits[w] = 0u;
        __syncwarp();
        const double *sc = a.scan + (size_t)b * a.n;
        double sn, cs;
        sincos(a.yaw[b], &sn, &cs);
        const double nsn = -sn;
        // pass 1: largest finite rotated coordinate per axis, needed only if an infinite one shows up (utils.py:30-31)
        int anyinf = 0;
        for (int i0 = lane; i0 < a.n; i0 += 128) {
            double r4[4];
#pragma unroll
            for (int u = 0; u < 4; u++) r4[u] = (i0 + 32 * u < a.n) ? __ldcs(sc + i0 + 32 * u) : 0.0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + 32 * u;
                if (i >= a.n) break;
                const double x = __dmul_rn(r4[u], a.bcos[i]), y = __dmul_rn(r4[u], a.bsin[i]);
                double xr = __fma_rn(nsn, y, __dmul_rn(cs, x)), yr = __fma_rn(cs, y, __dmul_rn(sn, x));
                xr = obs_fix(xr); yr = obs_fix(yr);
                if (isinf(xr) || isinf(yr)) { anyinf = 1; continue; }
                const unsigned ix = (unsigned)obs_cell_i(__dadd_rn(xr, a.half), a.res, inv_res);
                const unsigned iy = (unsigned)obs_cell_i(__dadd_rn(yr, a.half), a.res, inv_res);
                if (ix < (unsigned)nc && iy < (unsigned)nc) atomicOr(bits + iy * wpr + (ix >> 5), 1u << (ix & 31));
            }
        }
        if (__any_sync(FULL, anyinf)) {
            double mx = -INFINITY, my = -INFINITY;
            for (int i = lane; i < a.n; i += 32) {
                const double r = sc[i];
                const double x = __dmul_rn(r, a.bcos[i]), y = __dmul_rn(r, a.bsin[i]);
                const double xr = obs_fix(__fma_rn(nsn, y, __dmul_rn(cs, x))), yr = obs_fix(__fma_rn(cs, y, __dmul_rn(sn, x)));
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
                double xr = obs_fix(__fma_rn(nsn, y, __dmul_rn(cs, x))), yr = obs_fix(__fma_rn(cs, y, __dmul_rn(sn, x)));
                if (!isinf(xr) && !isinf(yr)) continue; // already marked
                if (isinf(xr)) xr = mx;
                if (isinf(yr)) yr = my;
                const double tx = trunc(__ddiv_rn(__dadd_rn(xr, a.half), a.res));
                const double ty = trunc(__ddiv_rn(__dadd_rn(yr, a.half), a.res));
                if (tx >= 0.0 && tx < fnc && ty >= 0.0 && ty < fnc)
                    atomicOr(bits + (int)ty * wpr + ((int)tx >> 5), 1u << ((int)tx & 31));
            }
        }
        __syncwarp();
        // horizontal pass: hor[y] bit x = OR of bits[y] over xs in [x - right, x + left]
        for (int i = lane; i < nc * wpr; i += 32) {
            const int y = i / wpr, w = i - y * wpr;
            const unsigned *row = bits + y * wpr;
            const unsigned lo = lcm_word(row, wpr, w - 1), mid = row[w], hi = lcm_word(row, wpr, w + 1);
            unsigned acc = mid;
            for (int s = 1; s <= right; s++) acc |= __funnelshift_l(lo, mid, s);  // sources at lower x
            for (int s = 1; s <= left; s++) acc |= __funnelshift_r(mid, hi, s);   // sources at higher x
            hor[i] = acc;
        }
        __syncwarp();
        // vertical pass + expansion to bytes: out[y] = OR of hor[ys] over ys in [y - down, y + up]
        unsigned char *o = a.out + (size_t)b * nc * nc;
        const int groups = (nc + 15) / 16; // 16 cells per store
        for (int i = lane; i < nc * groups; i += 32) {
            const int y = i / groups, gx = (i - y * groups) * 16;
            const int w = gx >> 5, sh = gx & 31;
            unsigned acc = 0u;
            const int ya = max(0, y - down), yb = min(nc - 1, y + up);
            for (int ys = ya; ys <= yb; ys++) acc |= hor[ys * wpr + w];
            const unsigned m16 = (acc >> sh) & 0xffffu;
            unsigned q[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const unsigned nib = (m16 >> (4 * u)) & 0xfu;
                // spread 4 bits to 4 bytes (0x00 / 0x01 each), then scale by the cell value
                const unsigned spread = (nib * 0x00204081u) & 0x01010101u;
                q[u] = spread * a.value;
            }
            (void)v4;
            unsigned char *dst = o + (size_t)y * nc + gx;
            if (gx + 16 <= nc && ((reinterpret_cast<size_t>(dst) & 15) == 0)) {
                __stcs(reinterpret_cast<uint4 *>(dst), make_uint4(q[0], q[1], q[2], q[3]));
            } else {
                for (int u = 0; u < 16 && gx + u < nc; u++) dst[u] = (unsigned char)((q[u >> 2] >> (8 * (u & 3))) & 0xffu);
            }
        }
        __syncwarp();
    }
}
