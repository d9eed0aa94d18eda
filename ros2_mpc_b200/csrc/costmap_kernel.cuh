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

// dilate_strip_kernel<KH, KW> — the dilation for a compile-time structuring element (the reference only ever uses
// ones((10,10)): both costmap publishers).  The two kernels above spend their time on shared-memory loads and FP64 maxima:
// kw + kh of each per cell (8.6 % of the HBM peak on 80 x 80 grids).  Here
//   * cells are converted to int32 once, while they are staged (trunc toward zero is monotone, so it commutes with the
//     maximum; a NaN never won an fmax and becomes INT_MIN, whose low byte is the 0 that the cast of a NaN produced),
//   * a thread takes a STRIP of DIL_SEG consecutive outputs of a row (then of a column), reads the strip and its halo into
//     registers and forms the window maxima by doubling (windows of 2, 4, 8, then 8 + 8 overlapping = 10): 96 integer
//     maxima and 29 + 20 shared-memory accesses per 20 cells instead of 180 + 220,
//   * row strides are odd, so the strips of a warp's 32 threads start in 32 different banks,
//   * the bytes leave through a shared-memory image in 16-byte stores.
// One CTA per tile (tile = whole grid when it fits: the 80 x 80 local costmap is exactly 320 strips per pass).
#define DIL_THREADS 320
#define DIL_SEG 20
#define DIL_LD 1
#define DIL_NEG (-2147483647 - 1)

__device__ __forceinline__ int dil_cvt(double v) { return (v != v) ? DIL_NEG : __double2int_rz(v); }

// v[t] <- max(v[t .. t+K-1]) for t < NOUT (in place; entries from NOUT on are scratch)
template <int K, int NOUT>
__device__ __forceinline__ void dil_window_max(int (&v)[NOUT + K - 1]) {
    constexpr int LEN = NOUT + K - 1;
    constexpr int P = K >= 32 ? 32 : K >= 16 ? 16 : K >= 8 ? 8 : K >= 4 ? 4 : K >= 2 ? 2 : 1; // largest power of two <= K
#pragma unroll
    for (int q = 1; q < P; q <<= 1) {
#pragma unroll
        for (int j = 0; j < LEN; j++)
            if (j + 2 * q <= LEN) v[j] = max(v[j], v[j + q]);
    }
    if (K > P) {
#pragma unroll
        for (int t = 0; t < NOUT; t++) v[t] = max(v[t], v[t + K - P]);
    }
}

struct DilateStripArgs {
    int B, H, W, TH, TW;   // tile size (TH == H and TW == W: whole-grid mode)
    int PS, PT;            // row strides (ints) of the staged tile and of the row maxima, both odd
    int src_rows;          // rows of the staged tile
    int obuf_off;          // byte offset of the output image in shared memory (multiple of 16)
    int raw_bytes, bar_off; // TMA instance: bytes of the raw float64 grid in front of the int32 tile, offset of the mbarrier
    const double *in;      // [B][H][W]
    unsigned char *out;    // [B][H][W]
};

template <int KH, int KW>
__global__ void __launch_bounds__(DIL_THREADS, 3) dilate_strip_kernel(const DilateStripArgs a) {
    extern __shared__ __align__(16) unsigned char cm_smem[];
    constexpr int ah = KH / 2, aw = KW / 2;
    const int H = a.H, W = a.W, TH = a.TH, TW = a.TW, PS = a.PS, PT = a.PT;
    const int NSX = (TW + DIL_SEG - 1) / DIL_SEG, NSY = (TH + DIL_SEG - 1) / DIL_SEG;
    const int SH = TH + KH - 1, CW = TW + KW - 1;
    int *src = reinterpret_cast<int *>(cm_smem);                 // [src_rows][PS]  column c <-> image x = x0 - aw + c
    int *tmp = src + (size_t)a.src_rows * PS;                    // [NSY*SEG + KH - 1][PT]  row r <-> image y = y0 - ah + r
    unsigned char *obuf = cm_smem + a.obuf_off;                  // [TH][TW]
    const bool whole = (TH == H && TW == W);
    const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH, per_grid = tiles_x * tiles_y;
    const bool vec_in = whole && (W & 1) == 0 && ((reinterpret_cast<size_t>(a.in) & 15) == 0);
    if (whole) {
        // the halo never changes from grid to grid: columns outside the image, rows of maxima outside the image
        for (int i = threadIdx.x; i < H * (KW - 1); i += DIL_THREADS) {
            const int r = i / (KW - 1), j = i - r * (KW - 1);
            src[r * PS + (j < aw ? j : W + j)] = DIL_NEG;
        }
        for (int i = threadIdx.x; i < (KH - 1) * TW; i += DIL_THREADS) {
            const int j = i / TW, x = i - j * TW;
            tmp[(j < ah ? j : H + j) * PT + x] = DIL_NEG;
        }
    }
    // (loop-invariant index arithmetic of the vectorised staging: this thread's first cell, the step between its loads)
    const int half = (H * W) >> 1;
    const int dy = (2 * DIL_THREADS) / W, dx = 2 * DIL_THREADS - dy * W;
    const int ys = (2 * (int)threadIdx.x) / W, xs = 2 * (int)threadIdx.x - ys * W;
    const int ntiles = a.B * per_grid; // (the host keeps B * tiles per grid below 2^31)
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int b = whole ? t : t / per_grid, tt = whole ? 0 : t - b * per_grid;
        const int ty = whole ? 0 : tt / tiles_x;
        const int y0 = ty * TH, x0 = (tt - ty * tiles_x) * TW;
        const double *g = a.in + (size_t)b * H * W;
        // staged rows that lie inside the image: r in [r_lo, r_hi)
        const int r_lo = max(0, ah - y0), r_hi = min(SH, H - y0 + ah), nr = r_hi - r_lo;
        if (vec_in) {
            const double2 *g2 = reinterpret_cast<const double2 *>(g);
            int y = ys, x = xs;
            // DIL_LD loads in flight per thread before the first conversion waits for one
            for (int i0 = threadIdx.x; i0 < half; i0 += DIL_LD * DIL_THREADS) {
                double2 v[DIL_LD];
#pragma unroll
                for (int u = 0; u < DIL_LD; u++)
                    if (i0 + u * DIL_THREADS < half) v[u] = __ldcs(g2 + i0 + u * DIL_THREADS);
#pragma unroll
                for (int u = 0; u < DIL_LD; u++) {
                    if (i0 + u * DIL_THREADS < half) {
                        int *d = src + y * PS + aw + x;
                        d[0] = dil_cvt(v[u].x); d[1] = dil_cvt(v[u].y);
                    }
                    x += dx; y += dy;
                    if (x >= W) { x -= W; y++; }
                }
            }
        } else {
            for (int i = threadIdx.x; i < nr * CW; i += DIL_THREADS) {
                const int rr = i / CW, c = i - rr * CW;
                const int y = y0 - ah + r_lo + rr, x = x0 - aw + c;
                src[rr * PS + c] = (x >= 0 && x < W) ? dil_cvt(__ldcs(g + (size_t)y * W + x)) : DIL_NEG;
            }
            for (int i = threadIdx.x; i < (SH - nr) * TW; i += DIL_THREADS) {
                const int j = i / TW, x = i - j * TW;
                tmp[(j < r_lo ? j : nr + j) * PT + x] = DIL_NEG;
            }
        }
        __syncthreads();
        // rows: one strip of DIL_SEG outputs per thread; consecutive threads take consecutive rows (odd stride PS)
        for (int i = threadIdx.x; i < nr * NSX; i += DIL_THREADS) {
            const int s = i / nr, rr = i - s * nr;
            const int *p = src + rr * PS + s * DIL_SEG;
            int v[DIL_SEG + KW - 1];
#pragma unroll
            for (int j = 0; j < DIL_SEG + KW - 1; j++) v[j] = p[j];
            dil_window_max<KW, DIL_SEG>(v);
            int *q = tmp + (r_lo + rr) * PT + s * DIL_SEG;
            if ((s + 1) * DIL_SEG <= TW) {
#pragma unroll
                for (int j = 0; j < DIL_SEG; j++) q[j] = v[j];
            } else {
#pragma unroll
                for (int j = 0; j < DIL_SEG; j++)
                    if (s * DIL_SEG + j < TW) q[j] = v[j];
            }
        }
        __syncthreads();
        // columns: consecutive threads take consecutive columns
        for (int i = threadIdx.x; i < TW * NSY; i += DIL_THREADS) {
            const int s = i / TW, x = i - s * TW;
            const int *p = tmp + (s * DIL_SEG) * PT + x;
            int v[DIL_SEG + KH - 1];
#pragma unroll
            for (int j = 0; j < DIL_SEG + KH - 1; j++) v[j] = p[j * PT];
            dil_window_max<KH, DIL_SEG>(v);
            unsigned char *ob = obuf + (s * DIL_SEG) * TW + x;
            if ((s + 1) * DIL_SEG <= TH) {
#pragma unroll
                for (int j = 0; j < DIL_SEG; j++) ob[j * TW] = (unsigned char)v[j];
            } else {
#pragma unroll
                for (int j = 0; j < DIL_SEG; j++)
                    if (s * DIL_SEG + j < TH) ob[j * TW] = (unsigned char)v[j];
            }
        }
        __syncthreads();
        unsigned char *o = a.out + (size_t)b * H * W;
        const int th = min(TH, H - y0), tw = min(TW, W - x0);
        unsigned char *o0 = o + (size_t)y0 * W + x0;
        if (tw == W && TW == W && ((th * W) & 15) == 0 && ((reinterpret_cast<size_t>(o0) & 15) == 0)) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(obuf);
            uint4 *d4 = reinterpret_cast<uint4 *>(o0);
            for (int i = threadIdx.x; i < (th * W) >> 4; i += DIL_THREADS) __stcs(d4 + i, s4[i]);
        } else {
            for (int i = threadIdx.x; i < th * tw; i += DIL_THREADS) {
                const int r = i / tw, c = i - r * tw;
                o0[(size_t)r * W + c] = obuf[r * TW + c];
            }
        }
        // no barrier here: the next tile's staging writes src and the halo rows of tmp only (both last read before the
        // barrier above), and obuf is not written again before two more barriers
    }
}

// dilate_strip_tma_kernel<KH, KW> — whole-grid tiles of the strip kernel with the grid fetched by the TMA engine: ONE thread
// issues a bulk copy (cp.async.bulk global -> shared, completion counted in bytes on an mbarrier) of the next grid's
// H*W*8 contiguous bytes as soon as the current grid has been converted to int32, so the fetch overlaps the row pass, the
// column pass and the stores of the current grid; no thread spends registers or issue slots on global loads and none
// waits for them at the head of a tile (profiles/r2_dilate_strip_ncu_summary.txt: 56 % of the strip kernel's stall samples).
// Two CTAs per SM (raw grid 51 KB + int32 tile 28 KB + row maxima 29 KB + image 6 KB for 80 x 80).
__device__ __forceinline__ unsigned dil_smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dil_tma_fetch(unsigned bar, unsigned dst, const void *src, unsigned bytes) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the generic-proxy reads of the buffer are done (barrier before)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    for (unsigned off = 0; off < bytes; off += 16384u) {
        const unsigned n = (bytes - off < 16384u) ? bytes - off : 16384u;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + off),
                     "l"(reinterpret_cast<const char *>(src) + off), "r"(n), "r"(bar)
                     : "memory");
    }
}
__device__ __forceinline__ bool dil_mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

template <int KH, int KW>
__global__ void __launch_bounds__(DIL_THREADS, 2) dilate_strip_tma_kernel(const DilateStripArgs a) {
    extern __shared__ __align__(16) unsigned char cm_smem[];
    constexpr int ah = KH / 2, aw = KW / 2;
    const int H = a.H, W = a.W, PS = a.PS, PT = a.PT, HW = H * W;
    const int NSX = (W + DIL_SEG - 1) / DIL_SEG, NSY = (H + DIL_SEG - 1) / DIL_SEG;
    const double2 *raw2 = reinterpret_cast<const double2 *>(cm_smem);      // [H*W] float64, written by the TMA engine
    int *src = reinterpret_cast<int *>(cm_smem + a.raw_bytes);             // [H][PS]
    int *tmp = src + (size_t)H * PS;                                       // [NSY*SEG + KH - 1][PT]
    unsigned char *obuf = cm_smem + a.obuf_off;                            // [H][W]
    const unsigned bar = dil_smem_addr(cm_smem + a.bar_off), raw_s = dil_smem_addr(cm_smem);
    const unsigned bytes = (unsigned)HW * 8u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < H * (KW - 1); i += DIL_THREADS) {
        const int r = i / (KW - 1), j = i - r * (KW - 1);
        src[r * PS + (j < aw ? j : W + j)] = DIL_NEG;
    }
    for (int i = threadIdx.x; i < (KH - 1) * W; i += DIL_THREADS) {
        const int j = i / W, x = i - j * W;
        tmp[(j < ah ? j : H + j) * PT + x] = DIL_NEG;
    }
    __syncthreads();
    if (threadIdx.x == 0 && (int)blockIdx.x < a.B) dil_tma_fetch(bar, raw_s, a.in + (size_t)blockIdx.x * HW, bytes);
    const int half = HW >> 1;
    const int dy = (2 * DIL_THREADS) / W, dx = 2 * DIL_THREADS - dy * W;
    const int ys = (2 * (int)threadIdx.x) / W, xs = 2 * (int)threadIdx.x - ys * W;
    unsigned parity = 0;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        while (!dil_mbar_try_wait(bar, parity)) {}
        parity ^= 1u;
        {
            int y = ys, x = xs;
            for (int i = threadIdx.x; i < half; i += DIL_THREADS) {
                const double2 v = raw2[i];
                int *d = src + y * PS + aw + x;
                d[0] = dil_cvt(v.x); d[1] = dil_cvt(v.y);
                x += dx; y += dy;
                if (x >= W) { x -= W; y++; }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && b + (int)gridDim.x < a.B) dil_tma_fetch(bar, raw_s, a.in + (size_t)(b + gridDim.x) * HW, bytes);
        for (int i = threadIdx.x; i < H * NSX; i += DIL_THREADS) {
            const int s = i / H, rr = i - s * H;
            const int *p = src + rr * PS + s * DIL_SEG;
            int v[DIL_SEG + KW - 1];
#pragma unroll
            for (int j = 0; j < DIL_SEG + KW - 1; j++) v[j] = p[j];
            dil_window_max<KW, DIL_SEG>(v);
            int *q = tmp + (ah + rr) * PT + s * DIL_SEG;
            if ((s + 1) * DIL_SEG <= W) {
#pragma unroll
                for (int j = 0; j < DIL_SEG; j++) q[j] = v[j];
            } else {
#pragma unroll
                for (int j = 0; j < DIL_SEG; j++)
                    if (s * DIL_SEG + j < W) q[j] = v[j];
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < W * NSY; i += DIL_THREADS) {
            const int s = i / W, x = i - s * W;
            const int *p = tmp + (s * DIL_SEG) * PT + x;
            int v[DIL_SEG + KH - 1];
#pragma unroll
            for (int j = 0; j < DIL_SEG + KH - 1; j++) v[j] = p[j * PT];
            dil_window_max<KH, DIL_SEG>(v);
            unsigned char *ob = obuf + (s * DIL_SEG) * W + x;
            if ((s + 1) * DIL_SEG <= H) {
#pragma unroll
                for (int j = 0; j < DIL_SEG; j++) ob[j * W] = (unsigned char)v[j];
            } else {
#pragma unroll
                for (int j = 0; j < DIL_SEG; j++)
                    if (s * DIL_SEG + j < H) ob[j * W] = (unsigned char)v[j];
            }
        }
        __syncthreads();
        unsigned char *o = a.out + (size_t)b * HW;
        if ((HW & 15) == 0 && ((reinterpret_cast<size_t>(o) & 15) == 0)) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(obuf);
            uint4 *d4 = reinterpret_cast<uint4 *>(o);
            for (int i = threadIdx.x; i < HW >> 4; i += DIL_THREADS) __stcs(d4 + i, s4[i]);
        } else {
            for (int i = threadIdx.x; i < HW; i += DIL_THREADS) o[i] = obuf[i];
        }
        // (src is next written after the mbarrier wait, tmp after the barrier behind the conversion, obuf after two more)
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

// inflate_bits_kernel — the same gather for cells_inflation <= 15 with the stamping sources as BIT rows: a window row is one
// funnel shift of two words, and only its set bits are visited.  Sources are sparse (obstacle cells), so a cell costs
// 2c+1 window extractions instead of (2c+1)^2 byte tests.  Staging is one warp per tile row: coalesced loads, the source
// predicate of 32 cells becomes one word by a warp vote.
__global__ void __launch_bounds__(COSTMAP_THREADS) inflate_bits_kernel(const InflateArgs a) {
    extern __shared__ __align__(16) unsigned char cm_smem[];
    const int c = a.c, n = 2 * c + 1;
    const int SH = a.TH + 2 * c, SW = a.TW + 2 * c;
    const int WPR = (SW + 31) / 32 + 1;                           // + one word so that the funnel shift may read w + 1
    double *src = reinterpret_cast<double *>(cm_smem);            // [SH][SW] original values
    double *Ms = src + (size_t)SH * SW;                           // [n][n]
    unsigned *obits = reinterpret_cast<unsigned *>(Ms + (size_t)n * n); // [SH][WPR] bit cc of row r = staged cell (r, cc) is a source
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = COSTMAP_THREADS / 32;
    const int tiles_x = (a.W + a.TW - 1) / a.TW, tiles_y = (a.H + a.TH - 1) / a.TH;
    const int per_grid = tiles_x * tiles_y;
    const unsigned nmask = (n >= 32) ? 0xffffffffu : ((1u << n) - 1u);
    for (int i = threadIdx.x; i < n * n; i += COSTMAP_THREADS) Ms[i] = a.M[i];
    for (int r = threadIdx.x; r < SH; r += COSTMAP_THREADS) obits[r * WPR + WPR - 1] = 0u;
    for (long long t = blockIdx.x; t < (long long)a.B * per_grid; t += gridDim.x) {
        const int b = (int)(t / per_grid), tt = (int)(t % per_grid);
        const int y0 = (tt / tiles_x) * a.TH, x0 = (tt % tiles_x) * a.TW;
        const double *g = a.in + (size_t)b * a.H * a.W;
        for (int r = wid; r < SH; r += nw) {
            const int y = y0 + r - c;
            const bool yin = (y >= 0 && y < a.H), ysrc = (y >= c && y + c < a.H);
            for (int c0 = 0; c0 < SW; c0 += 32) {
                const int cc = c0 + lane, x = x0 + cc - c;
                const bool inside = yin && cc < SW && x >= 0 && x < a.W;
                const double v = inside ? __ldcs(g + (size_t)y * a.W + x) : 1.0;
                if (cc < SW) src[r * SW + cc] = v;
                // a source: original value exactly 0 and the whole window inside the grid (costmap.py:10-14)
                const unsigned word = __ballot_sync(0xffffffffu, inside && ysrc && v == 0.0 && x >= c && x + c < a.W);
                if (lane == 0) obits[r * WPR + (c0 >> 5)] = word;
            }
        }
        __syncthreads();
        double *o = a.out + (size_t)b * a.H * a.W;
        for (int i = threadIdx.x; i < a.TH * a.TW; i += COSTMAP_THREADS) {
            const int r = i / a.TW, cc = i - r * a.TW;
            const int y = y0 + r, x = x0 + cc;
            if (y >= a.H || x >= a.W) continue;
            double m = src[(size_t)(r + c) * SW + cc + c];
            const int w = cc >> 5, sh = cc & 31;
            // window bit j <-> staged column cc + j <-> dx = c - j, inflation-matrix entry [dy + c][2c - j]
            for (int dy = -c; dy <= c; dy++) {
                const unsigned *brow = obits + (r + c - dy) * WPR + w;
                unsigned bits = __funnelshift_r(brow[0], brow[1], sh) & nmask;
                const double *mrow = Ms + (size_t)(dy + c) * n + 2 * c;
                while (bits) {
                    const int j = __ffs(bits) - 1;
                    m = fmin(m, mrow[-j]);
                    bits &= bits - 1;
                }
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

// KH = KW = 0: structuring element from the arguments (generic passes).  KH, KW > 0: compile-time element and a grid side
// that is a multiple of 16 (the publishers' case, 10 x 10 on 80 x 80): the shifts of the horizontal pass unroll, the
// vertical pass takes strips of LCM_VSEG rows per lane (window ORs by doubling in registers, as dil_window_max does for
// maxima) and leaves its words in shared memory, and the byte expansion reads ONE word per 16-byte store; the index
// divisions are multiplications (profiles/r2_lcm_v1_ncu_summary.txt: the generic passes were 68 % of the kernel's
// instructions, the kernel issue-bound at 78 % of the issue slots).
#define LCM_VSEG 8
template <int K, int NOUT>
__device__ __forceinline__ void lcm_window_or(unsigned (&v)[NOUT + K - 1]) {
    constexpr int LEN = NOUT + K - 1;
    constexpr int P = K >= 32 ? 32 : K >= 16 ? 16 : K >= 8 ? 8 : K >= 4 ? 4 : K >= 2 ? 2 : 1;
#pragma unroll
    for (int q = 1; q < P; q <<= 1) {
#pragma unroll
        for (int j = 0; j < LEN; j++)
            if (j + 2 * q <= LEN) v[j] |= v[j + q];
    }
    if (K > P) {
#pragma unroll
        for (int t = 0; t < NOUT; t++) v[t] |= v[t + K - P];
    }
}

template <int KH, int KW>
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
    // compile-time element: table byte of 8 cells -> their 8 image bytes, behind the warps' buffers (two loads per 16-byte store)
    uint2 *lut = reinterpret_cast<uint2 *>(cm_smem + (size_t)LCM_WARPS * per_warp);
    if (KH > 0) {
        for (int e = threadIdx.x; e < 256; e += LCM_WARPS * 32)
            lut[e] = make_uint2((((e & 0xfu) * 0x00204081u) & 0x01010101u) * a.value, ((((e >> 4) & 0xfu) * 0x00204081u) & 0x01010101u) * a.value);
        __syncthreads();
    }
    for (int b = blockIdx.x * LCM_WARPS + wid; b < a.B; b += gridDim.x * LCM_WARPS) {
        for (int w = lane; w < nc * wpr; w += 32) bits[w] = 0u;
        __syncwarp();
        const double *sc = a.scan + (size_t)b * a.n;
        {   // the warp's next scan on its way into L2 while this one is processed (one 128-byte line per lane)
            const long long bn = (long long)b + (long long)gridDim.x * LCM_WARPS;
            if (bn < a.B && lane * 16 < a.n) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.scan + (size_t)bn * a.n + lane * 16));
        }
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
        if (KH > 0) {
            // ---- compile-time element ----
            constexpr int kleft = KW - 1 - KW / 2, kright = KW / 2, kdown = KH / 2; // (rows above: KH - 1 - kdown)
            const unsigned mw = (1u << 20) / (unsigned)wpr + 1u; // i / wpr = (i * mw) >> 20 for i < 2^20 / wpr (i < 8192, wpr <= 16)
            for (int i = lane; i < nc * wpr; i += 32) {
                const int y = (int)(((unsigned)i * mw) >> 20), w = i - y * wpr;
                const unsigned *row = bits + y * wpr;
                const unsigned lo = lcm_word(row, wpr, w - 1), mid = row[w], hi = lcm_word(row, wpr, w + 1);
                unsigned acc = mid;
#pragma unroll
                for (int s = 1; s <= kright; s++) acc |= __funnelshift_l(lo, mid, s);
#pragma unroll
                for (int s = 1; s <= kleft; s++) acc |= __funnelshift_r(mid, hi, s);
                hor[i] = acc;
            }
            __syncwarp();
            // vertical pass: lane <-> (strip of LCM_VSEG rows, word column); ver[y] = OR of hor[ys], ys in [y - kdown, y + KH - 1 - kdown]
            unsigned *ver = bits; // (the occupancy bits are dead)
            const int nstrips = (nc + LCM_VSEG - 1) / LCM_VSEG;
            for (int i = lane; i < nstrips * wpr; i += 32) {
                const int st = (int)(((unsigned)i * mw) >> 20), w = i - st * wpr;
                const int yb = st * LCM_VSEG - kdown;
                unsigned v[LCM_VSEG + KH - 1];
#pragma unroll
                for (int j = 0; j < LCM_VSEG + KH - 1; j++) v[j] = ((unsigned)(yb + j) < (unsigned)nc) ? hor[(yb + j) * wpr + w] : 0u;
                lcm_window_or<KH, LCM_VSEG>(v);
#pragma unroll
                for (int j = 0; j < LCM_VSEG; j++)
                    if (st * LCM_VSEG + j < nc) ver[(st * LCM_VSEG + j) * wpr + w] = v[j];
            }
            __syncwarp();
            // bytes: 16 cells per store, consecutive lanes write consecutive 16-byte pieces of the image
            unsigned char *o = a.out + (size_t)b * nc * nc;
            const int groups = nc >> 4;
            const unsigned mg = (1u << 20) / (unsigned)groups + 1u; // i < nc * groups <= 16384 < 2^20 / groups
            for (int i = lane; i < nc * groups; i += 32) {
                const int y = (int)(((unsigned)i * mg) >> 20), g = i - y * groups;
                const unsigned m16 = (ver[y * wpr + (g >> 1)] >> ((g & 1) << 4)) & 0xffffu;
                const uint2 qa = lut[m16 & 0xffu], qb = lut[m16 >> 8];
                __stcs(reinterpret_cast<uint4 *>(o) + i, make_uint4(qa.x, qa.y, qb.x, qb.y));
            }
        } else {
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
        }
        __syncwarp();
    }
}
