"""Costmap inflation and dilation on the GPU, with the reference's function names.

  get_inflation_matrix   ros2_mpc/utils/costmap.py:44-59   (a (2c+1)^2 table: built once on the host)
  inflate_global         ros2_mpc/utils/costmap.py:5-20    -> b200mpc_inflate_batch
  inflate_local          ros2_mpc/utils/costmap.py:23-41   -> the reference's crop (Python slice semantics), then the same kernel
  dilate                 cv2.dilate(grid, np.ones((10, 10)), iterations=1).astype(np.uint8)
                         ros2_mpc/core/local_costmap_publisher.py:34-35, global_costmap_publisher.py -> b200mpc_dilate_batch
  local_costmap          the loop body of the local costmap publisher, scan -> grid (rotation = yaw) -> dilate -> uint8
                         ros2_mpc/core/local_costmap_publisher.py:29-35 -> b200mpc_local_costmap_batch (one fused kernel)

Every function takes one grid (H,W) like the reference or a batch (B,H,W); results are bit-exact with the reference's numba
functions and with OpenCV 4.13 (tests/golden/costmap_golden.npz).  The kernels need a CUDA device; there is no CPU path."""
import numpy as np

from .obstacles import _default_solver, beam_table


def get_inflation_matrix(cells_inflation, factor=1.3):
    """Centre 100, then concentric square rings whose value grows from the rim inwards by (100 / cells_inflation) / factor
    per ring (costmap.py:44-59).  Ring k (k = 0 at the rim) is the set of cells at Chebyshev distance c - k from the centre."""
    c = int(cells_inflation)
    n = 2 * c + 1
    decay = (1 / c) / factor
    i = np.arange(n)
    ring = np.minimum(np.minimum(i[:, None], i[None, :]), np.minimum(n - 1 - i[:, None], n - 1 - i[None, :]))  # 0 at the rim
    m = decay * (ring + 1) * 100
    m[c, c] = 100
    return m


def _as_batch(grid):
    g = np.asarray(grid, dtype=np.float64)
    return (g[None], True) if g.ndim == 2 else (g, False)


def inflate_global(occupancy_grid, inflation_matrix, cells_inflation, solver=None):
    g, single = _as_batch(occupancy_grid)
    out = (solver or _default_solver()).inflate_batch(g, inflation_matrix, cells_inflation)
    return out[0] if single else out


def inflate_local(occupancy_grid, inflation_matrix, cells_inflation, robot_position, costmap_size, solver=None):
    """The reference crops with int() bounds and Python slicing (negative bounds wrap, bounds beyond the grid clip) before
    it inflates; the crop is host index arithmetic, the inflation runs on the device."""
    g = np.asarray(occupancy_grid, dtype=np.float64)
    rows = slice(int(robot_position[1] - costmap_size / 2), int(robot_position[1] + costmap_size / 2))
    cols = slice(int(robot_position[0] - costmap_size / 2), int(robot_position[0] + costmap_size / 2))
    crop = np.ascontiguousarray(g[..., rows, cols])
    if crop.shape[-1] == 0 or crop.shape[-2] == 0:
        return crop.copy()
    return inflate_global(crop, inflation_matrix, cells_inflation, solver)


def dilate(grid, ksize=(10, 10), solver=None):
    g, single = _as_batch(grid)
    out = (solver or _default_solver()).dilate_batch(g, ksize[0], ksize[1])
    return out[0] if single else out


def local_costmap(scan, angles, resolution, costmap_size, yaw, ksize=(10, 10), solver=None):
    """uint8 image(s) the local costmap publisher sends: scan (n,) | (B,n), yaw scalar | (B,)."""
    scan = np.asarray(scan, dtype=np.float64)
    single = scan.ndim == 1
    scan = np.atleast_2d(scan)
    bc, bs = beam_table(scan.shape[1], angles)
    yaw = np.broadcast_to(np.asarray(yaw, dtype=np.float64), (scan.shape[0],))
    out = (solver or _default_solver()).local_costmap_batch(scan, bc, bs, yaw, costmap_size, resolution, ksize[0], ksize[1])
    return out[0] if single else out
