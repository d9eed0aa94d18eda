"""Obstacle-list construction: laser scan -> local occupancy grid -> fixed-length obstacle point list.

Product path (bottom of this file): `get_obstacles_gpu` / `get_obstacles_batch_gpu` -> libb200mpc.so
(b200mpc_obstacles_batch, csrc/obstacles_kernel.cuh).  It needs a CUDA device and fails loudly without one; it never
calls the numpy functions below.

The numpy functions (top of this file) are a batched, line-by-line mirror of the reference's producer
  get_obstacles                          scripts/point_follower_local_planner.py:88-118
  convert_laser_scan_to_occupancy_grid   utils/utils.py:5-43
  convert_to_map_coordinates             utils/utils.py:114-124
  rotate_coordinates                     utils/utils.py:145-152
used for two things only: building the synthetic workloads (synth.py: ray-cast scans -> obstacle lists, on machines
without a GPU as well) and as the second checker of the kernel in the tests.  They are bit-exact in the cell indices
against the reference's own numba helpers (tests/golden/obstacles_golden.npz).  Quirks kept on purpose:
  * beam angle i*(max-min)/n + min (the last beam stops one step short of angle_max);
  * the rotation-by-0.0 matrix product turns +-inf coordinates into NaN (0*inf), NaN becomes 0, so an
    infinite-range beam marks the robot's own cell (40,40);
  * int() truncates toward zero, so coordinates in (-0.05, 0) land in cell 0;
  * cells are emitted in np.where (row-major) order of the 180-degree-rotated grid, padded with the first
    obstacle, or all 100.0 when the scan hits nothing.
Deviation: more than `slots` cells make the reference raise ValueError; here
`overflow="truncate"` keeps the first `slots` cells (default for synthetic workloads), "raise" reproduces it.
"""
import numpy as np


def scan_to_cells(scan, angles, resolution, map_size, rotation=0.0):
    """utils.py:5-43 up to the cell indices. scan (B,n) float64; angles (2,) or (B,2).

    Returns (ix, iy, valid): integer cell indices (B,n) and the in-grid mask."""
    scan = np.atleast_2d(np.asarray(scan, dtype=np.float64))
    B, n = scan.shape
    angles = np.broadcast_to(np.asarray(angles, dtype=np.float64), (B, 2))
    amin, amax = angles[:, :1], angles[:, 1:2]
    num_cells = int(map_size / resolution)
    beam = np.arange(n)[None, :] * (amax - amin) / n + amin
    with np.errstate(invalid="ignore"):
        x = scan * np.cos(beam)
        y = scan * np.sin(beam)
        # rotate_coordinates(coordinates, rotation): [[c,-s],[s,c]] @ [x;y]
        c, s = np.cos(rotation), np.sin(rotation)
        xr = c * x + (-s) * y
        yr = s * x + c * y
    xr = np.where(np.isnan(xr), 0.0, xr)
    yr = np.where(np.isnan(yr), 0.0, yr)
    for arr in (xr, yr):
        inf = np.isinf(arr)
        if inf.any():
            # np.max(x[~isinf(x)]) per scan
            fin = np.where(inf, -np.inf, arr).max(axis=1, keepdims=True)
            np.copyto(arr, np.broadcast_to(fin, arr.shape), where=inf)
    xi = xr + (map_size / 2)
    yi = yr + (map_size / 2)
    ix = np.trunc(xi / resolution).astype(np.int64)
    iy = np.trunc(yi / resolution).astype(np.int64)
    valid = (ix >= 0) & (ix < num_cells) & (iy >= 0) & (iy < num_cells)
    return ix, iy, valid


def scan_to_occupancy_grid(scan, angles, resolution, map_size, rotation=0.0):
    """utils.py:5-43: (B,num_cells,num_cells) float64 grid with 100 at hit cells (grid[y, x])."""
    ix, iy, valid = scan_to_cells(scan, angles, resolution, map_size, rotation)
    B = ix.shape[0]
    num_cells = int(map_size / resolution)
    grid = np.zeros((B, num_cells, num_cells))
    b = np.broadcast_to(np.arange(B)[:, None], ix.shape)
    grid[b[valid], iy[valid], ix[valid]] = 100.0
    return grid


def get_obstacles(scan, angles, size, resolution, pos, yaw, slots, overflow="truncate"):
    """Batched get_obstacles (point_follower_local_planner.py:88-118).

    scan (B,n); angles (2,)|(B,2); pos (B,2); yaw (B,) (= ori[2]); slots = len(obstacles_x) = 160.
    Returns obstacles_x, obstacles_y (B,slots) float64 and the raw cell count per robot (B,)."""
    scan = np.atleast_2d(np.asarray(scan, dtype=np.float64))
    B = scan.shape[0]
    pos = np.broadcast_to(np.asarray(pos, dtype=np.float64), (B, 2))
    yaw = np.broadcast_to(np.asarray(yaw, dtype=np.float64), (B,))
    grid = scan_to_occupancy_grid(scan, angles, resolution, size * 2)
    occ = 1 - grid / 100
    occ = occ[:, ::-1, ::-1]  # np.rot90(k=2)
    nc = occ.shape[1]
    origin = (nc // 2) * resolution
    obs_x = np.empty((B, slots))
    obs_y = np.empty((B, slots))
    count = np.zeros(B, dtype=np.int64)
    for b in range(B):
        ii, jj = np.where(occ[b] == 0)
        count[b] = len(ii)
        if len(ii) == 0:  # IndexError branch: "No obstacles"
            obs_x[b, :] = 100.0
            obs_y[b, :] = 100.0
            continue
        if len(ii) > slots:
            if overflow == "raise":
                raise ValueError(f"could not broadcast input array from shape ({len(ii)},) into shape ({slots},)")
            ii, jj = ii[:slots], jj[:slots]
        # convert_to_map_coordinates: x[i,j] = -i*res + origin, y[i,j] = -j*res + origin
        ox = -ii * resolution + origin
        oy = -jj * resolution + origin
        c, s = np.cos(yaw[b]), np.sin(yaw[b])
        wx = c * ox + (-s) * oy + pos[b, 0]
        wy = s * ox + c * oy + pos[b, 1]
        obs_x[b, :] = wx[0]
        obs_y[b, :] = wy[0]
        obs_x[b, :len(wx)] = wx
        obs_y[b, :len(wy)] = wy
    return obs_x, obs_y, count


# ---- CUDA path (libb200mpc.so: b200mpc_obstacles_batch) --------------------------------------------------------------
def beam_table(n_beams, angles):
    """cos / sin of the beam angles exactly as utils.py:18-20 forms them (a property of the lidar, computed once on
    the host so that the device's cell indices are bit-exact with the reference)."""
    amin, amax = float(angles[0]), float(angles[1])
    beam = np.arange(n_beams) * (amax - amin) / n_beams + amin
    return np.cos(beam), np.sin(beam)


_SOLVERS = {}


def _default_solver(device=0):
    """The obstacle builder does not depend on the MPC variant; any handle on the device will do."""
    if device not in _SOLVERS:
        from . import _shim, load_params, make_params  # noqa: PLC0415
        _SOLVERS[device] = _shim.Solver(make_params("B", load_params()), device=device)
    return _SOLVERS[device]


def get_obstacles_batch_gpu(scan, angles, size, resolution, pos, yaw, slots, overflow="truncate", solver=None):
    """Batched get_obstacles on the GPU; same arguments and results as `get_obstacles` above (the numpy mirror)."""
    scan = np.atleast_2d(np.asarray(scan, dtype=np.float64))
    B, n = scan.shape
    angles = np.asarray(angles, dtype=np.float64)
    if angles.ndim != 1:
        if not np.all(angles == angles[0]):
            raise ValueError("the GPU path takes one (angle_min, angle_max) pair for the whole batch (one lidar model)")
        angles = angles[0]
    bc, bs = beam_table(n, angles)
    pos = np.ascontiguousarray(np.broadcast_to(np.asarray(pos, dtype=np.float64), (B, 2)))
    yaw = np.ascontiguousarray(np.broadcast_to(np.asarray(yaw, dtype=np.float64), (B,)))
    S = solver or _default_solver()
    ox, oy, cnt = S.obstacles_batch(scan, bc, bs, pos, yaw, size, resolution, slots)
    if overflow == "raise" and (cnt > slots).any():
        k = int(cnt[np.argmax(cnt > slots)])
        raise ValueError(f"could not broadcast input array from shape ({k},) into shape ({slots},)")
    return ox, oy, cnt.astype(np.int64)


def get_obstacles_gpu(scan_data, angles, size, resolution, pos, ori, obstacles_x, obstacles_y, solver=None):
    """Drop-in for get_obstacles(scan_data, angles, size, resolution, pos, ori, obstacles_x, obstacles_y)
    (scripts/point_follower_local_planner.py:88-118): one robot, ori = (roll, pitch, yaw), obstacles_x / obstacles_y
    only give the number of slots.  Raises ValueError when the scan marks more cells than there are slots, as the
    reference does; returns the 100.0 sentinel arrays when it marks none."""
    slots = int(np.size(obstacles_x))
    ox, oy, _ = get_obstacles_batch_gpu(np.asarray(scan_data)[None], angles, size, resolution, np.asarray(pos)[None, :2],
                                        np.asarray([ori[2]]), slots, overflow="raise", solver=solver)
    return ox[0].reshape(np.shape(obstacles_x)), oy[0].reshape(np.shape(obstacles_y))
