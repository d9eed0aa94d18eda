"""Seeded synthetic workloads for the configurations named in BASELINE.json (generators: SURVEY.md section 8d).

Nothing here is on the measured path: these functions only build the inputs (x0, goal / reference, obstacle
lists, initial controls) that the solve consumes.  The map is the reference's maps/map_carto.pgm, shipped as the
derived data file ros2_mpc_b200/data/map_carto_occ.npz (tests/golden/make_map_fixture.py writes it).
"""
import os

import numpy as np

from . import obstacles as _obs

MAP_FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "map_carto_occ.npz")


def load_map(path=None):
    z = np.load(path or MAP_FIXTURE)
    shape = tuple(int(v) for v in z["shape"])
    n = shape[0] * shape[1]
    occ = np.unpackbits(z["occ_bits"])[:n].reshape(shape).astype(bool)
    free = np.unpackbits(z["free_bits"])[:n].reshape(shape).astype(bool)
    return dict(occ=occ, free=free, resolution=float(z["resolution"]), origin=np.array(z["origin"], dtype=np.float64))


def clearance(m):
    """Distance [m] from every cell centre to the nearest occupied cell."""
    from scipy.ndimage import distance_transform_edt  # noqa: PLC0415
    return distance_transform_edt(~m["occ"]) * m["resolution"]


def world_to_cell(m, xy):
    ij = np.floor((np.asarray(xy) - m["origin"]) / m["resolution"]).astype(np.int64)
    return ij[..., 1], ij[..., 0]  # row (y), col (x)


def raycast(m, pos, yaw, n_beams=360, angle_min=0.0, angle_max=6.28, range_min=0.12, range_max=3.5, step=0.01):
    """Laser scan of the static map from poses (pos (B,2), yaw (B,)).  Beam i points along
    yaw + i*(angle_max-angle_min)/n + angle_min (the convention of utils.py:18).  A beam that hits nothing
    within range_max reports range_max (outside the 4 m local grid, so it marks no cell)."""
    pos = np.asarray(pos, dtype=np.float64)
    yaw = np.asarray(yaw, dtype=np.float64)
    B = pos.shape[0]
    H, W = m["occ"].shape
    ang = yaw[:, None] + (np.arange(n_beams)[None, :] * (angle_max - angle_min) / n_beams + angle_min)
    ca, sa = np.cos(ang), np.sin(ang)
    rng_out = np.full((B, n_beams), range_max)
    alive = np.ones((B, n_beams), dtype=bool)
    occ = m["occ"]
    ox, oy, res = m["origin"][0], m["origin"][1], m["resolution"]
    nsteps = int(round((range_max - range_min) / step)) + 1
    for t in range(nsteps):
        if not alive.any():
            break
        r = range_min + t * step
        col = np.floor((pos[:, 0:1] + r * ca - ox) / res).astype(np.int64)
        row = np.floor((pos[:, 1:2] + r * sa - oy) / res).astype(np.int64)
        inside = (row >= 0) & (row < H) & (col >= 0) & (col < W)
        hit = np.zeros_like(alive)
        hit[inside] = occ[row[inside], col[inside]]
        newhit = alive & hit
        rng_out[newhit] = r
        alive &= ~hit
    return rng_out, np.array([angle_min, angle_max])


def _sample_free_poses(m, clr, rng, B, min_clearance):
    rows, cols = np.where(m["free"] & (clr >= min_clearance))
    pick = rng.integers(0, len(rows), size=B)
    jitter = rng.uniform(0.0, 1.0, size=(B, 2))
    x = m["origin"][0] + (cols[pick] + jitter[:, 0]) * m["resolution"]
    y = m["origin"][1] + (rows[pick] + jitter[:, 1]) * m["resolution"]
    yaw = rng.uniform(0.0, 2 * np.pi, size=B)
    return np.stack([x, y, yaw], axis=1)


def robots_on_map(B=4096, seed=0, params=None, slots=160, min_clearance=0.3, goal_range=(0.3, 1.0), m=None):
    """Config 3: B random initial states / goals on map_carto with scan-derived obstacle lists."""
    m = m or load_map()
    clr = clearance(m)
    rng = np.random.Generator(np.random.PCG64(seed))
    x0 = _sample_free_poses(m, clr, rng, B, min_clearance)
    goal = np.zeros((B, 3))
    todo = np.arange(B)
    H, W = m["occ"].shape
    for _ in range(200):
        if len(todo) == 0:
            break
        rho = rng.uniform(goal_range[0], goal_range[1], size=len(todo))
        phi = rng.uniform(0.0, 2 * np.pi, size=len(todo))
        g = x0[todo, :2] + np.stack([rho * np.cos(phi), rho * np.sin(phi)], axis=1)
        r, c = world_to_cell(m, g)
        ok = (r >= 0) & (r < H) & (c >= 0) & (c < W)
        ok[ok] &= m["free"][r[ok], c[ok]]
        goal[todo[ok], :2] = g[ok]
        todo = todo[~ok]
    if len(todo):
        goal[todo, :2] = x0[todo, :2]
    goal[:, 2] = rng.uniform(0.0, 2 * np.pi, size=B)
    size = 2.0 if params is None else params["costmap_size"]
    res = 0.05 if params is None else params["resolution"]
    scan, angles = raycast(m, x0[:, :2], x0[:, 2])
    obs_x, obs_y, count = _obs.get_obstacles(scan, angles, size, res, x0[:, :2], x0[:, 2], slots)
    return dict(x0=x0, goal=goal, obs_x=obs_x, obs_y=obs_y, obs_count=count, scan=scan, angles=angles)


def warm_start_seeds(n_seeds, N, u_lo, u_hi, first_seed=1):
    """Config 4: u_init = clip(N(0, 0.05^2), bounds) from PCG64(seed = first_seed + s); (n_seeds, N, 2)."""
    out = np.empty((n_seeds, N, 2))
    lo, hi = np.asarray(u_lo), np.asarray(u_hi)
    for s in range(n_seeds):
        rng = np.random.Generator(np.random.PCG64(first_seed + s))
        out[s] = np.clip(rng.normal(0.0, 0.05, size=(N, 2)), lo, hi)
    return out


def dense_obstacle_field(x0, slots=160, seed=2, lattice=0.05, r_in=0.3, r_out=1.5):
    """Config 5: `slots` distinct obstacle points per robot on a `lattice` grid in the annulus r_in..r_out."""
    rng = np.random.Generator(np.random.PCG64(seed))
    k = int(np.ceil(r_out / lattice))
    gi, gj = np.meshgrid(np.arange(-k, k + 1), np.arange(-k, k + 1), indexing="ij")
    d = np.hypot(gi, gj) * lattice
    cand = np.stack([gi[(d >= r_in) & (d <= r_out)], gj[(d >= r_in) & (d <= r_out)]], axis=1) * lattice
    B = x0.shape[0]
    obs_x = np.empty((B, slots))
    obs_y = np.empty((B, slots))
    for b in range(B):
        sel = rng.choice(len(cand), size=slots, replace=False)
        obs_x[b] = x0[b, 0] + cand[sel, 0]
        obs_y[b] = x0[b, 1] + cand[sel, 1]
    return obs_x, obs_y


def straight_reference(x0, goal, N):
    """Tracking-variant reference (pxf (B,3N), puf (B,2N)): N points on the segment x0 -> goal, heading along it."""
    B = x0.shape[0]
    t = (np.arange(1, N + 1) / N)[None, :, None]
    xy = x0[:, None, :2] * (1 - t) + goal[:, None, :2] * t
    head = np.arctan2(goal[:, 1] - x0[:, 1], goal[:, 0] - x0[:, 0])
    pxf = np.concatenate([xy, np.broadcast_to(head[:, None, None], (B, N, 1))], axis=2).reshape(B, 3 * N)
    puf = np.tile(np.array([0.1, 0.0]), (B, N))
    return pxf, puf
