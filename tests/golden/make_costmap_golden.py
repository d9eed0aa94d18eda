"""Golden vectors for the costmap operations, generated in the build container from the
reference's own code and from OpenCV (the GPU box has neither /root/reference nor a guarantee of the same cv2 build):

  utils/costmap.py:5-59            get_inflation_matrix, inflate_global, inflate_local — the unmodified numba functions
  utils/utils.py:5-43              convert_laser_scan_to_occupancy_grid with rotation = yaw (what the local costmap publisher
                                   calls, core/local_costmap_publisher.py:29-31)
  cv2.dilate(grid, np.ones((kh, kw)), iterations=1).astype(np.uint8)      core/local_costmap_publisher.py:34-35 (cv2 4.13)

Output: tests/golden/costmap_golden.npz (bit-packed / uint8 where the data allows, to stay small)."""
import importlib.util
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/ros2_mpc/utils"


def load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    cm, ut = load("costmap"), load("utils")
    rng = np.random.default_rng(2026)
    out = {"cv2_version": np.array(cv2.__version__)}
    # ---- inflation matrices ----
    for c, f in ((2, 1.3), (4, 1.3), (5, 1.3), (3, 2.0)):
        out[f"matrix_c{c}_f{f}"] = cm.get_inflation_matrix(c, f)
    # ---- inflate_global / inflate_local: grids with values {0 (source), 50, 100}; sources also in the border band ----
    cases = []
    for i, (H, W, c, dens) in enumerate(((80, 80, 4, 0.01), (37, 53, 5, 0.03), (20, 20, 2, 0.1), (64, 200, 3, 0.005), (9, 9, 4, 0.2))):
        g = np.full((H, W), 100.0)
        g[rng.random((H, W)) < 0.1] = 50.0
        g[rng.random((H, W)) < dens] = 0.0
        g[0, 0] = 0.0; g[H - 1, W - 1] = 0.0; g[H // 2, 0] = 0.0  # clipped windows: skipped by the reference
        M = cm.get_inflation_matrix(c)
        out[f"infl{i}_grid"] = g.astype(np.uint8)
        out[f"infl{i}_c"] = np.array(c)
        out[f"infl{i}_out"] = cm.inflate_global(g.copy(), M, c)
        cases.append((g, M, c))
    out["n_infl"] = np.array(len(cases))
    # inflate_local: crops incl. bounds beyond the grid and negative (wrapping) starts, as Python slicing does
    g, M, c = cases[0]
    locs = [((40.0, 40.0), 30), ((10.0, 70.0), 30), ((5.0, 5.0), 30), ((79.0, 40.0), 20), ((40.5, 12.5), 25)]
    for i, (pos, size) in enumerate(locs):
        out[f"local{i}_pos"] = np.array(pos)
        out[f"local{i}_size"] = np.array(size)
        out[f"local{i}_out"] = cm.inflate_local(g.copy(), M, c, np.array(pos), size)
    out["n_local"] = np.array(len(locs))
    # ---- dilation: occupancy grids 0 / 100 (scan-like), a float grid with arbitrary values, several kernel sizes ----
    dil = []
    g1 = np.zeros((80, 80)); g1[rng.integers(0, 80, 60), rng.integers(0, 80, 60)] = 100.0
    g2 = np.zeros((224, 314)); g2[rng.random((224, 314)) < 0.02] = 100.0
    g3 = np.round(rng.uniform(0, 255, (33, 47)), 1)
    g4 = np.zeros((80, 80)); g4[0, :] = 100.0; g4[:, 79] = 100.0; g4[79, 0] = 100.0
    for i, (g, ks) in enumerate(((g1, (10, 10)), (g2, (10, 10)), (g3, (3, 5)), (g4, (10, 10)), (g1, (8, 8)), (g3, (1, 1)), (g1, (7, 4)))):
        out[f"dil{i}_grid"] = g
        out[f"dil{i}_k"] = np.array(ks)
        out[f"dil{i}_out"] = cv2.dilate(g, np.ones(ks), iterations=1).astype(np.uint8)
        dil.append(i)
    out["n_dil"] = np.array(len(dil))
    # ---- local costmap publisher: scan -> grid rotated by yaw -> dilate -> uint8 ----
    n = 360
    B = 48
    scans = np.round(rng.uniform(0.12, 3.5, (B, n)), 2)
    scans[rng.random((B, n)) < 0.04] = np.inf
    scans[rng.random((B, n)) < 0.01] = np.nan
    scans[3] = np.inf                      # nothing in range
    scans[4, ::2] = 0.0                    # hits on the robot's own cell
    yaw = rng.uniform(-np.pi, np.pi, B)
    yaw[0] = 0.0; yaw[1] = np.pi / 2; yaw[2] = -np.pi
    angles = np.array([0.0, 6.28])
    imgs = np.empty((B, 80, 80), np.uint8)
    grids = np.empty((B, 80, 80), np.uint8)
    for b in range(B):
        g = ut.convert_laser_scan_to_occupancy_grid(scans[b].copy(), angles, 0.05, 4.0, yaw[b])
        grids[b] = g.astype(np.uint8)
        imgs[b] = cv2.dilate(g, np.ones((10, 10)), iterations=1).astype(np.uint8)
    out["lcm_scan"] = scans; out["lcm_yaw"] = yaw; out["lcm_angles"] = angles
    out["lcm_grid_bits"] = np.packbits(grids > 0)
    out["lcm_img_bits"] = np.packbits(imgs > 0)   # the images hold 0 / 100 only
    assert set(np.unique(imgs)) <= {0, 100}
    np.savez_compressed(os.path.join(HERE, "costmap_golden.npz"), **out)
    print("wrote costmap_golden.npz", os.path.getsize(os.path.join(HERE, "costmap_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
