"""Costmap inflation / dilation and the simulated lidar against goldens generated from the
reference's own numba functions and OpenCV 4.13 (tests/golden/make_costmap_golden.py).

CPU part: the host-side pieces (inflation matrix, the gather formulation the kernels implement, written out in numpy) against
the goldens.  GPU part (-m gpu): the kernels through the C ABI, bit-exact."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(G, "costmap_golden.npz"))


# ---- numpy statements of what the kernels compute (checkers; the product never calls them) ---------------------------
def np_inflate(g, M, c):
    """out[p] = min(g[p], min over sources q (g[q] == 0, full window inside) with |p - q|_inf <= c of M[p - q + c])."""
    H, W = g.shape
    out = g.copy()
    src = (g == 0)
    src[:c, :] = False; src[H - c:, :] = False; src[:, :c] = False; src[:, W - c:] = False
    if c == 0:
        src = (g == 0)
    for qi, qj in zip(*np.where(src)):
        win = out[qi - c:qi + c + 1, qj - c:qj + c + 1]
        np.minimum(win, M, out=win)
    return out


def np_dilate(g, kh, kw):
    H, W = g.shape
    ah, aw = kh // 2, kw // 2
    pad = np.full((H + kh - 1, W + kw - 1), -np.inf)
    pad[ah:ah + H, aw:aw + W] = g
    m = np.full((H, W), -np.inf)
    for i in range(kh):
        for j in range(kw):
            m = np.maximum(m, pad[i:i + H, j:j + W])
    return m.astype(np.uint8)


def test_inflation_matrix_matches_reference(gold):
    from ros2_mpc_b200 import costmap as cm
    for c, f in ((2, 1.3), (4, 1.3), (5, 1.3), (3, 2.0)):
        assert np.array_equal(cm.get_inflation_matrix(c, f), gold[f"matrix_c{c}_f{f}"])


def test_gather_formulation_equals_reference_stamping(gold):
    """The reference stamps windows in scan order; the kernel gathers.  Same result (min is order independent)."""
    from ros2_mpc_b200 import costmap as cm
    for i in range(int(gold["n_infl"])):
        g, c = gold[f"infl{i}_grid"].astype(np.float64), int(gold[f"infl{i}_c"])
        assert np.array_equal(np_inflate(g, cm.get_inflation_matrix(c), c), gold[f"infl{i}_out"]), i


def test_dilate_conventions_match_opencv(gold):
    for i in range(int(gold["n_dil"])):
        kh, kw = (int(v) for v in gold[f"dil{i}_k"])
        assert np.array_equal(np_dilate(gold[f"dil{i}_grid"], kh, kw), gold[f"dil{i}_out"]), i


def test_rotated_scan_grid_matches_reference(gold):
    """The numpy mirror of convert_laser_scan_to_occupancy_grid with rotation = yaw (the local costmap publisher's call)."""
    from ros2_mpc_b200 import obstacles as ob
    scans, yaw = gold["lcm_scan"], gold["lcm_yaw"]
    ref = np.unpackbits(gold["lcm_grid_bits"])[:scans.shape[0] * 6400].reshape(-1, 80, 80).astype(bool)
    mism = 0
    for b in range(scans.shape[0]):
        g = ob.scan_to_occupancy_grid(scans[b:b + 1], gold["lcm_angles"], 0.05, 4.0, rotation=yaw[b])[0] > 0
        mism += int((g != ref[b]).sum())
    # numpy rounds c*x + (-s)*y twice, the reference's np.dot fuses one product: a cell may flip when a coordinate sits
    # within an ulp of a cell edge (the kernel uses the fused form and must match exactly, see the gpu test)
    assert mism <= 2


def test_costmap_entry_points_fail_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from ros2_mpc_b200 import costmap as cm
    with pytest.raises(RuntimeError):
        cm.dilate(np.zeros((8, 8)))


# ---- GPU ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_inflate_matches_reference(gold, built):
    from ros2_mpc_b200 import costmap as cm
    for i in range(int(gold["n_infl"])):
        g, c = gold[f"infl{i}_grid"].astype(np.float64), int(gold[f"infl{i}_c"])
        M = cm.get_inflation_matrix(c)
        assert np.array_equal(cm.inflate_global(g, M, c), gold[f"infl{i}_out"]), i
    # a batch of grids in one launch, and inflate_local's crops (bounds beyond the grid, as Python slicing clips them)
    g, c = gold["infl0_grid"].astype(np.float64), int(gold["infl0_c"])
    M = cm.get_inflation_matrix(c)
    batch = np.stack([g, g[::-1].copy(), g.T.copy()])
    out = cm.inflate_global(batch, M, c)
    for k in range(3):
        assert np.array_equal(out[k], np_inflate(batch[k], M, c))
    # an inflation matrix without any symmetry (an index flip would hide behind the reference's ring matrix), on the bit-row
    # kernel (cells_inflation <= 15: 3, 15) and on the generic one (16); grids wider than one tile and narrower than a window
    rng = np.random.default_rng(11)
    for (H, W, cc) in ((80, 80, 3), (70, 150, 15), (90, 140, 16), (9, 40, 4), (33, 33, 0)):
        gg = np.full((2, H, W), 100.0)
        gg[rng.random((2, H, W)) < 0.2] = 37.5
        gg[rng.random((2, H, W)) < 0.02] = 0.0
        Mr = np.round(rng.uniform(1, 99, (2 * cc + 1, 2 * cc + 1)), 3)
        oo = cm.inflate_global(gg, Mr, cc)
        for k in range(2):
            assert np.array_equal(oo[k], np_inflate(gg[k], Mr, cc)), (H, W, cc)
    for i in range(int(gold["n_local"])):
        o = cm.inflate_local(g, M, c, gold[f"local{i}_pos"], int(gold[f"local{i}_size"]))
        assert o.shape == gold[f"local{i}_out"].shape and np.array_equal(o, gold[f"local{i}_out"]), i


@pytest.mark.gpu
def test_gpu_dilate_matches_opencv(gold, built):
    from ros2_mpc_b200 import costmap as cm
    for i in range(int(gold["n_dil"])):
        kh, kw = (int(v) for v in gold[f"dil{i}_k"])
        assert np.array_equal(cm.dilate(gold[f"dil{i}_grid"], (kh, kw)), gold[f"dil{i}_out"]), i
    rng = np.random.default_rng(5)
    grids = (rng.random((257, 80, 80)) < 0.01) * 100.0
    out = cm.dilate(grids)
    for b in (0, 100, 256):
        assert np.array_equal(out[b], np_dilate(grids[b], 10, 10))
    # the 10 x 10 strip kernel on other shapes: whole grid with an odd width, smaller than the structuring element, tiled
    # (several tiles per grid, ragged last tiles), arbitrary non-negative values (truncated toward zero by the cast)
    for shape in ((33, 47), (7, 5), (1, 1), (81, 80), (80, 82), (100, 200), (224, 314), (161, 19)):
        g = np.round(rng.uniform(0, 255.9, (3,) + shape), 2) * (rng.random((3,) + shape) < 0.05)
        o = cm.dilate(g)
        for b in range(3):
            assert np.array_equal(o[b], np_dilate(g[b], 10, 10)), shape


@pytest.mark.gpu
def test_gpu_local_costmap_matches_reference(gold, built):
    """scan -> grid rotated by yaw -> dilate -> uint8 in one kernel == the reference's numba function + cv2.dilate."""
    from ros2_mpc_b200 import costmap as cm
    scans, yaw = gold["lcm_scan"], gold["lcm_yaw"]
    B = scans.shape[0]
    ref = np.unpackbits(gold["lcm_img_bits"])[:B * 6400].reshape(B, 80, 80) * np.uint8(100)
    img = cm.local_costmap(scans, gold["lcm_angles"], 0.05, 2.0, yaw)
    assert img.dtype == np.uint8 and img.shape == (B, 80, 80)
    assert np.array_equal(img, ref), np.argwhere((img != ref).any(axis=(1, 2))).ravel()
    one = cm.local_costmap(scans[7], gold["lcm_angles"], 0.05, 2.0, yaw[7])
    assert np.array_equal(one, ref[7])
    # the generic dilation of the float64 grids gives the same images
    from ros2_mpc_b200 import obstacles as ob
    grid = np.unpackbits(gold["lcm_grid_bits"])[:B * 6400].reshape(B, 80, 80) * 100.0
    assert np.array_equal(cm.dilate(grid), ref)
    # another structuring element takes the kernel's generic passes: same image as the generic dilation of the grid
    assert np.array_equal(cm.local_costmap(scans, gold["lcm_angles"], 0.05, 2.0, yaw, ksize=(7, 4)), cm.dilate(grid, (7, 4)))


@pytest.mark.gpu
def test_gpu_raycast_matches_host_sensor_model(built):
    """The device lidar == synth.raycast (the host model that builds the workloads), sample for sample."""
    from ros2_mpc_b200 import sensors, synth
    m = synth.load_map()
    w = synth.robots_on_map(B=192, seed=3)
    scan_g, ang = sensors.raycast_gpu(m, w["x0"])
    assert np.array_equal(ang, w["angles"])
    assert np.array_equal(scan_g, w["scan"]), np.abs(scan_g - w["scan"]).max()
    # poses outside the map and on occupied cells
    pose = np.array([[-50.0, 3.0, 0.3], [m["origin"][0] + 0.01, m["origin"][1] + 0.01, 1.0], [0.0, 0.0, -2.0]])
    s_h, _ = synth.raycast(m, pose[:, :2], pose[:, 2])
    s_g, _ = sensors.raycast_gpu(m, pose)
    assert np.array_equal(s_g, s_h)
    # other lidar models: fewer beams, another field of view / range / step
    s_h, _ = synth.raycast(m, w["x0"][:16, :2], w["x0"][:16, 2], n_beams=90, angle_min=-1.5, angle_max=1.5, range_max=2.0, step=0.02)
    s_g, _ = sensors.raycast_gpu(m, w["x0"][:16], n_beams=90, angle_min=-1.5, angle_max=1.5, range_max=2.0, step=0.02)
    assert np.array_equal(s_g, s_h)


@pytest.mark.gpu
def test_gpu_get_headings_matches_reference(built):
    """get_headings on the device against outputs of the reference's own function (refgen_golden.npz): velocities bit-exact,
    headings / omega to the last ulps of atan2."""
    from ros2_mpc_b200 import references as rf
    g = np.load(os.path.join(G, "refgen_golden.npz"))
    for pi in range(int(g["n_paths"])):
        h, v, o = rf.get_headings(g[f"path{pi}_xy"], 0.2)
        assert np.array_equal(v, g[f"path{pi}_velocity"])
        assert np.max(np.abs(h - g[f"path{pi}_heading"])) <= 1e-15
        assert np.max(np.abs(o - g[f"path{pi}_omega"])) <= 1e-15
        assert len(o) == len(h) - 1
    # a batch of paths in one launch
    xy = np.stack([g["path0_xy"], g["path0_xy"][::-1]])
    hb, vb, ob_ = rf.get_headings(xy, 0.2)
    assert hb.shape == (2, len(g["path0_xy"])) and np.array_equal(vb[0], g["path0_velocity"])


def test_window_maximum_by_doubling_is_the_window_maximum():
    """The strip kernels' register scheme (costmap_kernel.cuh: dil_window_max / lcm_window_or), restated: in-place doubling to the
    largest power of two P <= K, then one combine of two overlapping windows of P.  Any K up to 32, maxima and ORs."""
    rng = np.random.default_rng(3)
    nout = 20
    for K in range(1, 33):
        ln = nout + K - 1
        P = 1 << (K.bit_length() - 1)
        for op, v in ((max, list(rng.integers(-1000, 1000, ln))), (lambda a, b: a | b, list(rng.integers(0, 1 << 32, ln)))):
            ref = v[:]
            want = []
            for t in range(nout):
                acc = ref[t]
                for j in range(1, K):
                    acc = op(acc, ref[t + j])
                want.append(acc)
            q = 1
            while q < P:
                for j in range(ln):
                    if j + 2 * q <= ln:
                        v[j] = op(v[j], v[j + q])
                q <<= 1
            if K > P:
                for t in range(nout):
                    v[t] = op(v[t], v[t + K - P])
            assert v[:nout] == want, K
