"""Host-side logic that needs no GPU: params lookup, variant tables, workload generators, bench arithmetic."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from ros2_mpc_b200 import load_params, make_params, params as P, synth, _shim
from ros2_mpc_b200.sharding import contiguous_shard


def test_params_file_has_the_reference_keys(params):
    assert params["dt"] == 0.2 and params["N"] == 30
    assert params["Q"] == [1.0, 1.0, 0.005] and params["R"] == [1.0, 1.0]
    assert P.obstacle_slots(params) == 160
    for k in P.REQUIRED_KEYS:
        assert k in params


def test_params_lookup_order(tmp_path, monkeypatch):
    share = tmp_path / "share"
    (share / "config").mkdir(parents=True)
    (share / "config" / "params.yaml").write_text(open(P._PACKAGED).read().replace("N: 30", "N: 12"))
    monkeypatch.setenv("ROS2_MPC_SHARE", str(share))
    assert load_params()["N"] == 12
    monkeypatch.delenv("ROS2_MPC_SHARE")
    assert load_params()["N"] == 30
    assert load_params(str(share / "config" / "params.yaml"))["N"] == 12
    bad = tmp_path / "bad.yaml"
    bad.write_text("dt: 0.2\n")
    with pytest.raises(KeyError):
        load_params(str(bad))


def test_reference_params_file_is_read_unchanged_when_present():
    ref = "/root/reference/config/params.yaml"
    if not os.path.exists(ref):
        pytest.skip("reference checkout not present")
    assert load_params(ref) == load_params()


@pytest.mark.parametrize("variant,obstacles", [("A", None), ("B", None), ("B", True), ("C", None)])
def test_product_and_oracle_variant_tables_agree(params, built, variant, obstacles):
    """Two independently written tables (ros2_mpc_b200/variants.py, oracle/oracle.py) of SURVEY.md App. A."""
    a = make_params(variant, params, obstacles=obstacles)
    b = O.variant_params(variant, params, obstacles=obstacles)
    for f in ("N", "M", "dt", "integrator", "ref_kind", "kappa", "obs_form", "obs_k0", "obs_k1", "max_iter",
              "obs_r", "tol", "acceptable_tol", "mu_init", "acceptable_iter", "max_soc"):
        assert getattr(a, f) == getattr(b, f), f
    if a.obs_form:
        assert a.obs_c == b.obs_c
    for f in ("Q", "R", "u_lo", "u_hi"):
        assert list(getattr(a, f)) == list(getattr(b, f)), f


def test_variant_constants(params, built):
    a, b, c = (make_params(v, params) for v in "ABC")
    assert list(b.u_lo) == [-0.05, -0.2] and list(b.u_hi) == [0.15, 0.2] and list(b.R) == [0.5, 0.5]
    assert b.kappa == 0.5 and b.obs_form == _shim.OBS_NONE
    assert list(a.Q) == [0.00005, 0.05, 0.05] and a.obs_form == _shim.OBS_EXPLOG and a.obs_c == 5.0
    assert (a.obs_k0, a.obs_k1) == (0, 30)
    assert c.integrator == _shim.EULER and c.ref_kind == _shim.REF_TRAJ and c.kappa == 5.0
    with pytest.raises(ValueError):
        make_params("D", params)


def test_synth_workload_is_seeded_and_well_formed():
    m = synth.load_map()
    assert m["occ"].shape == (224, 314) and int(m["occ"].sum()) == 3452
    a = synth.robots_on_map(B=24, seed=0, m=m)
    b = synth.robots_on_map(B=24, seed=0, m=m)
    for k in ("x0", "goal", "obs_x", "obs_y"):
        assert np.array_equal(a[k], b[k])
    r, c = synth.world_to_cell(m, a["x0"][:, :2])
    assert m["free"][r, c].all()
    d = np.hypot(*(a["goal"][:, :2] - a["x0"][:, :2]).T)
    assert (d >= 0.3 - 1e-9).all() and (d <= 1.0 + 1e-9).all()
    assert a["obs_x"].shape == (24, 160)
    ui = synth.warm_start_seeds(3, 30, [-0.05, -0.2], [0.15, 0.2])
    assert ui.shape == (3, 30, 2) and ui[..., 0].max() <= 0.15 and ui[..., 0].min() >= -0.05
    ox, oy = synth.dense_obstacle_field(a["x0"][:4])
    dd = np.hypot(ox - a["x0"][:4, :1], oy - a["x0"][:4, 1:2])
    assert (dd >= 0.3 - 1e-9).all() and (dd <= 1.5 + 1e-9).all()
    assert all(len({(round(x, 6), round(y, 6)) for x, y in zip(ox[i], oy[i])}) == 160 for i in range(4))


def test_raycast_hits_walls():
    m = synth.load_map()
    clr = synth.clearance(m)
    r, c = np.unravel_index(np.argmax(clr * m["free"]), clr.shape)
    pos = np.array([[m["origin"][0] + (c + 0.5) * 0.05, m["origin"][1] + (r + 0.5) * 0.05]])
    scan, ang = synth.raycast(m, pos, np.array([0.0]))
    assert scan.shape == (1, 360) and scan.min() >= clr[r, c] - 0.1
    assert (scan < 3.5).any()


def test_contiguous_shards_partition_the_batch():
    for B in (0, 1, 7, 4096, 1048576):
        for world in (1, 2, 3, 8):
            spans = [contiguous_shard(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1


def test_bench_flop_model(params, built):
    import bench
    b = make_params("B", params)
    a = make_params("A", params)
    it = np.array([1]); ls = np.array([0])
    assert bench.algorithmic_flops(b, it, ls)[0] == pytest.approx(548 * 30 + 20 * 7 * 30)       # 2.06e4
    assert bench.algorithmic_flops(a, it, ls)[0] == pytest.approx(548 * 30 + 34 * 31 * 160 + 20 * (210 + 31 * 160))
    inb, outb = bench.io_bytes_per_solve(a, True)
    assert inb == 8 * (3 + 3 + 2 * 30 + 2 * 160)


def test_gpu_only_helpers_fail_loudly_without_a_device(built):
    """The producers around the solve have no CPU fallback either: without a CUDA device the drop-ins raise, they do
    not quietly run the numpy mirror."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from ros2_mpc_b200 import obstacles as ob, references as rf
    ob._SOLVERS.clear()
    scan = np.full(360, 1.0)
    with pytest.raises(RuntimeError, match="no CUDA device|CPU fallback"):
        ob.get_obstacles_gpu(scan, np.array([0.0, 6.28]), 2.0, 0.05, np.zeros(2), np.zeros(3), np.ones(160), np.ones(160))
    with pytest.raises(RuntimeError, match="no CUDA device|CPU fallback"):
        rf.get_goal_for_mpc(np.zeros((4, 2)), np.zeros(4), np.zeros(5), np.zeros(2))
    # the numpy mirror itself (workload generation, checker) of course runs
    ox, oy, cnt = ob.get_obstacles(scan, np.array([0.0, 6.28]), 2.0, 0.05, np.zeros((1, 2)), np.zeros(1), 160)
    assert ox.shape == (1, 160) and cnt[0] > 0


def test_bench_strong_scaling_shards_tile_the_batch():
    """bench.py's strong-scaling leg: rank r of n takes the seed blocks [r*S/n, (r+1)*S/n) of ONE batch — the shards, put
    end to end, are the whole batch (every array), for the variants with and without per-problem obstacle lists."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    y = load_params()
    robots, seeds, world = 16, 8, 4
    for variant in ("B", "A", "C"):
        whole = bench.build_workload(variant, robots, seeds, 0, y)
        parts = [bench.build_workload(variant, robots, seeds, 0, y, seed_lo=r * seeds // world, seed_hi=(r + 1) * seeds // world)
                 for r in range(world)]
        assert sum(p["B"] for p in parts) == whole["B"] == robots * seeds
        for k in ("x0", "xref", "uref", "u_init", "obs_x", "obs_y"):
            if whole[k] is None:
                assert all(p[k] is None for p in parts)
            else:
                assert np.array_equal(np.concatenate([p[k] for p in parts]), whole[k]), (variant, k)
        # weak scaling: another rank's batch differs in its warm-start seeds only
        other = bench.build_workload(variant, robots, seeds, seeds, y)
        assert np.array_equal(other["x0"], whole["x0"]) and not np.array_equal(other["u_init"], whole["u_init"])
