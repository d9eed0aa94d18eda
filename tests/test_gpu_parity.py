"""GPU parity tests: the CUDA path (through the C ABI / the drop-in Mpc classes) against the CPU oracle and the
golden fixtures.  Tolerances are the ones BASELINE.json states: <= 1e-5 relative on the optimal cost, <= 1e-4
absolute on u0 and the predicted states, identical status."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
COST_RTOL, U_ATOL, X_ATOL = 1e-5, 1e-4, 1e-4


@pytest.fixture(scope="module")
def env(params, built):
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    from oracle import oracle as O
    from ros2_mpc_b200 import _shim, make_params, synth
    return dict(O=O, shim=_shim, make=make_params, synth=synth, y=params)


@pytest.fixture(scope="module")
def robots(env):
    return env["synth"].robots_on_map(B=384, seed=0)


def _inputs(env, variant, w):
    kw, xr = {}, w["goal"]
    if variant == "A":
        kw = dict(obs_x=w["obs_x"], obs_y=w["obs_y"])
    if variant == "C":
        pxf, puf = env["synth"].straight_reference(w["x0"], w["goal"], env["y"]["N"])
        xr, kw = pxf, dict(uref=puf)
    return xr, kw


def _assert_parity(out, ref, need_frac=0.5, nonconvex_slack=0.0, certify=None):
    """Status identical on every problem; on converged problems cost / controls / states within tolerance.

    nonconvex_slack: fraction of the problems allowed to differ — in status, or (both converged) by sitting on a
    *different* local optimum.  Only the obstacle cost forms use it: exp(c/s) makes the NLP non-convex and
    ill-conditioned, the iteration is chaotic near obstacle points, and ulp-level differences (CUDA vs glibc exp)
    occasionally steer the two implementations apart.  Variants B / C must match on every problem (slack 0).
    certify(b, X, U) -> oracle kkt_certificate dict: when given, every problem that converged in both solutions but to
    different points must satisfy IPOPT's (scaled) first-order conditions in BOTH — "a different local optimum" is
    proven, not assumed.  Returns (problems with another status, problems on another optimum)."""
    n = len(ref["status"])
    dstat = out["status"] != ref["status"]
    ok = np.isin(ref["status"], (0, 1)) & np.isin(out["status"], (0, 1))
    assert np.isin(ref["status"], (0, 1)).mean() >= need_frac
    with np.errstate(invalid="ignore"):
        dc = np.abs(out["cost"] - ref["cost"]) / np.abs(ref["cost"])
        dU = np.abs(out["U"] - ref["U"]).reshape(n, -1).max(1)
        dX = np.abs(out["X"] - ref["X"]).reshape(n, -1).max(1)
        bad = ok & ~((dc <= COST_RTOL) & (dU <= U_ATOL) & (dX <= X_ATOL))
    assert (bad | dstat).sum() <= nonconvex_slack * n, (int(dstat.sum()), int(bad.sum()), n, out["status"][dstat],
                                                       ref["status"][dstat], dc[bad], dU[bad], dX[bad])
    if certify is not None:
        for b in np.where(bad)[0]:
            for sol in (out, ref):
                c = certify(int(b), sol["X"][b], sol["U"][b])
                tol = 1e-6 if sol["status"][b] == 0 else 1e-4   # tol 1e-8 / acceptable_tol 1e-6, recomputed independently
                assert c["defect"] <= 1e-4 and c["complementarity"] <= tol, (int(b), int(sol["status"][b]), c)
    assert np.array_equal(out["X"][:, 0, :], ref["X"][:, 0, :])  # x_opt[:,0] == x0 bit-exact
    return int(dstat.sum()), int(bad.sum())


@pytest.mark.parametrize("variant", ["A", "B", "C"])
def test_nlp_functions_match_oracle(env, robots, variant):
    """K1-K3 (transcription, dynamics + derivatives, obstacle cost) element-wise against the oracle."""
    O, shim = env["O"], env["shim"]
    p, po = env["make"](variant, env["y"]), O.variant_params(variant, env["y"])
    S = shim.Solver(p)
    rng = np.random.default_rng(4)
    B, N = 48, p.N
    w = {k: v[:B] for k, v in robots.items() if isinstance(v, np.ndarray) and v.shape[0] == 384}
    xr, kw = _inputs(env, variant, w)
    X = rng.normal(0, 0.3, (B, N + 1, 3)) + w["x0"][:, None, :]
    X[:, 0, :] = w["x0"]
    U = rng.uniform(-0.2, 0.2, (B, N, 2)); lam = rng.normal(0, 1, (B, N, 3))
    g = S.eval_batch(w["x0"], xr, X, U, lam=lam, obj_scale=0.7, **kw)
    for b in range(B):
        o = O.evaluate(po, w["x0"][b], xr[b], X[b], U[b], lam=lam[b], obj_scale=0.7, **{k: v[b] for k, v in kw.items()})
        for key in ("f", "c", "grad", "stages"):
            a, r = np.asarray(g[key][b], dtype=float), np.asarray(o[key], dtype=float)
            fin = np.isfinite(r)
            assert np.array_equal(np.isfinite(a), fin)
            assert np.all(np.abs(a[fin] - r[fin]) <= 1e-11 * (1 + np.abs(r[fin]))), (variant, key, b)
    S.close()


@pytest.mark.parametrize("variant", ["A", "B", "C"])
def test_nlp_functions_match_reference_golden(env, variant):
    """The kernel's objective / defects == the reference's own Opti problem (nlp_golden.npz)."""
    nlp = np.load(os.path.join(G, "nlp_golden.npz"))
    v = variant
    S = env["shim"].Solver(env["make"](v, env["y"]))
    P = nlp[f"{v}_X"].shape[0]
    x0 = np.tile(nlp[f"{v}_x0"], (P, 1))
    kw, xr = {}, np.tile(nlp[f"{v}_goal"], (P, 1))
    if v == "C":
        xr, kw = np.tile(nlp["C_pf"], (P, 1)), dict(uref=np.tile(nlp["C_puf"], (P, 1)))
    if v == "A":
        kw = dict(obs_x=nlp["A_obs_x"], obs_y=nlp["A_obs_y"])  # shared list, obs_stride 0
    g = S.eval_batch(x0, xr, nlp[f"{v}_X"], nlp[f"{v}_U"], **kw)
    assert np.all(np.abs(g["f"] - nlp[f"{v}_f"]) <= 1e-12 * np.abs(nlp[f"{v}_f"]))
    assert np.allclose(g["c"], nlp[f"{v}_c"], rtol=0, atol=1e-13)
    S.close()


@pytest.mark.parametrize("variant", ["A", "B", "C"])
def test_solve_matches_oracle_on_map_problems(env, robots, variant):
    """Config 3 problems (random poses / goals on map_carto, scan-derived obstacle lists)."""
    O = env["O"]
    xr, kw = _inputs(env, variant, robots)
    S = env["shim"].Solver(env["make"](variant, env["y"]))
    out = S.solve_batch(robots["x0"], xr, **kw)
    ref = O.solve_batch(O.variant_params(variant, env["y"]), robots["x0"], xr, **kw)
    po = O.variant_params(variant, env["y"])
    cert = None
    if variant == "A":
        cert = lambda b, X, U: O.kkt_certificate(po, robots["x0"][b], xr[b], X, U, obs_x=kw["obs_x"][b], obs_y=kw["obs_y"][b])  # noqa: E731
    # variant A: measured 0.3 % of the converged problems end on another (certified) KKT point; 1 % is the ceiling
    nst, nopt = _assert_parity(out, ref, nonconvex_slack=0.01 if variant == "A" else 0.0, certify=cert)
    print(f"variant {variant}: {nst} of {len(ref['status'])} problems with another status, {nopt} converged to another optimum")
    if variant in "BC":
        assert np.isin(ref["status"], (0, 1)).all()
        assert np.array_equal(out["iters"], ref["iters"])
    # active set on converged problems: same controls at their bounds
    ok = (ref["status"] == 0) & (np.abs(out["U"] - ref["U"]).reshape(len(ref["status"]), -1).max(1) <= U_ATOL)
    p = S.params
    for i in range(2):
        at_hi = lambda U: np.abs(U[ok][:, :, i] - p.u_hi[i]) <= 1e-6  # noqa: E731
        at_lo = lambda U: np.abs(U[ok][:, :, i] - p.u_lo[i]) <= 1e-6  # noqa: E731
        # a control may sit within 1e-6 of the bound in one solution and 1.1e-6 in the other: compare with slack
        assert (at_hi(out["U"]) != at_hi(ref["U"])).mean() <= 1e-3
        assert (at_lo(out["U"]) != at_lo(ref["U"])).mean() <= 1e-3
    S.close()


def test_solve_matches_dense_kkt_oracle(env, robots):
    """Against the oracle's dense LDL^T backend (no Riccati structure shared with the kernel)."""
    O = env["O"]
    B = 24
    S = env["shim"].Solver(env["make"]("B", env["y"]))
    out = S.solve_batch(robots["x0"][:B], robots["goal"][:B])
    ref = O.solve_batch(O.variant_params("B", env["y"], linear_solver=1), robots["x0"][:B], robots["goal"][:B])
    _assert_parity(out, ref, need_frac=1.0)
    S.close()


def test_config1_and_reference_golden_solutions(env):
    """Config 1 (perform_mpc defaults) and the optima the reference's own perform_mpc returned with SLSQP."""
    sol = np.load(os.path.join(G, "solve_golden.npz"))
    from ros2_mpc_b200 import MpcPointStabilization, MpcPointStabilizationLocal, MpcTracking
    N = 30
    u0 = np.zeros((2, N))
    mb = MpcPointStabilizationLocal()
    assert (mb.N, mb.n_controls, mb.n_states, mb.dt) == (30, 2, 3, 0.2)
    for tag in ("B1", "B2", "B3"):
        u = mb.perform_mpc(u0, sol[f"{tag}_x0"], sol[f"{tag}_goal"])
        assert u.shape == (2,) and np.max(np.abs(u - sol[f"{tag}_u0"])) <= U_ATOL
        assert abs(mb.last_cost - float(sol[f"{tag}_cost"])) <= COST_RTOL * abs(float(sol[f"{tag}_cost"]))
    u = mb.perform_mpc(u0)  # defaults: x0 = 0, goal = (10,10,0)
    assert np.max(np.abs(u - sol["B2_u0"])) <= U_ATOL
    ma = MpcPointStabilization()
    with pytest.raises(RuntimeError):
        ma.perform_mpc(u0)  # obstacle parameters never set while the cost is active
    for tag in ("A1", "A2"):
        x_opt, u_opt = ma.perform_mpc(u0, sol[f"{tag}_x0"], sol[f"{tag}_goal"], sol[f"{tag}_obs_x"], sol[f"{tag}_obs_y"])
        assert x_opt.shape == (3, N + 1) and u_opt.shape == (2, N)
        assert np.max(np.abs(x_opt - sol[f"{tag}_X"])) <= X_ATOL and np.max(np.abs(u_opt - sol[f"{tag}_U"])) <= U_ATOL
        assert abs(ma.last_cost - float(sol[f"{tag}_cost"])) <= COST_RTOL * float(sol[f"{tag}_cost"])
        assert np.array_equal(x_opt[:, 0], sol[f"{tag}_x0"])
    # obstacles=None: the values of the previous call persist (opti.set_value is simply not repeated): solving A2's problem
    # again without passing its wall must give A2's solution, not the sentinel-obstacle one
    x_opt, _ = ma.perform_mpc(u0, sol["A2_x0"], sol["A2_goal"])
    assert np.max(np.abs(x_opt - sol["A2_X"])) <= X_ATOL
    mc = MpcTracking()
    x_opt, u0c = mc.perform_mpc(u0, sol["C1_x0"], sol["C1_pf"].reshape(-1, 1), sol["C1_puf"].reshape(-1, 1))
    assert np.max(np.abs(x_opt - sol["C1_X"])) <= X_ATOL and np.max(np.abs(u0c - sol["C1_u0"])) <= U_ATOL
    for m in (ma, mb, mc):
        m.close()


def test_failed_solve_raises_like_opti_solve(env):
    from ros2_mpc_b200 import MpcPointStabilization, SolveError
    ma = MpcPointStabilization()
    ox = np.full(160, 100.0); oy = np.full(160, 100.0)
    ox[3], oy[3] = 0.001, 0.0  # on the Opti start guess X = 0: exp(c/s) overflows
    with pytest.raises(RuntimeError, match="Invalid_Number_Detected") as ei:
        ma.perform_mpc(np.zeros((2, 30)), np.array([1.0, 1.0, 0.0]), np.array([2.0, 2.0, 0.0]), ox, oy)
    assert isinstance(ei.value, SolveError) and ei.value.status == -13
    r = ma.perform_mpc_batch(None, np.array([[1.0, 1.0, 0.0]]), np.array([[2.0, 2.0, 0.0]]), ox, oy)
    assert r["status"][0] == -13  # batch API reports, does not raise
    ma.close()


@pytest.mark.parametrize("field", ["stated", "easier"])
@pytest.mark.parametrize("N", [10, 25, 50, 100])
def test_horizon_sweep_config5(env, N, field):
    """Config 5: N in {10,25,50,100}, halved control box, 160 distinct obstacle points, variant-A cost form.
    "stated": the field SURVEY section 8d specifies (annulus 0.3-1.5 m around the start, IPOPT's max_iter 3000);
    "easier": annulus 0.6-1.5 m and max_iter 300 (round 1's case, kept as a second data point)."""
    O = env["O"]
    B = 48
    w = env["synth"].robots_on_map(B=B, seed=5)
    r_in, max_iter = (0.3, 3000) if field == "stated" else (0.6, 300)
    ox, oy = env["synth"].dense_obstacle_field(w["x0"], seed=2, r_in=r_in)
    over = dict(u_lo=[-0.025, -0.1], u_hi=[0.075, 0.1], max_iter=max_iter)
    p = env["make"]("A", env["y"], N=N, **over)
    po = O.variant_params("A", env["y"], N=N, **over)
    po.obs_k1 = N
    assert p.obs_k1 == N
    S = env["shim"].Solver(p)
    out = S.solve_batch(w["x0"], w["goal"], obs_x=ox, obs_y=oy)
    ref = O.solve_batch(po, w["x0"], w["goal"], obs_x=ox, obs_y=oy)
    cert = lambda b, X, U: O.kkt_certificate(po, w["x0"][b], w["goal"][b], X, U, obs_x=ox[b], obs_y=oy[b])  # noqa: E731
    # the stated field at N = 100: every predicted position sits inside the obstacle annulus, a fifth of the problems converge.
    # The stated field starts 0.3 m from the nearest obstacle point: the iteration is chaotic there, and with 48 problems the
    # count of problems that an ulp in exp() tips to another status / optimum moves between 2 and 4 from build to build
    # (each "another optimum" is certified as a KKT point below): 10 % for the stated field, 5 % for the easier one.
    hard = field == "stated" and N == 100
    nst, nopt = _assert_parity(out, ref, need_frac=0.15 if hard else 0.3, nonconvex_slack=0.1 if field == "stated" else 0.05, certify=cert)
    print(f"config 5 {field} N={N}: converged {np.isin(ref['status'], (0, 1)).mean():.2f}, another status {nst}, another optimum {nopt}, "
          f"status counts {dict(zip(*np.unique(out['status'], return_counts=True)))}")
    S.close()


def test_gauss_obstacle_form_matches_oracle_and_reference_sources(env, robots):
    """Variant B with its obstacle cost enabled (obstacles=True): the gauss form c*exp(-s) the reference builds in
    define_obstacles_cost_function (local_planner_point_stabilization.py:60-67) and then drops from the objective.
    (1) the evaluation kernel's obstacle cost against values of the reference's own method traced through the casadi
    stand-in (nlp_golden.npz, Bobs_*); (2) value / gradient / Hessian against the oracle; (3) solves against the oracle."""
    O, shim = env["O"], env["shim"]
    g = np.load(os.path.join(G, "nlp_golden.npz"))
    p, p0 = env["make"]("B", env["y"], obstacles=True), env["make"]("B", env["y"])
    assert p.obs_form == shim.OBS_GAUSS and (p.obs_k0, p.obs_k1) == (0, p.N - 1)
    S, S0 = shim.Solver(p), shim.Solver(p0)
    Xs, Us = g["Bobs_X"], g["Bobs_U"]
    n = Xs.shape[0]
    x0 = np.tile(g["Bobs_x0"], (n, 1)); goal = np.tile([1.1, 0.2, 0.5], (n, 1))
    f_with = S.eval_batch(x0, goal, Xs, Us, obs_x=g["Bobs_obs_x"], obs_y=g["Bobs_obs_y"])["f"]
    f_without = S0.eval_batch(x0, goal, Xs, Us)["f"]
    assert np.max(np.abs((f_with - f_without) - g["Bobs_f"]) / g["Bobs_f"]) <= 1e-12
    S0.close()
    po = O.variant_params("B", env["y"], obstacles=True)
    B = 64
    w = {k: v[:B] for k, v in robots.items() if isinstance(v, np.ndarray) and v.shape[0] == 384}
    rng = np.random.default_rng(8)
    X = rng.normal(0, 0.3, (B, p.N + 1, 3)) + w["x0"][:, None, :]
    X[:, 0, :] = w["x0"]
    U = rng.uniform(-0.05, 0.15, (B, p.N, 2)); lam = rng.normal(0, 1, (B, p.N, 3))
    ev = S.eval_batch(w["x0"], w["goal"], X, U, lam=lam, obj_scale=0.3, obs_x=w["obs_x"], obs_y=w["obs_y"])
    for b in range(0, B, 4):
        o = O.evaluate(po, w["x0"][b], w["goal"][b], X[b], U[b], lam=lam[b], obj_scale=0.3, obs_x=w["obs_x"][b], obs_y=w["obs_y"][b])
        for key in ("f", "c", "grad", "stages"):
            a, r = np.asarray(ev[key][b], dtype=float), np.asarray(o[key], dtype=float)
            assert np.all(np.abs(a - r) <= 1e-11 * (1 + np.abs(r))), (key, b)
    out = S.solve_batch(w["x0"], w["goal"], obs_x=w["obs_x"], obs_y=w["obs_y"])
    ref = O.solve_batch(po, w["x0"], w["goal"], obs_x=w["obs_x"], obs_y=w["obs_y"])
    cert = lambda b, X_, U_: O.kkt_certificate(po, w["x0"][b], w["goal"][b], X_, U_, obs_x=w["obs_x"][b], obs_y=w["obs_y"][b])  # noqa: E731
    nst, nopt = _assert_parity(out, ref, need_frac=0.9, nonconvex_slack=0.02, certify=cert)
    print(f"gauss form: converged {np.isin(ref['status'], (0, 1)).mean():.2f}, another status {nst}, another optimum {nopt}")
    # the obstacle cost matters: the optimum differs from the obstacle-free variant B on most problems
    S.close()


def test_edge_cases(env, robots):
    O, shim = env["O"], env["shim"]
    p = env["make"]("B", env["y"])
    S = shim.Solver(p)
    # empty batch
    out = S.solve_batch(np.zeros((0, 3)), np.zeros((0, 3)))
    assert out["X"].shape == (0, 31, 3)
    # batch of one, ragged batch sizes around the 4-warps-per-CTA granularity
    for B in (1, 3, 5, 129):
        o = S.solve_batch(robots["x0"][:B], robots["goal"][:B])
        r = O.solve_batch(O.variant_params("B", env["y"]), robots["x0"][:B], robots["goal"][:B])
        _assert_parity(o, r, need_frac=1.0)
    # goal == start; start guess on / outside the bounds (slack push); far goal (inertia correction path)
    x0 = np.array([[0.3, 0.4, 1.0], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]])
    goal = np.array([[0.3, 0.4, 1.0], [10.0, 10.0, 0.0], [-30.0, 40.0, 3.0]])
    ui = np.zeros((3, 30, 2)); ui[0, :, 0] = 0.15; ui[0, :, 1] = -0.2; ui[1, :, 0] = 5.0; ui[2, ::2, 1] = -7.0
    o = S.solve_batch(x0, goal, u_init=ui)
    r = O.solve_batch(O.variant_params("B", env["y"]), x0, goal, u_init=ui.reshape(3, -1))
    _assert_parity(o, r, need_frac=1.0)
    S.close()
    # shared obstacle list (obs_stride = 0) == the same list replicated per problem
    pa = env["make"]("A", env["y"])
    Sa = shim.Solver(pa)
    B = 16
    x0 = np.tile([[0.0, 0.0, 0.0]], (B, 1)); goal = np.c_[np.linspace(0.5, 1.5, B), np.linspace(-0.3, 0.3, B), np.zeros(B)]
    wx, wy = np.linspace(-1, 2, 160), np.full(160, 1.0)
    a = Sa.solve_batch(x0, goal, obs_x=wx, obs_y=wy)
    b = Sa.solve_batch(x0, goal, obs_x=np.tile(wx, (B, 1)), obs_y=np.tile(wy, (B, 1)))
    assert np.array_equal(a["X"], b["X"]) and np.array_equal(a["status"], b["status"])
    with pytest.raises(RuntimeError, match="obstacle"):
        Sa.solve_batch(x0, goal)
    Sa.close()


def test_max_iter_status(env, robots):
    O = env["O"]
    p = env["make"]("B", env["y"], max_iter=5)
    S = env["shim"].Solver(p)
    out = S.solve_batch(robots["x0"][:32], robots["goal"][:32])
    ref = O.solve_batch(O.variant_params("B", env["y"], max_iter=5), robots["x0"][:32], robots["goal"][:32])
    assert (out["status"] == -1).all() and np.array_equal(out["status"], ref["status"])
    assert (out["iters"] == 5).all()
    assert np.allclose(out["X"], ref["X"], atol=1e-9)
    S.close()


def test_full_size_properties_config4(env):
    """At BASELINE sizes the oracle is too slow, so check size-independent properties on 4096 robots x 64 seeds
    (262,144 problems, the lane-per-problem kernel): every problem converges; the warm-start seeds of a robot reach
    the optimum of its cold start (cost <= 1e-5 rel, u0 / states <= 1e-4) except on the few robots whose NLP has
    several local optima (turn left / turn right); the two solve kernels agree problem by problem; repeated
    launches are bit-identical; the device-buffer and host-buffer entry points agree; a sample matches the oracle."""
    import torch
    O, shim, synth = env["O"], env["shim"], env["synth"]
    R, Sd, N = 4096, 64, 30
    w = synth.robots_on_map(B=R, seed=0)
    p = env["make"]("B", env["y"])
    S = shim.Solver(p)
    ui = synth.warm_start_seeds(Sd, N, list(p.u_lo), list(p.u_hi))
    B = R * Sd
    x0 = np.tile(w["x0"], (Sd, 1)); goal = np.tile(w["goal"], (Sd, 1))
    u_init = np.repeat(ui, R, axis=0).reshape(B, N, 2)
    out = S.solve_batch(x0, goal, u_init=u_init)
    assert S.last_kernel_kind == shim.KERNEL_LANE
    assert np.isin(out["status"], (0, 1)).all()
    cold = S.solve_batch(w["x0"], w["goal"])
    X = out["X"].reshape(Sd, R, N + 1, 3); U = out["U"].reshape(Sd, R, N, 2); c = out["cost"].reshape(Sd, R)
    same = ((np.abs(c - cold["cost"][None]) / cold["cost"][None] <= COST_RTOL)
            & (np.abs(U - cold["U"][None]).reshape(Sd, R, -1).max(2) <= U_ATOL)
            & (np.abs(X - cold["X"][None]).reshape(Sd, R, -1).max(2) <= X_ATOL))
    assert same.mean() >= 0.97          # measured: 98.4 % of the (robot, seed) pairs
    assert same.all(0).mean() >= 0.90   # measured: 94 % of the robots are unimodal over all 64 seeds
    # the warp-per-problem kernel on the same batch: same status, same optimum, same iteration path
    S.set_kernel(shim.KERNEL_WARP)
    wout = S.solve_batch(x0, goal, u_init=u_init)
    S.set_kernel(shim.KERNEL_AUTO)
    _assert_parity(out, wout, need_frac=1.0)
    assert (out["iters"] == wout["iters"]).mean() >= 0.999
    again = S.solve_batch(x0, goal, u_init=u_init)
    assert np.array_equal(again["X"], out["X"]) and np.array_equal(again["U"], out["U"])
    # device-buffer entry point on torch's stream
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    dx0, dg, dui = t(x0), t(goal), t(u_init)
    dX = torch.empty((B, N + 1, 3), dtype=torch.float64, device=dev); dU = torch.empty((B, N, 2), dtype=torch.float64, device=dev)
    dc = torch.empty(B, dtype=torch.float64, device=dev)
    ds = torch.empty(B, dtype=torch.int32, device=dev); di = torch.empty_like(ds); dl = torch.empty_like(ds)
    S.solve_batch_device(B, dx0.data_ptr(), dg.data_ptr(), 0, 0, 0, 0, dui.data_ptr(), dX.data_ptr(), dU.data_ptr(),
                         dc.data_ptr(), ds.data_ptr(), di.data_ptr(), dl.data_ptr(),
                         stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(dX.cpu().numpy(), out["X"]) and np.array_equal(ds.cpu().numpy(), out["status"])
    # oracle on a strided sample
    idx = np.arange(0, B, B // 256)
    ref = O.solve_batch(O.variant_params("B", env["y"]), x0[idx], goal[idx], u_init=u_init[idx].reshape(len(idx), -1))
    sub = {k: v[idx] for k, v in out.items()}
    _assert_parity(sub, ref, need_frac=1.0)
    S.close()


@pytest.mark.parametrize("warm", [False, True])
def test_closed_loop_config2(env, warm):
    """Config 2: one robot driven along a path to a goal on map_carto, re-solving every control step with the look-ahead
    goal of get_goal_for_mpc (look_ahead_distance from params.yaml).  cold: zeros as the start guess every step, as the
    reference does (scripts/point_follower_local_planner.py:174).  warm (new feature): the previous plan shifted by one
    stage; GPU and oracle get the same u_init and must agree on that warm solve, step by step."""
    O = env["O"]
    from ros2_mpc_b200 import MpcPointStabilizationLocal, references as rf
    y = env["y"]
    start = np.array([-2.965, 2.315, 0.0])
    goal = np.array([-1.6, 2.9, 0.0, 0.0, 0.4])                      # (x, y, -, -, yaw) as GoalSubscriber delivers it
    mpc = MpcPointStabilizationLocal()
    po = O.variant_params("B", y)
    xg, xo = start.copy(), start.copy()
    N = mpc.N
    u_prev = np.zeros((2, N))
    reached, lookahead_used = False, 0
    for step in range(160):
        # the global planner re-plans from the robot's position every cycle (scripts/global_path_publisher.py:70-133): a
        # straight path from where the robot is; get_goal_for_mpc takes its first point beyond the look-ahead distance
        path = np.linspace(xg[:2], goal[:2], 40)
        head, _, _ = rf.get_headings(path, y["dt"], solver=mpc._solver)
        look = rf.get_goal_for_mpc(path, head, goal, xg[:2], y["look_ahead_distance"], solver=mpc._solver)
        lookahead_used += int(np.linalg.norm(look[:2] - goal[:2]) > 1e-9)
        u_init = u_prev if warm else np.zeros((2, N))
        ug = mpc._solve(u_init, xg, look, None, None)[1]            # full plan (2,N); perform_mpc returns its first column
        ro = O.solve(po, xo, look, u_init=u_init)
        assert ro["status"] == 0 and mpc.last_status == 0
        assert np.max(np.abs(ug - ro["U"])) <= U_ATOL
        assert abs(mpc.last_cost - ro["cost"]) <= COST_RTOL * abs(ro["cost"])
        assert mpc.last_iterations == ro["stats"]["iters"]
        # plant: the same RK4 unicycle, dt = 0.2
        for x, u in ((xg, ug[:, 0]), (xo, ro["U"][:, 0])):
            th, v, w_ = x[2], u[0], u[1]
            tm, te = th + 0.1 * w_, th + 0.2 * w_
            x[0] += 0.2 * v / 6 * (np.cos(th) + 4 * np.cos(tm) + np.cos(te))
            x[1] += 0.2 * v / 6 * (np.sin(th) + 4 * np.sin(tm) + np.sin(te))
            x[2] = te
        u_prev = np.concatenate([ug[:, 1:], ug[:, -1:]], axis=1)    # shifted plan (the GPU's; both solvers get it)
        assert np.max(np.abs(xg - xo)) <= 1e-3
        if np.linalg.norm(xg[:2] - goal[:2]) <= y["goal_threshold"]:
            reached = True
            break
    assert reached and lookahead_used >= 10
    mpc.close()


def test_control_step_matches_reference_statements(env):
    """The limiter / goal-reached logic of control_step_kernel against outputs of the reference's own statements
    (scripts/point_follower_local_planner.py:196-231, cut out of main() with ast and executed:
    tests/golden/make_control_golden.py): 3000 single steps around both thresholds, and 64 sequences of 40 steps with
    u_last / GOAL_FLAG carried on the device."""
    import torch
    g = np.load(os.path.join(G, "control_golden.npz"))
    y, shim = env["y"], env["shim"]
    S = shim.Solver(env["make"]("B", y))
    dev = torch.device("cuda", 0)
    N = y["N"]
    t = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dtype=dt)  # noqa: E731
    thr = float(g["goal_threshold"])

    def run(u, u_last, x0, goal, flag):
        B = u.shape[0]
        U = np.zeros((B, N, 2)); U[:, 0] = u
        dU, dul, dx0, dgoal = t(U), t(u_last), t(x0), t(goal)
        dstate = t(x0); dflag = t(flag.astype(np.int32), torch.int32)
        dstatus = torch.zeros(B, dtype=torch.int32, device=dev); dcmd = torch.zeros((B, 2), dtype=torch.float64, device=dev)
        S.control_step_device(B, dU.data_ptr(), dstatus.data_ptr(), dstate.data_ptr(), dx0.data_ptr(), dul.data_ptr(),
                              dgoal.data_ptr(), 5, dflag.data_ptr(), thr, 0.03, False, dcmd.data_ptr(), 0,
                              stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        return dcmd.cpu().numpy(), dflag.cpu().numpy().astype(bool), dul.cpu().numpy()

    cmd, flag, ul = run(g["s_u"], g["s_u_last"], g["s_x0"], g["s_goal"], g["s_flag"])
    assert np.array_equal(flag, g["s_flag_out"])
    assert np.array_equal(cmd, g["s_cmd"])
    assert np.array_equal(ul, g["s_u_last_out"])
    Q, T = g["q_u"].shape[:2]
    ul, fl = np.zeros((Q, 2)), np.zeros(Q, bool)
    for k in range(T):
        x0 = np.c_[g["q_x0"][:, k], np.zeros(Q)]
        cmd, fl, ul = run(g["q_u"][:, k], ul, x0, g["q_goal"], fl)
        assert np.array_equal(cmd, g["q_cmd"][:, k]) and np.array_equal(fl, g["q_flag"][:, k]), k
        assert np.array_equal(ul, g["q_u_last"][:, k]), k
    S.close()


# ---- lane-per-problem kernel (large-batch path) -----------------------------------------------------------------

@pytest.mark.parametrize("variant", ["B", "C"])
def test_lane_kernel_matches_oracle(env, robots, variant):
    """The lane-per-problem kernel (forced; normally chosen from 30720 problems on) against the oracle on the
    config-3 problems, cold start and with random warm-start seeds."""
    O, shim, synth = env["O"], env["shim"], env["synth"]
    xr, kw = _inputs(env, variant, robots)
    p = env["make"](variant, env["y"])
    S = shim.Solver(p)
    S.set_kernel(shim.KERNEL_LANE)
    po = O.variant_params(variant, env["y"])
    B, N = robots["x0"].shape[0], p.N
    ui = synth.warm_start_seeds(4, N, list(p.u_lo), list(p.u_hi), first_seed=11)
    u_init = np.repeat(ui, B // 4, axis=0).reshape(B, N, 2)
    for u in (None, u_init):
        out = S.solve_batch(robots["x0"], xr, u_init=u, **kw)
        assert S.last_kernel_kind == shim.KERNEL_LANE
        ref = O.solve_batch(po, robots["x0"], xr, u_init=None if u is None else u.reshape(B, -1), **kw)
        _assert_parity(out, ref, need_frac=1.0)
        # same algorithm, rounding-level arithmetic differences: the iteration paths coincide
        assert (out["iters"] == ref["iters"]).mean() >= 0.99
        assert abs(out["ls"].mean() - ref["ls"].mean()) <= 0.02
    S.close()


def test_lane_kernel_agrees_with_warp_kernel(env):
    shim, synth = env["shim"], env["synth"]
    w = synth.robots_on_map(B=4096, seed=3)
    S = shim.Solver(env["make"]("B", env["y"]))
    S.set_kernel(shim.KERNEL_LANE)
    a = S.solve_batch(w["x0"], w["goal"])
    S.set_kernel(shim.KERNEL_WARP)
    b = S.solve_batch(w["x0"], w["goal"])
    assert S.last_kernel_kind == shim.KERNEL_WARP
    _assert_parity(a, b, need_frac=1.0)
    assert (a["iters"] == b["iters"]).mean() >= 0.99
    # automatic choice by batch size
    S.set_kernel(shim.KERNEL_AUTO)
    S.solve_batch(w["x0"][:100], w["goal"][:100])
    assert S.last_kernel_kind == shim.KERNEL_WARP
    mid, big = 5, 8  # 20 480 problems: still the warp kernel; 32 768: the lane kernel
    S.solve_batch(np.tile(w["x0"], (mid, 1)), np.tile(w["goal"], (mid, 1)))
    assert S.last_kernel_kind == shim.KERNEL_WARP
    S.solve_batch(np.tile(w["x0"], (big, 1)), np.tile(w["goal"], (big, 1)))
    assert S.last_kernel_kind == shim.KERNEL_LANE
    S.close()
    # with the obstacle cost the lane kernel takes over later (131 072 problems): 4 096 problems stay on the warp kernel
    Sa = shim.Solver(env["make"]("A", env["y"]))
    Sa.solve_batch(w["x0"][:64], w["goal"][:64], obs_x=w["obs_x"][:64], obs_y=w["obs_y"][:64])
    assert Sa.last_kernel_kind == shim.KERNEL_WARP
    Sa.close()


def test_lane_kernel_with_obstacle_cost_matches_oracle(env, robots, monkeypatch):
    """The lane-per-problem kernel carries the obstacle cost (variant A: mpc_point_stabilization.py:46-53,100): obstacle sums
    by the warp for one problem at a time, cached per stage in the workspace; restoration candidates evaluated by the warp.
    Forced here (AUTO takes it from 131 072 problems on); the instance of variant A and the generic instance (run-time
    switch), cold start and warm-start seeds; then variant B with its gauss obstacle cost (same instance, other form)."""
    O, shim, synth = env["O"], env["shim"], env["synth"]
    y, w = env["y"], robots
    B = w["x0"].shape[0]
    for variant, mk, generic in (("A", {}, False), ("A", {}, True), ("B", dict(obstacles=True), False)):
        p = env["make"](variant, y, **mk)
        po = O.variant_params(variant, y, **mk)
        if generic:
            monkeypatch.setenv("B200MPC_LANE_GENERIC", "1")
        S = shim.Solver(p)
        monkeypatch.delenv("B200MPC_LANE_GENERIC", raising=False)
        S.set_kernel(shim.KERNEL_LANE)
        ui = synth.warm_start_seeds(4, p.N, list(p.u_lo), list(p.u_hi), first_seed=11)
        u_init = np.repeat(ui, B // 4, axis=0).reshape(B, p.N, 2)
        kw = dict(obs_x=w["obs_x"], obs_y=w["obs_y"])
        cert = lambda b, X, U: O.kkt_certificate(po, w["x0"][b], w["goal"][b], X, U, **{k: v[b] for k, v in kw.items()})  # noqa: E731
        for u in (None, u_init):
            out = S.solve_batch(w["x0"], w["goal"], u_init=u, **kw)
            assert S.last_kernel_kind == shim.KERNEL_LANE
            ref = O.solve_batch(po, w["x0"], w["goal"], u_init=None if u is None else u.reshape(B, -1), **kw)
            # non-convex: a few problems may end on another (certified) KKT point or with another status (see _assert_parity)
            nst, nopt = _assert_parity(out, ref, need_frac=0.9, nonconvex_slack=0.02, certify=cert if u is None else None)
            both = np.isin(out["status"], (0, 1)) & np.isin(ref["status"], (0, 1))
            print(f"lane kernel, variant {variant} {mk} generic={generic}: another status {nst}, another optimum {nopt}, "
                  f"iterations identical {np.mean((out['iters'] == ref['iters'])[both]):.3f}")
            assert np.mean((out["iters"] == ref["iters"])[both]) >= 0.9
        S.close()


@pytest.mark.parametrize("variant", ["A", "B", "C"])
def test_straggler_handover_lane_to_warp_kernel(env, robots, variant, monkeypatch):
    """A lane-kernel launch lasts as long as its slowest problem, so the lane kernel exports the complete solver state of a
    problem that has taken B200MPC_HAND_ITER iterations (iterate, multipliers, barrier parameter, filter, counters) and a
    second launch of the warp kernel resumes it at the top of the interior-point loop.  Here the threshold is 3, so nearly
    every problem changes kernels in mid-solve; the results must still be the oracle's, iteration for iteration."""
    O, shim = env["O"], env["shim"]
    xr, kw = _inputs(env, variant, robots)
    po = O.variant_params(variant, env["y"])
    ref = O.solve_batch(po, robots["x0"], xr, **kw)
    monkeypatch.setenv("B200MPC_HAND_ITER", "3")
    S = shim.Solver(env["make"](variant, env["y"]))
    monkeypatch.delenv("B200MPC_HAND_ITER")
    S.set_kernel(shim.KERNEL_LANE)
    n0 = S.launch_count
    out = S.solve_batch(robots["x0"], xr, **kw)
    assert S.launch_count - n0 == 2  # lane kernel + the resuming warp kernel
    S.close()
    _assert_parity(out, ref, need_frac=0.9 if variant == "A" else 1.0, nonconvex_slack=0.02 if variant == "A" else 0.0)
    both = np.isin(out["status"], (0, 1)) & np.isin(ref["status"], (0, 1))
    assert np.mean((out["iters"] == ref["iters"])[both]) >= (0.9 if variant == "A" else 0.99)


@pytest.mark.parametrize("variant", ["A", "B"])
def test_straggler_handover_with_a_full_record_buffer(env, robots, variant, monkeypatch):
    """A lane that wants to leave but finds the record buffer full stays in the lane kernel for good (it re-evaluates its
    point and goes on).  Room for 16 records and a threshold of 3: 16 problems change kernels, the others take the
    buffer-full path; every problem must still match the oracle."""
    O, shim = env["O"], env["shim"]
    xr, kw = _inputs(env, variant, robots)
    ref = O.solve_batch(O.variant_params(variant, env["y"]), robots["x0"], xr, **kw)
    monkeypatch.setenv("B200MPC_HAND_ITER", "3")
    monkeypatch.setenv("B200MPC_HAND_CAP", "16")
    S = shim.Solver(env["make"](variant, env["y"]))
    S.set_kernel(shim.KERNEL_LANE)
    out = S.solve_batch(robots["x0"], xr, **kw)
    S.close()
    _assert_parity(out, ref, need_frac=0.9 if variant == "A" else 1.0, nonconvex_slack=0.02 if variant == "A" else 0.0)
    both = np.isin(out["status"], (0, 1)) & np.isin(ref["status"], (0, 1))
    assert np.mean((out["iters"] == ref["iters"])[both]) >= (0.9 if variant == "A" else 0.99)


def test_lane_kernel_edge_cases(env, robots):
    O, shim = env["O"], env["shim"]
    po = O.variant_params("B", env["y"])
    S = shim.Solver(env["make"]("B", env["y"]))
    S.set_kernel(shim.KERNEL_LANE)
    assert S.solve_batch(np.zeros((0, 3)), np.zeros((0, 3)))["X"].shape == (0, 31, 3)
    # ragged batch sizes around the warp (32 lanes) and CTA granularity
    for B in (1, 31, 33, 383):
        o = S.solve_batch(robots["x0"][:B], robots["goal"][:B])
        r = O.solve_batch(po, robots["x0"][:B], robots["goal"][:B])
        _assert_parity(o, r, need_frac=1.0)
    # goal == start; start guess on / outside the bounds (slack push); far goals (inertia correction, long solves)
    x0 = np.array([[0.3, 0.4, 1.0], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]])
    goal = np.array([[0.3, 0.4, 1.0], [10.0, 10.0, 0.0], [-30.0, 40.0, 3.0]])
    ui = np.zeros((3, 30, 2)); ui[0, :, 0] = 0.15; ui[0, :, 1] = -0.2; ui[1, :, 0] = 5.0; ui[2, ::2, 1] = -7.0
    o = S.solve_batch(x0, goal, u_init=ui)
    r = O.solve_batch(po, x0, goal, u_init=ui.reshape(3, -1))
    _assert_parity(o, r, need_frac=1.0)
    S.close()
    # iteration limit
    S = shim.Solver(env["make"]("B", env["y"], max_iter=5))
    S.set_kernel(shim.KERNEL_LANE)
    o = S.solve_batch(robots["x0"][:64], robots["goal"][:64])
    r = O.solve_batch(O.variant_params("B", env["y"], max_iter=5), robots["x0"][:64], robots["goal"][:64])
    assert (o["status"] == -1).all() and np.array_equal(o["status"], r["status"]) and (o["iters"] == 5).all()
    assert np.allclose(o["X"], r["X"], atol=1e-9)
    S.close()


@pytest.mark.parametrize("N", [10, 50, 100])
def test_lane_kernel_horizon_sweep(env, N):
    """Horizons other than params.yaml's 30 (the workspace is sized per handle), halved control box (config 5)."""
    O, shim, synth = env["O"], env["shim"], env["synth"]
    w = synth.robots_on_map(B=96, seed=6)
    over = dict(u_lo=[-0.025, -0.1], u_hi=[0.075, 0.1], max_iter=300)
    S = shim.Solver(env["make"]("B", env["y"], N=N, **over))
    S.set_kernel(shim.KERNEL_LANE)
    out = S.solve_batch(w["x0"], w["goal"])
    ref = O.solve_batch(O.variant_params("B", env["y"], N=N, **over), w["x0"], w["goal"])
    _assert_parity(out, ref, need_frac=0.9)
    S.close()


@pytest.mark.parametrize("variant", ["B", "C"])
def test_streamed_host_solve_matches_plain(env, variant, monkeypatch):
    """Page-locked host buffers: the batch is streamed (chunked H2D while the kernel runs, per-chunk D2H on
    completion flags).  Every problem must come out bit-identical to the plain copy-in / solve / copy-out path,
    including a ragged last chunk.  (A streamed solve finishes every problem in the lane kernel — the per-chunk flags count
    its results — while the plain path hands its stragglers to the warp kernel, which agrees to rounding, not to the bit:
    the hand-over is switched off for the bit-wise comparison and compared at the parity tolerances below.)"""
    import torch
    monkeypatch.setenv("B200MPC_HAND_ITER", "0")
    shim, synth = env["shim"], env["synth"]
    w = synth.robots_on_map(B=4096, seed=5)
    rep, B = 33, 33 * 4096 - 77
    N = env["y"]["N"]
    x0 = np.tile(w["x0"], (rep, 1))[:B]
    xr = np.tile(w["goal"], (rep, 1))[:B]
    kw = {}
    if variant == "C":
        pxf, puf = synth.straight_reference(w["x0"], w["goal"], N)
        xr = np.tile(pxf, (rep, 1))[:B]
        kw["uref"] = np.tile(puf, (rep, 1))[:B]
    ui = synth.warm_start_seeds(rep, N, [-0.05, -0.2], [0.15, 0.2], first_seed=11)
    kw["u_init"] = np.repeat(ui, 4096, axis=0)[:B]
    S = shim.Solver(env["make"](variant, env["y"]))
    plain = S.solve_batch(x0, xr, **kw)
    assert S.last_kernel_kind == shim.KERNEL_LANE and S.last_solve_chunks == 0

    def pin(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

    out = dict(X=pin(np.zeros((B, N + 1, 3))), U=pin(np.zeros((B, N, 2))), cost=pin(np.zeros(B)),
               status=pin(np.full(B, 99, np.int32)), iters=pin(np.zeros(B, np.int32)), ls=pin(np.zeros(B, np.int32)))
    for rnd in range(2):  # twice: the flags and counters must reset between calls
        for v in out.values():
            v[...] = 7
        st = S.solve_batch(pin(x0), pin(xr), out=out, **{k: pin(v) for k, v in kw.items()})
        assert S.last_solve_chunks >= 2
        for k in ("status", "iters", "ls", "cost", "X", "U"):
            assert np.array_equal(st[k], plain[k]), (k, rnd)
    S.close()
    # default handle: plain path with the straggler hand-over against the streamed result
    monkeypatch.delenv("B200MPC_HAND_ITER")
    S = shim.Solver(env["make"](variant, env["y"]))
    handed = S.solve_batch(x0, xr, **kw)
    S.close()
    _assert_parity(handed, {k: np.array(v) for k, v in st.items()}, need_frac=1.0)
    assert (handed["iters"] == st["iters"]).mean() >= 0.999


# ---- obstacle-list construction on the GPU ----------------------------------------------------
def test_gpu_obstacles_match_reference_golden(env):
    """Cell indexing bit-exact, coordinates <= 1e-12, against the outputs of the reference's own numba helpers
    (tests/golden/obstacles_golden.npz), including NaN / inf beams, the empty scan and the overflow case."""
    from ros2_mpc_b200 import obstacles as ob
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "obstacles_golden.npz"))
    S = g["scans"].shape[0]
    groups = {}
    for i in range(S):  # one lidar geometry per call
        groups.setdefault(tuple(g["angles"][i]), []).append(i)
    for ang, idx in groups.items():
        idx = np.array(idx)
        ox, oy, cnt = ob.get_obstacles_batch_gpu(g["scans"][idx], np.array(ang), 2.0, 0.05, g["pos"][idx], g["yaw"][idx], 160)
        assert np.array_equal(cnt, g["count"][idx])
        assert np.allclose(ox, g["obs_x"][idx], rtol=0, atol=1e-12)
        assert np.allclose(oy, g["obs_y"][idx], rtol=0, atol=1e-12)
    none = np.where(g["count"] == 0)[0]
    assert len(none) >= 1
    over = np.where(g["count"] > 160)[0]
    with pytest.raises(ValueError):
        ob.get_obstacles_gpu(g["scans"][over[0]], g["angles"][over[0]], 2.0, 0.05, g["pos"][over[0]],
                             (0.0, 0.0, g["yaw"][over[0]]), np.ones(160), np.ones(160))
    i = none[0]
    x, y = ob.get_obstacles_gpu(g["scans"][i], g["angles"][i], 2.0, 0.05, g["pos"][i], (0.0, 0.0, g["yaw"][i]),
                                np.ones(160), np.ones(160))
    assert np.all(x == 100.0) and np.all(y == 100.0)


def test_gpu_obstacles_match_numpy_mirror_on_map_scans(env, robots):
    """4096 ray-cast scans of map_carto + random NaN / inf / out-of-range beams: identical cells (the occupied-cell
    set is recovered from the coordinates), identical counts, coordinates to 1e-12; other grid sizes and slot counts."""
    from ros2_mpc_b200 import obstacles as ob
    w = env["synth"].robots_on_map(B=4096, seed=0)
    scan = w["scan"].copy()
    rng = np.random.default_rng(5)
    m = rng.random(scan.shape)
    scan[m < 0.02] = np.inf
    scan[(m >= 0.02) & (m < 0.03)] = np.nan
    scan[(m >= 0.03) & (m < 0.04)] = -np.inf
    scan[(m >= 0.04) & (m < 0.05)] *= -1.0
    scan[7] = np.inf
    scan[8] = 3.5
    for size, res, slots in ((2.0, 0.05, 160), (2.0, 0.05, 40), (1.5, 0.1, 64), (3.0, 0.025, 1024)):
        rx, ry, rc = ob.get_obstacles(scan, w["angles"], size, res, w["x0"][:, :2], w["x0"][:, 2], slots)
        gx, gy, gc = ob.get_obstacles_batch_gpu(scan, w["angles"], size, res, w["x0"][:, :2], w["x0"][:, 2], slots)
        assert np.array_equal(gc, rc), (size, res, slots)
        assert np.max(np.abs(gx - rx)) <= 1e-12 and np.max(np.abs(gy - ry)) <= 1e-12, (size, res, slots)
    assert (rc > slots).any() or slots == 1024


# ---- reference producers on the GPU -----------------------------------------------------------
def test_gpu_reference_producers_match_reference_golden(env):
    """get_goal_for_mpc and get_reference_trajectory: bit-exact against outputs of the reference's own functions
    (tests/golden/refgen_golden.npz: straight, curved, shorter-than-horizon and long paths; robots on path points, near
    the path end, on the odometry raster, far away), shared and per-robot paths, and through the one-robot drop-ins."""
    from ros2_mpc_b200 import references as rf, MpcTracking
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "refgen_golden.npz"))
    mpc = MpcTracking()
    for pi in range(int(g["n_paths"])):
        pxy, head, vel, om = (g[f"path{pi}_{k}"] for k in ("xy", "heading", "velocity", "omega"))
        pos, goal = g[f"path{pi}_pos"], g[f"path{pi}_goal"]
        h2, v2, o2 = rf.get_headings(pxy, 0.2)
        assert np.array_equal(v2, vel) and np.max(np.abs(h2 - head)) <= 1e-15 and np.max(np.abs(o2 - om)) <= 1e-15
        gp, idx = rf.get_goals_batch(pxy, head, goal, pos, 0.5)
        assert np.array_equal(gp, g[f"path{pi}_goal_pose"]), pi
        assert (idx >= -1).all() and (idx < len(pxy)).all()
        pxf, puf, near = rf.get_reference_trajectories_batch(pos, goal, pxy, head, vel, om, mpc._solver)
        assert np.array_equal(pxf, g[f"path{pi}_pxf"]), pi
        assert np.array_equal(puf, g[f"path{pi}_puf"]), pi
        # per-robot copies of the path give the same answers
        R = pos.shape[0]
        gp2, _ = rf.get_goals_batch(np.tile(pxy, (R, 1, 1)), np.tile(head, (R, 1)), goal, pos, 0.5)
        assert np.array_equal(gp2, gp)
        pxf2, puf2, _ = mpc._solver.reftraj_batch(np.tile(pxy, (R, 1, 1)), np.tile(head, (R, 1)), np.tile(vel, (R, 1)),
                                                  np.tile(om, (R, 1)), pos, goal)
        assert np.array_equal(pxf2, pxf) and np.array_equal(puf2, puf)
        # one-robot drop-ins with the reference's signatures and shapes
        one = rf.get_goal_for_mpc(pxy, head.reshape(-1, 1), goal[3], pos[3], 0.5)
        assert one.shape == (3,) and np.array_equal(one, gp[3])
        a, b = rf.get_reference_trajectory(pos[5], goal[5], pxy, head, vel, om, mpc)
        assert a.shape == (90, 1) and b.shape == (60, 1)
        assert np.array_equal(a.ravel(), pxf[5]) and np.array_equal(b.ravel(), puf[5])
    # the produced references feed the tracking solve directly (smooth path 0, a robot standing on the path)
    p0 = np.array([g["path0_xy"][10, 0], g["path0_xy"][10, 1], g["path0_heading"][10]])
    a, b = rf.get_reference_trajectory(p0, g["path0_goal"][0], g["path0_xy"], g["path0_heading"], g["path0_velocity"],
                                       g["path0_omega"], mpc)
    x_opt, u = mpc.perform_mpc(np.zeros((2, mpc.N)), p0, a, b)
    assert x_opt.shape == (3, mpc.N + 1) and np.all(np.isfinite(u))
    mpc.close()


# ---- device-side closed loop ------------------------------------------------------------------------
@pytest.mark.parametrize("warm", [False, True])
def test_fleet_closed_loop_matches_host_loop(env, warm):
    """96 robots, 70 control steps, all on the device (goals -> solve -> limiter / goal logic / plant / next
    measurement) against the node loop written out with numpy on the host (the reference's expressions,
    scripts/point_follower_local_planner.py:172-231, around the same host-buffer solve).  Commands, flags and
    the quantised measurements must agree exactly; some robots reach their goal, some are still under way."""
    import torch  # noqa: F401
    from ros2_mpc_b200 import references as rf
    from ros2_mpc_b200.fleet import FleetPointStabilization
    y, shim = env["y"], env["shim"]
    rng = np.random.default_rng(42)
    B, T, N = 96, 70, y["N"]
    t = np.linspace(0, 1, 60)
    path = np.stack([1.6 * t, 0.4 * np.sin(2.5 * t)], axis=1)
    head, _, _ = rf.get_headings(path, y["dt"])
    start = np.c_[rng.normal(0, 0.15, B), rng.normal(0, 0.15, B), rng.uniform(-0.6, 0.6, B)]
    start[::3, :2] += path[35] + 0.0  # a third of the fleet starts far along the path and arrives within the run
    goal = np.tile(np.r_[path[-1], 0.0, 0.0, head[-1]], (B, 1))
    goal[:, :2] += rng.normal(0, 0.02, (B, 2))
    fleet = FleetPointStabilization(start, goal, path, head, params=y, warm_start=warm, quantise=True)
    fleet.step(T)
    dev = fleet.snapshot()
    fleet.close()

    # the same loop on the host
    S = shim.Solver(env["make"]("B", y))
    state = start.copy()
    x0 = np.round(state, 2); x0[:, 2] = x0[:, 2] % (2 * np.pi)
    u_last = np.zeros((B, 2)); flag = np.zeros(B, dtype=bool); cmd = np.zeros((B, 2)); u_init = None
    for step in range(T):
        goal_mpc, _ = rf.get_goals_batch(path, head, goal, x0[:, :2], y["look_ahead_distance"], solver=S)
        out = S.solve_batch(x0, goal_mpc, u_init=u_init)
        ok = np.isin(out["status"], (0, 1))
        for b in range(B):
            u = out["U"][b, 0] if ok[b] else np.zeros(2)
            if flag[b]:
                cmd[b] = 0.0
            elif np.linalg.norm(u - u_last[b]) > 0.03:
                cmd[b] = u_last[b] + 0.03
                u_last[b] = u
            else:
                cmd[b] = u
                u_last[b] = u
            if np.linalg.norm(x0[b, 0:2] - goal[b, 0:2]) > y["goal_threshold"]:
                flag[b] = False
            elif not flag[b]:
                cmd[b] = 0.0
                flag[b] = True
        th, v, w_ = state[:, 2].copy(), cmd[:, 0], cmd[:, 1]
        tm, te = th + 0.5 * y["dt"] * w_, th + y["dt"] * w_
        state[:, 0] += y["dt"] / 6.0 * v * (np.cos(th) + 4 * np.cos(tm) + np.cos(te))
        state[:, 1] += y["dt"] / 6.0 * v * (np.sin(th) + 4 * np.sin(tm) + np.sin(te))
        state[:, 2] = te
        x0 = np.round(state, 2); x0[:, 2] = x0[:, 2] % (2 * np.pi)
        if warm:
            u_init = np.concatenate([out["U"][:, 1:], out["U"][:, -1:]], axis=1)
            u_init[~ok] = 0.0
    S.close()
    assert np.array_equal(dev["goal_flag"].astype(bool), flag)
    assert flag.any() and not flag.all()
    assert np.array_equal(dev["cmd"], cmd)
    assert np.array_equal(dev["u_last"], u_last)
    assert np.array_equal(dev["x0"], x0)
    assert np.max(np.abs(dev["state"] - state)) <= 1e-9
    assert np.isin(dev["status"], (0, 1)).all()


def test_gpu_producers_edge_cases(env):
    """Empty batches, one-point paths, beam counts that are not multiples of the warp size, a single robot."""
    from ros2_mpc_b200 import obstacles as ob, references as rf
    S = env["shim"].Solver(env["make"]("C", env["y"]))
    N = env["y"]["N"]
    # empty batches are accepted and do nothing
    ox, oy, cnt = S.obstacles_batch(np.zeros((0, 360)), np.ones(360), np.zeros(360), np.zeros((0, 2)), np.zeros(0), 2.0, 0.05, 160)
    assert ox.shape == (0, 160) and cnt.shape == (0,)
    gp, idx = S.goals_batch(np.zeros((4, 2)), np.zeros(4), np.zeros((0, 5)), np.zeros((0, 2)), 0.5)
    assert gp.shape == (0, 3)
    # one-point path: every robot gets that point (or the final goal when it is near)
    pxy, head = np.array([[1.0, 2.0]]), np.array([7.5])
    goal = np.array([[1.0, 2.0, 0, 0, -1.0], [5.0, 5.0, 0, 0, 9.0]])
    pos = np.array([[1.2, 2.1], [0.0, 0.0]])
    gp, idx = S.goals_batch(pxy, head, goal, pos, 0.5)
    assert np.array_equal(gp[0], [1.0, 2.0, (-1.0) % (2 * np.pi)]) and idx[0] == -1
    assert np.array_equal(gp[1], [1.0, 2.0, 7.5 % (2 * np.pi)]) and idx[1] == 0
    pxf, puf, near = S.reftraj_batch(pxy, head, np.array([0.3]), np.array([0.1]), np.array([[4.0, 4.0, 0.0]]), goal[:1, :3])
    assert np.array_equal(pxf[0], np.tile([1.0, 2.0, 7.5], N)) and np.array_equal(puf[0], np.tile([0.3, 0.1], N)) and near[0] == 0
    # beam counts 1, 33, 100, 361, 719 against the numpy mirror (random ranges incl. NaN / inf), one robot and many
    rng = np.random.default_rng(3)
    for n in (1, 33, 100, 361, 719):
        for B in (1, 37):
            scan = np.round(rng.uniform(0.1, 3.0, (B, n)), 2)
            scan[rng.random((B, n)) < 0.05] = np.inf
            scan[rng.random((B, n)) < 0.03] = np.nan
            ang = np.array([-1.0, 5.0])
            pos, yaw = rng.uniform(-3, 3, (B, 2)), rng.uniform(-3, 3, B)
            rx, ry, rc = ob.get_obstacles(scan, ang, 2.0, 0.05, pos, yaw, 160)
            gx, gy, gc = ob.get_obstacles_batch_gpu(scan, ang, 2.0, 0.05, pos, yaw, 160, solver=S)
            assert np.array_equal(gc, rc), (n, B)
            assert np.max(np.abs(gx - rx)) <= 1e-12 and np.max(np.abs(gy - ry)) <= 1e-12, (n, B)
    S.close()


@pytest.mark.parametrize("variant", ["B", "C"])
def test_two_sweep_lane_kernel_matches_oracle(env, robots, variant, monkeypatch):
    """The experimental two-sweep lane kernel (csrc/tpp_fused.cuh, B200MPC_LANE_FUSED=1: trial sweep fused with the
    next iteration's Riccati sweep) solves the same problems to the same optima in the same number of iterations."""
    O, shim, synth = env["O"], env["shim"], env["synth"]
    monkeypatch.setenv("B200MPC_LANE_FUSED", "1")
    S = shim.Solver(env["make"](variant, env["y"]))
    monkeypatch.delenv("B200MPC_LANE_FUSED")
    S.set_kernel(shim.KERNEL_LANE)
    po = O.variant_params(variant, env["y"])
    w = robots
    N = env["y"]["N"]
    xr, kw = w["goal"], {}
    if variant == "C":
        pxf, puf = synth.straight_reference(w["x0"], w["goal"], N)
        xr, kw = pxf, dict(uref=puf)
    B = w["x0"].shape[0]
    ui = synth.warm_start_seeds(B, N, [-0.05, -0.2], [0.15, 0.2], first_seed=31)
    for u_init in (None, ui):
        out = S.solve_batch(w["x0"], xr, u_init=u_init, **kw)
        ref = O.solve_batch(po, w["x0"], xr, u_init=None if u_init is None else u_init.reshape(B, -1), **kw)
        _assert_parity(out, ref, need_frac=1.0)
        assert (out["iters"] == ref["iters"]).mean() >= 0.99
    # ragged: fewer problems than lanes of one warp, and a batch that leaves lanes idle at the end
    for nb in (1, 33):
        out = S.solve_batch(w["x0"][:nb], xr[:nb], **{k: v[:nb] for k, v in kw.items()})
        ref = O.solve_batch(po, w["x0"][:nb], xr[:nb], **{k: v[:nb] for k, v in kw.items()})
        _assert_parity(out, ref, need_frac=1.0)
    S.close()


@pytest.mark.parametrize("N", [1, 2, 3, 33, 127])
def test_extreme_horizons_both_kernels(env, N):
    """Shortest horizons (one to three stages), one stage more than a warp, and the longest supported one."""
    O, shim = env["O"], env["shim"]
    w = env["synth"].robots_on_map(B=64 if N < 100 else 16, seed=9)
    p = env["make"]("B", env["y"], N=N)
    po = O.variant_params("B", env["y"], N=N)
    S = shim.Solver(p)
    ref = O.solve_batch(po, w["x0"], w["goal"])
    for kind in (shim.KERNEL_WARP, shim.KERNEL_LANE):
        S.set_kernel(kind)
        out = S.solve_batch(w["x0"], w["goal"])
        assert S.last_kernel_kind == kind
        _assert_parity(out, ref, need_frac=0.9)
        assert out["X"].shape == (w["x0"].shape[0], N + 1, 3)
    S.close()


def test_lane_kernel_instances(env, robots, monkeypatch):
    """The lane kernels exist once per problem family (RK4 + goal, Euler + trajectory) and once with run-time switches.
    The generic instance must agree with the specialised ones, and it alone serves the mixed families (here the
    tracking cost on RK4 dynamics, and the goal cost on Euler dynamics)."""
    O, shim, synth = env["O"], env["shim"], env["synth"]
    y, w = env["y"], robots
    N = y["N"]
    pxf, puf = synth.straight_reference(w["x0"], w["goal"], N)
    for variant, over in (("B", {}), ("C", {}), ("C", dict(integrator=shim.RK4)), ("B", dict(integrator=shim.EULER))):
        xr, kw = (w["goal"], {}) if variant == "B" else (pxf, dict(uref=puf))
        po = O.variant_params(variant, y, **over)
        ref = O.solve_batch(po, w["x0"], xr, **kw)
        outs = []
        for generic in (False, True):
            if generic:
                monkeypatch.setenv("B200MPC_LANE_GENERIC", "1")
            S = shim.Solver(env["make"](variant, y, **over))
            monkeypatch.delenv("B200MPC_LANE_GENERIC", raising=False)
            S.set_kernel(shim.KERNEL_LANE)
            out = S.solve_batch(w["x0"], xr, **kw)
            S.close()
            _assert_parity(out, ref, need_frac=0.95)
            outs.append(out)
        assert np.array_equal(outs[0]["status"], outs[1]["status"]) and np.array_equal(outs[0]["iters"], outs[1]["iters"])
        assert np.max(np.abs(outs[0]["U"] - outs[1]["U"])) <= 1e-9


def test_closed_loop_with_obstacles_variant_a(env):
    """Config 2 with the obstacle cost active (variant A): every control step runs scan (ray-cast of map_carto) ->
    obstacle list (GPU, first 160 cells) -> look-ahead goal (GPU) -> solve (GPU) -> plant.  Each step's solve is checked
    against the oracle on identical inputs, and the robot must keep its distance from the occupied cells.  (Whether it
    approaches the goal is the reference's business: variant A weighs the x error with 5e-5 against obstacle terms of
    order one, and the oracle steers exactly the same way.)"""
    O, synth = env["O"], env["synth"]
    from ros2_mpc_b200 import MpcPointStabilization, obstacles as ob, references as rf
    y = env["y"]
    m = synth.load_map()
    clr = synth.clearance(m)
    start = np.array([-2.965, 2.315, 0.3])
    goal5 = np.array([-1.9, 2.9, 0.0, 0.0, 0.5])
    path = np.linspace(start[:2], goal5[:2], 40)
    head, _, _ = rf.get_headings(path, y["dt"])
    mpc = MpcPointStabilization()
    po = O.variant_params("A", y)
    N = mpc.N
    x = start.copy()
    d0 = np.linalg.norm(x[:2] - goal5[:2])
    worst_u, min_clear, n_cmp = 0.0, 1e9, 0
    for step in range(45):
        scan, angles = synth.raycast(m, x[None, :2], x[None, 2:3].ravel())
        ox, oy, cnt = ob.get_obstacles_batch_gpu(scan, angles, y["costmap_size"], y["resolution"], x[None, :2], x[None, 2], 160)
        rx, ry, rc = ob.get_obstacles(scan, angles, y["costmap_size"], y["resolution"], x[None, :2], x[None, 2], 160)
        assert np.array_equal(cnt, rc) and np.max(np.abs(ox - rx)) <= 1e-12 and np.max(np.abs(oy - ry)) <= 1e-12
        gm = rf.get_goal_for_mpc(path, head.reshape(-1, 1), goal5, x, y["look_ahead_distance"])
        x0 = np.array([x[0], x[1], x[2] % (2 * np.pi)])
        x_opt, u_opt = mpc.perform_mpc(np.zeros((2, N)), x0, gm, ox[0], oy[0])
        ro = O.solve(po, x0, gm, obs_x=ox[0], obs_y=oy[0])
        if ro["status"] == 0 and abs(ro["cost"] - mpc.last_cost) <= 1e-5 * abs(ro["cost"]):  # same local optimum
            worst_u = max(worst_u, float(np.max(np.abs(u_opt - ro["U"]))))
            n_cmp += 1
        u = u_opt[:, 0]
        th = x[2]
        tm, te = th + 0.5 * y["dt"] * u[1], th + y["dt"] * u[1]
        x[0] += y["dt"] / 6 * u[0] * (np.cos(th) + 4 * np.cos(tm) + np.cos(te))
        x[1] += y["dt"] / 6 * u[0] * (np.sin(th) + 4 * np.sin(tm) + np.sin(te))
        x[2] = te
        r, c = synth.world_to_cell(m, x[:2])
        min_clear = min(min_clear, float(clr[r, c]))
    mpc.close()
    assert n_cmp >= 40 and worst_u <= U_ATOL
    assert np.isfinite(x).all() and np.linalg.norm(x[:2] - start[:2]) > 0.1 and d0 > 0
    assert min_clear > 0.2


def test_fleet_with_obstacle_cost_runs_on_the_device(env):
    """Variant-A fleet on map_carto, every control step entirely on the device (fleet.FleetObstacleAvoidance): ray-cast of the
    shared map from the true pose -> obstacle list at the measured pose -> look-ahead goal -> solve with the obstacle cost
    -> limiter / goal logic / plant / next measurement.  Checked against the same loop on the host, stage by stage through
    the host-buffer entry points, with the node's expressions written out in numpy: scans, obstacle lists, commands, flags
    and measurements must agree exactly.  (Whether a robot keeps clear of the walls is the reference's business: the list
    holds the first 160 cells of the scan, a failed solve commands (0, 0), and the limiter adds 0.03 to both components.)"""
    from ros2_mpc_b200 import obstacles as ob, references as rf, sensors
    from ros2_mpc_b200.fleet import FleetObstacleAvoidance
    y, shim, synth = env["y"], env["shim"], env["synth"]
    m = synth.load_map()
    B, T, K = 48, 25, 24
    w = synth.robots_on_map(B=B, seed=7, m=m)
    start = w["x0"].copy()
    tt = np.linspace(0, 1, K)[None, :, None]
    path = start[:, None, :2] * (1 - tt) + w["goal"][:, None, :2] * tt           # one straight path per robot
    head = np.stack([rf.get_headings(path[b], y["dt"])[0] for b in range(B)])
    goal = np.c_[w["goal"][:, :2], np.zeros((B, 2)), w["goal"][:, 2]]
    fleet = FleetObstacleAvoidance(start, goal, path, head, m, params=y)
    fleet.step(T)
    dev = fleet.snapshot()
    fleet.close()

    S = shim.Solver(env["make"]("A", y))
    bits = sensors.map_bits(m)
    H, W = m["occ"].shape
    angles = np.array([0.0, 6.28])
    state = start.copy()
    x0 = np.round(state, 2); x0[:, 2] = x0[:, 2] % (2 * np.pi)
    u_last = np.zeros((B, 2)); flag = np.zeros(B, dtype=bool); cmd = np.zeros((B, 2))
    hit = False
    for step in range(T):
        scan = S.raycast_batch(bits, H, W, m["origin"], m["resolution"], state)
        ox, oy, cnt = ob.get_obstacles_batch_gpu(scan, angles, y["costmap_size"], y["resolution"], x0[:, :2], x0[:, 2], 160, solver=S)
        goal_mpc, _ = rf.get_goals_batch(path, head, goal, x0[:, :2], y["look_ahead_distance"], solver=S)
        out = S.solve_batch(x0, goal_mpc, obs_x=ox, obs_y=oy)
        ok = np.isin(out["status"], (0, 1))
        for b in range(B):
            u = out["U"][b, 0] if ok[b] else np.zeros(2)
            if flag[b]:
                cmd[b] = 0.0
            elif np.linalg.norm(u - u_last[b]) > 0.03:
                cmd[b] = u_last[b] + 0.03
                u_last[b] = u
            else:
                cmd[b] = u
                u_last[b] = u
            if np.linalg.norm(x0[b, 0:2] - goal[b, 0:2]) > y["goal_threshold"]:
                flag[b] = False
            elif not flag[b]:
                cmd[b] = 0.0
                flag[b] = True
        th, v, w_ = state[:, 2].copy(), cmd[:, 0], cmd[:, 1]
        tm, te = th + 0.5 * y["dt"] * w_, th + y["dt"] * w_
        state[:, 0] += y["dt"] / 6.0 * v * (np.cos(th) + 4 * np.cos(tm) + np.cos(te))
        state[:, 1] += y["dt"] / 6.0 * v * (np.sin(th) + 4 * np.sin(tm) + np.sin(te))
        state[:, 2] = te
        x0 = np.round(state, 2); x0[:, 2] = x0[:, 2] % (2 * np.pi)
        for b in range(B):
            r, c = synth.world_to_cell(m, state[b, :2])
            hit = hit or bool(m["occ"][r, c])
    S.close()
    # the last step's sensor data and obstacle lists, then everything the loop carries
    assert np.array_equal(dev["scan"], scan) and np.array_equal(dev["obs_count"], cnt)
    assert np.array_equal(dev["obs_x"], ox) and np.array_equal(dev["obs_y"], oy)
    assert np.array_equal(dev["status"], out["status"]) and np.array_equal(dev["U"], out["U"])
    assert np.array_equal(dev["goal_flag"].astype(bool), flag)
    assert np.array_equal(dev["cmd"], cmd) and np.array_equal(dev["u_last"], u_last)
    assert np.array_equal(dev["x0"], x0)
    assert np.max(np.abs(dev["state"] - state)) <= 1e-9
    # (a failed solve commands (0, 0), as the node would raise; measured: 39 of 48 solves of the last step succeed — in
    # closed loop the robots sit where the exp(c/s) terms are large, the hard end of config 3)
    print(f"variant-A fleet: last step {np.isin(dev['status'], (0, 1)).mean():.2f} converged, entered an occupied cell: {hit}, "
          f"largest displacement {np.linalg.norm(state[:, :2] - start[:, :2], axis=1).max():.2f} m")
    assert np.isin(dev["status"], (0, 1)).mean() >= 0.7
    assert np.linalg.norm(state[:, :2] - start[:, :2], axis=1).max() > 0.1   # the fleet moves


@pytest.mark.parametrize("variant", ["A", "B"])
def test_iteration_bounded_warp_kernel_launches(env, variant, monkeypatch):
    """The warp kernel can export a problem's solver state as the lane kernel does and resume it in a later launch
    (B200MPC_WARP_CAPS: a cascade of iteration-bounded launches; off by default, it does not pay).  With bounds of 5 and 12
    iterations nearly every problem crosses two launch boundaries; results must be bit-identical to the single launch."""
    shim, synth = env["shim"], env["synth"]
    w = synth.robots_on_map(B=4096, seed=9)
    xr, kw = _inputs(env, variant, w)
    S = shim.Solver(env["make"](variant, env["y"]))
    S.set_kernel(shim.KERNEL_WARP)
    one = S.solve_batch(w["x0"], xr, **kw)
    S.close()
    monkeypatch.setenv("B200MPC_WARP_CAPS", "5,12")
    S = shim.Solver(env["make"](variant, env["y"]))
    monkeypatch.delenv("B200MPC_WARP_CAPS")
    S.set_kernel(shim.KERNEL_WARP)
    n0 = S.launch_count
    out = S.solve_batch(w["x0"], xr, **kw)
    assert S.launch_count - n0 == 3
    S.close()
    for k in ("status", "iters", "ls", "cost", "X", "U"):
        assert np.array_equal(out[k], one[k], equal_nan=True), k


def test_restoration_second_stage_on_both_kernels(env):
    """The problems of tests/golden/resto_golden.npz (first restoration stage finds nothing: the roll-out runs into an obstacle
    point) on the warp kernel and on the lane kernel (warp-cooperative restoration): status, optimum and iteration count of
    the oracle, and the cost scipy's SLSQP reaches where it converges."""
    O, shim = env["O"], env["shim"]
    g = np.load(os.path.join(G, "resto_golden.npz"))
    po = O.variant_params("A", env["y"])
    kw = dict(obs_x=g["obs_x"], obs_y=g["obs_y"])
    ref = O.solve_batch(po, g["x0"], g["goal"], **kw)
    assert (ref["status"] == 0).all()
    for kind in (shim.KERNEL_WARP, shim.KERNEL_LANE):
        S = shim.Solver(env["make"]("A", env["y"]))
        S.set_kernel(kind)
        out = S.solve_batch(g["x0"], g["goal"], **kw)
        S.close()
        _assert_parity(out, ref, need_frac=1.0)
        assert np.array_equal(out["iters"], ref["iters"])
        ok = np.isfinite(g["slsqp_cost"])
        assert np.all(np.abs(out["cost"][ok] - g["slsqp_cost"][ok]) <= COST_RTOL * np.abs(g["slsqp_cost"][ok]))


def test_handles_with_different_shared_memory_needs_coexist(env, robots):
    """The warp kernel's dynamic shared-memory limit is a per-kernel attribute shared by all handles of the process: a
    handle created later with a smaller need (no obstacle lists) must not break the launches of an earlier one."""
    shim, w = env["shim"], robots
    Sa = shim.Solver(env["make"]("A", env["y"]))          # obstacle lists in shared memory: > 48 KB
    Sb = shim.Solver(env["make"]("B", env["y"]))          # created afterwards, needs less
    Sn = shim.Solver(env["make"]("B", env["y"], N=5))     # and a tiny one
    k = 32
    for _ in range(2):
        a = Sa.solve_batch(w["x0"][:k], w["goal"][:k], obs_x=w["obs_x"][:k], obs_y=w["obs_y"][:k])
        b = Sb.solve_batch(w["x0"][:k], w["goal"][:k])
        n = Sn.solve_batch(w["x0"][:k], w["goal"][:k])
        assert np.isin(b["status"], (0, 1)).all() and np.isin(n["status"], (0, 1)).all() and a["X"].shape == (k, 31, 3)
    for S in (Sa, Sb, Sn):
        S.close()


def test_handle_lifecycle_releases_device_memory(env, robots):
    """Creating, using (lane kernel: ~0.9 GB workspace; streamed and zero-copy host paths) and destroying handles
    repeatedly must give the device memory back."""
    import torch
    shim, w = env["shim"], robots
    x0, goal = np.tile(w["x0"], (96, 1)), np.tile(w["goal"], (96, 1))   # 36 864 problems -> lane kernel

    def cycle():
        S = shim.Solver(env["make"]("B", env["y"]))
        o = S.solve_batch(x0, goal)
        assert S.last_kernel_kind == shim.KERNEL_LANE and np.isin(o["status"], (0, 1)).all()
        S.solve_batch(w["x0"][:1], w["goal"][:1])
        S.obstacles_batch(w["scan"][:8], *__import__("ros2_mpc_b200.obstacles", fromlist=["beam_table"]).beam_table(360, w["angles"]),
                          w["x0"][:8, :2], w["x0"][:8, 2], 2.0, 0.05, 160)
        S.close()

    cycle()
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(6):
        cycle()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 64 * 1024 * 1024, (free0, free1)


def test_c_abi_argument_errors(env):
    """Bad arguments come back as negative codes with a message — never a crash, never an exception across the ABI."""
    import ctypes as C
    shim = env["shim"]
    L = shim.lib()
    p = env["make"]("B", env["y"])
    for field, val in (("N", 0), ("N", 500), ("dt", 0.0)):
        q = env["make"]("B", env["y"])
        setattr(q, field, val)
        assert not L.b200mpc_create(C.byref(q), 0)
        assert len(L.b200mpc_last_error(None)) > 0
    assert not L.b200mpc_create(C.byref(p), 99) and b"device" in L.b200mpc_last_error(None)
    S = shim.Solver(p)
    h = S._h
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    x = np.zeros((4, 3)); X = np.zeros((4, 31, 3)); U = np.zeros((4, 30, 2)); st = np.zeros(4, np.int32)
    P = lambda a: a.ctypes.data_as(dp)
    E_ARG = -1
    assert L.b200mpc_solve_batch(h, -1, P(x), P(x), None, None, None, 0, None, P(X), P(U), None, st.ctypes.data_as(ip), None, None) == E_ARG
    assert L.b200mpc_solve_batch(h, 4, None, P(x), None, None, None, 0, None, P(X), P(U), None, st.ctypes.data_as(ip), None, None) == E_ARG
    assert L.b200mpc_solve_batch(h, 4, P(x), P(x), None, None, None, 0, None, None, P(U), None, st.ctypes.data_as(ip), None, None) == E_ARG
    assert b"NULL" in L.b200mpc_last_error(h)
    assert L.b200mpc_solve_batch(None, 4, P(x), P(x), None, None, None, 0, None, P(X), P(U), None, st.ctypes.data_as(ip), None, None) == E_ARG
    assert L.b200mpc_set_kernel(h, 7) == E_ARG
    sc = np.ones((2, 8)); t = np.ones(8); ox = np.zeros((2, 160)); cnt = np.zeros(2, np.int32)
    assert L.b200mpc_obstacles_batch(h, 2, 8, P(sc), P(t), P(t), P(x), P(x), 2.0, 0.05, 0, P(ox), P(ox), cnt.ctypes.data_as(ip)) == E_ARG
    assert L.b200mpc_obstacles_batch(h, 2, 8, P(sc), P(t), P(t), P(x), P(x), 2.0, 0.0, 160, P(ox), P(ox), cnt.ctypes.data_as(ip)) == E_ARG
    assert L.b200mpc_obstacles_batch(h, 2, 8, P(sc), P(t), P(t), P(x), P(x), 500.0, 0.05, 160, P(ox), P(ox), cnt.ctypes.data_as(ip)) == E_ARG
    assert L.b200mpc_goals_batch(h, 2, 0, P(x), P(x), 0, P(x), P(x), 0.5, P(x), None) == E_ARG
    assert L.b200mpc_reftraj_batch(h, 2, 4, P(x), P(x), P(x), P(x), 9, 0, P(x), P(x), P(X), P(U), None) == E_ARG
    # the handle is still usable afterwards
    o = S.solve_batch(np.zeros((1, 3)), np.array([[0.5, 0.2, 0.0]]))
    assert o["status"][0] == 0
    S.close()


def test_scan_recursion_matches_serial_recursion(env, robots, monkeypatch):
    """Small batches run the warp kernel with the parallel-in-time (scan) form of the Riccati recursion, batches that
    fill the machine with the serial lane-parallel form (B200MPC_KKT_SCAN=0/1 forces one).  Both against the oracle
    and against each other: same status, same optimum; the scan's iterates differ at the 1e-9 level, so iteration
    counts may differ on a few problems."""
    O, shim = env["O"], env["shim"]
    w = robots
    po = O.variant_params("B", env["y"])
    ref = O.solve_batch(po, w["x0"], w["goal"])
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("B200MPC_KKT_SCAN", mode)
        S = shim.Solver(env["make"]("B", env["y"]))
        monkeypatch.delenv("B200MPC_KKT_SCAN")
        S.set_kernel(shim.KERNEL_WARP)
        outs[mode] = S.solve_batch(w["x0"], w["goal"])
        S.close()
        _assert_parity(outs[mode], ref, need_frac=1.0)
    assert (outs["0"]["iters"] == ref["iters"]).all()
    assert (outs["1"]["iters"] == ref["iters"]).mean() >= 0.99
    assert np.max(np.abs(outs["0"]["U"] - outs["1"]["U"])) <= 1e-6
    # horizons the scan does not cover (more than 32 stages) fall back to the serial form
    monkeypatch.setenv("B200MPC_KKT_SCAN", "1")
    S = shim.Solver(env["make"]("B", env["y"], N=50))
    monkeypatch.delenv("B200MPC_KKT_SCAN")
    o = S.solve_batch(w["x0"][:16], w["goal"][:16])
    r = O.solve_batch(O.variant_params("B", env["y"], N=50), w["x0"][:16], w["goal"][:16])
    _assert_parity(o, r, need_frac=1.0)
    S.close()


# ---- multi-GPU product entry point --------------------------------------------------------------------
def test_multi_gpu_sharded_solve_matches_single_gpu(env, robots):
    """b200mpc_solve_batch_multi / Mpc(devices=[...]).perform_mpc_batch: one batch sharded over the GPUs of the node (one
    handle + host thread per device, results written straight into one set of host arrays) is bit-identical to the same
    batch solved on one GPU — small batches (warp kernel), a streamed large batch (lane kernel) and the obstacle variant."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (gpurun --gpus 2)")
    shim, synth, y = env["shim"], env["synth"], env["y"]
    devs = list(range(min(n, 4)))
    for variant, B in (("B", 383), ("A", 96), ("C", 200)):
        w = {k: v[:B] for k, v in robots.items() if isinstance(v, np.ndarray) and v.shape[0] == 384}
        xr, kw = _inputs(env, variant, w)
        p = env["make"](variant, y)
        S1, SM = shim.Solver(p, device=0), shim.MultiSolver(p, devs)
        a, b = S1.solve_batch(w["x0"], xr, **kw), SM.solve_batch(w["x0"], xr, **kw)
        for k in ("X", "U", "cost", "status", "iters", "ls"):
            assert np.array_equal(a[k], b[k]), (variant, k)
        S1.close(); SM.close()
    # a large batch through page-locked buffers: every shard is streamed by its own device
    R, Sd, N = 4096, 80, y["N"]
    wm = synth.robots_on_map(B=R, seed=0)
    p = env["make"]("B", y)
    ui = synth.warm_start_seeds(Sd, N, list(p.u_lo), list(p.u_hi))
    B = R * Sd
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    x0, goal = pin(np.tile(wm["x0"], (Sd, 1))), pin(np.tile(wm["goal"], (Sd, 1)))
    u_init = pin(np.repeat(ui, R, axis=0).reshape(B, N, 2))

    def outs():
        return dict(X=pin(np.empty((B, N + 1, 3))), U=pin(np.empty((B, N, 2))), cost=pin(np.empty(B)),
                    status=pin(np.empty(B, np.int32)), iters=pin(np.empty(B, np.int32)), ls=pin(np.empty(B, np.int32)))

    S1, SM = shim.Solver(p, device=0), shim.MultiSolver(p, devs[:2])
    a = S1.solve_batch(x0, goal, u_init=u_init, out=outs())
    b = SM.solve_batch(x0, goal, u_init=u_init, out=outs())
    assert all(s.last_solve_chunks > 0 for s in SM._solvers)          # both shards (163 840 problems each) were streamed
    for k in ("X", "U", "cost", "status", "iters", "ls"):
        assert np.array_equal(a[k], b[k]), k
    S1.close(); SM.close()
    # the drop-in class
    from ros2_mpc_b200 import MpcPointStabilizationLocal
    m1, m2 = MpcPointStabilizationLocal(), MpcPointStabilizationLocal(devices=devs[:2])
    r1 = m1.perform_mpc_batch(None, robots["x0"][:101], robots["goal"][:101])
    r2 = m2.perform_mpc_batch(None, robots["x0"][:101], robots["goal"][:101])
    assert np.array_equal(r1["u_opt"], r2["u_opt"]) and np.array_equal(r1["status"], r2["status"])
    m1.close(); m2.close()
