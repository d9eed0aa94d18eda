"""CPU tests of the oracle (oracle/mpc_oracle.c): derivatives, linear algebra, known answers, scipy cross-check.

The reference has no tests or golden vectors, so these are the pins the repo creates
(SURVEY.md section 8c): finite differences for every analytic derivative, agreement of the structure-exploiting
Riccati backend with an independent dense LDL^T of the full KKT matrix, analytic known answers, and agreement of
the optimum with scipy's SLSQP on the same multiple-shooting NLP."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as O


def _random_point(p, rng, variant):
    N = p.N
    x0 = np.array([0.3, -0.2, 0.7])
    X = rng.normal(0, 0.5, (N + 1, 3)); X[0] = x0
    U = rng.uniform(-0.2, 0.2, (N, 2))
    lam = rng.normal(0, 1, (N, 3))
    kw = {}
    xref = np.array([1.0, 0.5, 0.3])
    if variant == "C":
        xref = rng.normal(0, 1, 3 * N); kw["uref"] = rng.uniform(-0.1, 0.1, 2 * N)
    if p.obs_form != O.OBS_NONE:
        ang = rng.uniform(0, 2 * np.pi, p.M); rr = rng.uniform(2.0, 4.0, p.M)
        kw["obs_x"] = rr * np.cos(ang); kw["obs_y"] = rr * np.sin(ang)
    return x0, xref, X, U, lam, kw


def _unpack(z, N, x0):
    X = np.vstack([x0, z[:3 * N].reshape(N, 3)])
    U = z[3 * N:].reshape(N, 2)
    return X, U


@pytest.mark.parametrize("variant,obstacles", [("A", None), ("B", None), ("B", True), ("C", None)])
def test_derivatives_match_finite_differences(variant, obstacles):
    rng = np.random.default_rng(3)
    p = O.variant_params(variant, N=6, obstacles=obstacles)
    N = p.N
    x0, xref, X, U, lam, kw = _random_point(p, rng, variant)
    base = O.evaluate(p, x0, xref, X, U, lam=lam, **kw)
    z0 = np.concatenate([X[1:].ravel(), U.ravel()])

    def f_of(z):
        Xz, Uz = _unpack(z, N, x0)
        return O.evaluate(p, x0, xref, Xz, Uz, **kw)["f"]

    def c_of(z):
        Xz, Uz = _unpack(z, N, x0)
        return O.evaluate(p, x0, xref, Xz, Uz, **kw)["c"].ravel()

    def lag_grad(z):
        Xz, Uz = _unpack(z, N, x0)
        e = O.evaluate(p, x0, xref, Xz, Uz, lam=lam, **kw)
        return e["grad"] + jac_from_stages(e["stages"], N, p.dt).T @ lam.ravel()

    def jac_from_stages(st, N, dt):
        J = np.zeros((3 * N, 5 * N))
        for k in range(N):
            a13, a23, b11, b12, b21, b22 = st[k, :6]
            A = np.array([[1, 0, a13], [0, 1, a23], [0, 0, 1]])
            Bm = np.array([[b11, b12], [b21, b22], [0, dt]])
            J[3 * k:3 * k + 3, 3 * k:3 * k + 3] = np.eye(3)          # d c_{k+1} / d X_{k+1}
            if k >= 1:
                J[3 * k:3 * k + 3, 3 * (k - 1):3 * (k - 1) + 3] = -A  # d c_{k+1} / d X_k
            J[3 * k:3 * k + 3, 3 * N + 2 * k:3 * N + 2 * k + 2] = -Bm
        return J

    h = 1e-6
    n = len(z0)
    g_fd = np.zeros(n); J_fd = np.zeros((3 * N, n)); H_fd = np.zeros((n, n))
    for i in range(n):
        e = np.zeros(n); e[i] = h
        g_fd[i] = (f_of(z0 + e) - f_of(z0 - e)) / (2 * h)
        J_fd[:, i] = (c_of(z0 + e) - c_of(z0 - e)) / (2 * h)
        H_fd[:, i] = (lag_grad(z0 + e) - lag_grad(z0 - e)) / (2 * h)
    assert np.allclose(base["grad"], g_fd, rtol=1e-6, atol=1e-7)
    assert np.allclose(jac_from_stages(base["stages"], N, p.dt), J_fd, rtol=1e-6, atol=1e-8)
    # Lagrangian Hessian assembled from the stage blocks
    H = np.zeros((n, n))
    for k in range(N + 1):
        Hk = base["stages"][k, 6:31].reshape(5, 5)
        idx = [3 * (k - 1) + i if k >= 1 else -1 for i in range(3)] + [3 * N + 2 * k + i if k < N else -1 for i in range(2)]
        for i in range(5):
            for j in range(5):
                if idx[i] >= 0 and idx[j] >= 0:
                    H[idx[i], idx[j]] += Hk[i, j]
    assert np.allclose(H, H_fd, rtol=2e-5, atol=2e-6)
    assert np.allclose(H, H.T)


def test_closed_form_rk4_equals_staged_rk4():
    # the derivative code uses F = (x + dt v C/6, ...); the value code stages k1..k4 like the reference
    p = O.variant_params("B", N=1)
    rng = np.random.default_rng(0)
    for _ in range(50):
        x = rng.normal(0, 2, 3); u = rng.uniform(-1, 1, 2)
        X = np.vstack([x, np.zeros(3)]); U = u[None]
        c = O.evaluate(p, x, np.zeros(3), X, U)["c"][0]
        dt = p.dt; th, v, w = x[2], u[0], u[1]
        tm, te = th + dt * w / 2, th + dt * w
        F = np.array([x[0] + dt * v / 6 * (np.cos(th) + 4 * np.cos(tm) + np.cos(te)),
                      x[1] + dt * v / 6 * (np.sin(th) + 4 * np.sin(tm) + np.sin(te)), th + dt * w])
        assert np.allclose(-c, F, rtol=0, atol=1e-14)


def test_ldl_solve_and_inertia_against_numpy():
    rng = np.random.default_rng(5)
    L = O.lib()
    for n, m in ((12, 5), (40, 17), (7, 0)):
        Hm = rng.normal(size=(n, n)); Hm = Hm @ Hm.T + 0.1 * np.eye(n)
        if rng.random() < 0.5 and n > 2:
            Hm[0, 0] -= 50.0  # make the (1,1) block indefinite
        J = rng.normal(size=(m, n))
        K = np.block([[Hm, J.T], [J, np.zeros((m, m))]]) if m else Hm
        b = rng.normal(size=n + m)
        A = np.ascontiguousarray(K.copy()); x = b.copy()
        inertia = (C.c_int * 3)()
        info = L.orc_ldl_solve(n + m, A.ctypes.data_as(C.POINTER(C.c_double)), x.ctypes.data_as(C.POINTER(C.c_double)), inertia)
        assert info == 0
        ev = np.linalg.eigvalsh(K)
        assert list(inertia) == [int((ev > 0).sum()), int((ev < 0).sum()), 0]
        assert np.allclose(K @ x, b, atol=1e-8 * max(1, np.abs(b).max()))


@pytest.mark.parametrize("variant", ["A", "B", "C"])
def test_riccati_and_dense_backends_agree(variant):
    rng = np.random.default_rng(11)
    N = 10
    pr = O.variant_params(variant, N=N, linear_solver=0)
    pdn = O.variant_params(variant, N=N, linear_solver=1)
    for _ in range(4):
        x0 = np.r_[rng.uniform(-1, 1, 2), rng.uniform(0, 6.28)]
        goal = x0 + np.r_[rng.uniform(-0.6, 0.6, 2), rng.uniform(-1, 1)]
        kw = {}; xr = goal
        if variant == "C":
            t = (np.arange(1, N + 1) / N)[:, None]
            xr = (x0 * (1 - t) + goal * t).ravel(); kw["uref"] = np.tile([0.1, 0.0], N)
        if variant == "A":
            t = np.linspace(-1, 1, pr.M)
            kw.update(obs_x=x0[0] + t, obs_y=x0[1] + 1.0 + 0 * t)
        a = O.solve(pr, x0, xr, **kw); b = O.solve(pdn, x0, xr, **kw)
        assert a["status"] == b["status"] == 0
        assert a["stats"]["iters"] == b["stats"]["iters"]
        assert np.allclose(a["X"], b["X"], atol=1e-9) and np.allclose(a["U"], b["U"], atol=1e-9)
        assert abs(a["cost"] - b["cost"]) <= 1e-9 * abs(a["cost"])


def test_known_answers_variant_b():
    """BASELINE.md section 4 anchors: J* = 63.0480930889 (scipy L-BFGS-B / SLSQP), u0* = (0.15, 0.2);
    X_N carries no cost, so omega_{N-1} = 0 and v_{N-1} = min(v_max, root of 2 R v = kappa e^{-kappa v})."""
    p = O.variant_params("B")
    r = O.solve(p, [0, 0, 0], [1, 1, 0])
    assert r["status"] == 0
    # IPOPT relaxes the bounds by 1e-8 (bound_relax_factor), which moves the optimum by ~1.4e-6 in cost
    assert abs(r["cost"] - 63.0480930889) <= 1e-5 * 63.05
    assert np.allclose(r["U"][:, 0], [0.15, 0.2], atol=1e-6)
    assert np.allclose(r["U"][:, -1], [0.15, 0.0], atol=1e-6)
    assert np.allclose(r["X"][:, -1], [0.74605, 0.44350, 0.81959], atol=2e-5)
    assert np.array_equal(r["X"][:, 0], [0, 0, 0])
    assert int(np.sum(np.abs(r["U"][0] - 0.15) < 1e-6)) == 30
    assert int(np.sum(np.abs(np.abs(r["U"][1]) - 0.2) < 1e-6)) == 17


def test_known_answer_variant_a_sentinel_obstacles():
    p = O.variant_params("A")
    ox = np.full(160, 100.0)
    r = O.solve(p, [0, 0, 0], [10, 10, 0], obs_x=ox, obs_y=ox)
    assert r["status"] == 0
    assert abs(r["cost"] - 5134.1828) <= 1e-5 * 5134.0
    assert np.allclose(r["U"][:, 0], [0.2, 0.1], atol=1e-6)
    assert r["cost"] > 31 * 160  # every obstacle term is >= 1


def test_invalid_number_when_an_obstacle_sits_on_the_start_guess():
    # the Opti start guess is X = 0: an obstacle point on the world origin overflows exp(c/s) (SURVEY.md hard parts)
    p = O.variant_params("A")
    ox = np.full(160, 100.0); oy = np.full(160, 100.0)
    ox[7] = 0.001; oy[7] = 0.0
    r = O.solve(p, [1.0, 1.0, 0.0], [2, 2, 0], obs_x=ox, obs_y=oy)
    assert r["status"] == -13


def _scipy_solve(p, x0, goal):
    from scipy.optimize import minimize
    N = p.N
    z0 = np.zeros(5 * N)

    def f(z):
        X, U = _unpack(z, N, x0)
        e = O.evaluate(p, x0, goal, X, U)
        return e["f"], e["grad"]

    def c(z):
        X, U = _unpack(z, N, x0)
        return O.evaluate(p, x0, goal, X, U)["c"].ravel()

    lo = np.r_[np.full(3 * N, -np.inf), np.tile([p.u_lo[0], p.u_lo[1]], N)]
    hi = np.r_[np.full(3 * N, np.inf), np.tile([p.u_hi[0], p.u_hi[1]], N)]
    res = minimize(f, z0, jac=True, method="SLSQP", bounds=list(zip(lo, hi)),
                   constraints=[{"type": "eq", "fun": c}], options={"maxiter": 500, "ftol": 1e-14})
    return res


def test_optimum_agrees_with_scipy_slsqp():
    """Independent third-party NLP solver on the same multiple-shooting problem (N=10 to keep it quick)."""
    p = O.variant_params("B", N=10)
    x0 = np.array([0.2, -0.1, 0.4]); goal = np.array([0.9, 0.5, 1.0])
    res = _scipy_solve(p, x0, goal)
    r = O.solve(p, x0, goal)
    assert r["status"] == 0
    assert abs(r["cost"] - res.fun) <= 1e-5 * abs(res.fun)
    X, U = _unpack(res.x, 10, x0)
    assert np.allclose(r["U"].T, U, atol=1e-4)
    assert np.allclose(r["X"].T, X, atol=1e-4)


def test_kkt_certificate_at_the_returned_point():
    """Stationarity / feasibility / complementarity recomputed from orc_eval, independent of the solver loop."""
    p = O.variant_params("B")
    x0 = np.array([0.0, 0.0, 0.0]); goal = np.array([1.0, 1.0, 0.0])
    r = O.solve(p, x0, goal)
    X, U = r["X"].T.copy(), r["U"].T.copy()
    N = p.N
    e = O.evaluate(p, x0, goal, X, U)
    assert np.abs(e["c"]).max() <= 1e-8
    # multipliers from the adjoint recursion lam_N = -g_N, lam_k = A_k' lam_{k+1} - g_k
    g = e["grad"]; st = e["stages"]
    lam = np.zeros((N + 2, 3))
    for k in range(N, 0, -1):
        gk = g[3 * (k - 1):3 * (k - 1) + 3]
        if k == N:
            lam[k] = -gk
        else:
            a13, a23 = st[k, 0], st[k, 1]
            A = np.array([[1, 0, a13], [0, 1, a23], [0, 0, 1]])
            lam[k] = A.T @ lam[k + 1] - gk
    # reduced gradient w.r.t. U_k: g_u - B_k' lam_{k+1}; must vanish off the bounds and have the right sign on them
    for k in range(N):
        b11, b12, b21, b22 = st[k, 2:6]
        Bm = np.array([[b11, b12], [b21, b22], [0, p.dt]])
        rg = g[3 * N + 2 * k:3 * N + 2 * k + 2] - Bm.T @ lam[k + 1]
        for i in range(2):
            at_lo = U[k, i] <= p.u_lo[i] + 1e-6
            at_hi = U[k, i] >= p.u_hi[i] - 1e-6
            if at_lo:
                assert rg[i] >= -1e-6
            elif at_hi:
                assert rg[i] <= 1e-6
            else:
                assert abs(rg[i]) <= 1e-6


def test_warm_start_reaches_the_same_optimum():
    p = O.variant_params("B")
    x0 = np.array([0.1, 0.2, 0.3]); goal = np.array([0.8, 0.7, 1.0])
    cold = O.solve(p, x0, goal)
    rng = np.random.default_rng(1)
    u = np.clip(rng.normal(0, 0.05, (2, 30)), np.array([[-0.05], [-0.2]]), np.array([[0.15], [0.2]]))
    warm = O.solve(p, x0, goal, u_init=u)
    assert cold["status"] == warm["status"] == 0
    assert abs(cold["cost"] - warm["cost"]) <= 1e-7 * cold["cost"]
    assert np.allclose(cold["U"], warm["U"], atol=1e-5)


def test_batch_matches_single_and_threads():
    p = O.variant_params("B", N=12)
    rng = np.random.default_rng(2)
    B = 24
    x0 = np.c_[rng.uniform(-1, 1, (B, 2)), rng.uniform(0, 6.28, B)]
    goal = x0 + np.c_[rng.uniform(-0.5, 0.5, (B, 2)), rng.uniform(-1, 1, B)]
    r1 = O.solve_batch(p, x0, goal, nthreads=1)
    r3 = O.solve_batch(p, x0, goal, nthreads=3)
    assert np.array_equal(r1["X"], r3["X"]) and np.array_equal(r1["status"], r3["status"])
    s = O.solve(p, x0[5], goal[5])
    assert np.array_equal(s["X"].T, r1["X"][5]) and s["cost"] == r1["cost"][5]


def test_restoration_second_stage_reaches_the_third_party_optimum():
    """Four config-3 problems of the obstacle-active variant on which the roll-out of the slacks runs into an obstacle point
    (tests/golden/make_resto_golden.py): with the first restoration stage alone the solve ended Restoration_Failed at a
    cost of 2.5e8 ... 4.6e24; the second stage (the plan shrunk towards standing still, lowest barrier objective) lets all
    four converge, and where scipy's SLSQP reaches a KKT point from the same cold start it is the same optimum."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resto_golden.npz"))
    p = O.variant_params("A")
    for n in range(len(g["index"])):
        kw = dict(obs_x=g["obs_x"][n], obs_y=g["obs_y"][n])
        r = O.solve(p, g["x0"][n], g["goal"][n], **kw)
        assert r["status"] == 0 and r["stats"]["n_resto"] >= 1, (int(g["index"][n]), r["status"], r["stats"])
        c = O.kkt_certificate(p, g["x0"][n], g["goal"][n], r["X"].T.copy(), r["U"].T.copy(), **kw)
        assert c["defect"] <= 1e-8 and c["complementarity"] <= 1e-6, c
        if np.isfinite(g["slsqp_cost"][n]):
            assert abs(r["cost"] - g["slsqp_cost"][n]) <= 1e-5 * abs(g["slsqp_cost"][n])
    assert np.isfinite(g["slsqp_cost"]).sum() >= 2
