"""CPU tests that pin the oracle (oracle/nlo_oracle.cc) before anything is compared with it.

Pins, all from the reference's own logs / fixtures:
  * PnP known answer, exact:  results/reproj_amd64.txt:5,10  (COST 2.33228e-11, iter 6, pose)
  * NDT 6-DoF fixture, +-2 % band (depends on Eigen eigenvector sign / unordered_map order):
    results/maha_amd64_simple.txt:10-14,24
  * fixture sizes: 954 605 room points, 96 NDT cells at 1.0 m, 630 PnP points
plus self-consistency the reference never checks: finite-difference Jacobians, loss derivatives,
thread-split semantics, long-double accumulation.
"""
import os

import numpy as np
import pytest

from nonlinear_optimizer_for_slam_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _pose0(oracle):
    return oracle.pose_from_Rt(np.eye(3), np.zeros(3))


def test_pnp_known_answer(oracle):
    X = oracle.pnp_reference_points()
    assert X.shape == (630, 3)
    Xs, px, K = syn.pnp_fixture()
    np.testing.assert_array_equal(X, Xs)
    pose, it, cost, trace = oracle.reproj_solve(X, px, K, _pose0(oracle), oracle.LOSS_EXPONENTIAL,
                                                [1.0, 1.0])
    assert it == 6                                   # "iter: 6"
    assert "%.5e" % cost == "2.33228e-11"            # "COST: 2.33228e-11"
    R, t = oracle.pose_to_Rt(pose)
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
    Ti = np.linalg.inv(T)                            # the test prints Solve(...).inverse()
    q = oracle.rotmat_to_quat(Ti[:3, :3])
    assert ["%.6g" % v for v in Ti[:3, 3]] == ["-0.1", "0.123", "-0.5"]
    assert "%.6g" % q[0] == "-2.38636e-09" and "%.6g" % q[1] == "5.42421e-11"
    assert "%.6g" % q[2] == "0.0499792" and "%.6g" % q[3] == "0.99875"


def test_golden_vectors_on_disk(oracle):
    """tests/golden/*.npz were written by tests/golden/make_golden.py from this oracle at the
    commit that pinned it; a later edit of the oracle must not move them."""
    g = np.load(os.path.join(GOLDEN, "pnp_fixture_trace.npz"))
    X, px, K = syn.pnp_fixture()
    pose, it, cost, trace = oracle.reproj_solve(X, px, K, _pose0(oracle), 1, [1.0, 1.0])
    assert it == int(g["iterations"])
    np.testing.assert_allclose(trace, g["trace"], rtol=1e-12, atol=1e-18)
    np.testing.assert_allclose(pose, g["pose"], rtol=0, atol=1e-14)
    g = np.load(os.path.join(GOLDEN, "ndt_small_sums.npz"))
    H, gr, c = oracle.ndt6_assemble(g["point"], g["mean"], g["sqrt_info"], g["R"], g["t"], 1, [1.0, 1.0])
    np.testing.assert_allclose(H, g["H21"], rtol=1e-12)
    np.testing.assert_allclose(gr, g["g6"], rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(c, g["cost"], rtol=1e-12)
    H, gr, c = oracle.ndt3_assemble(g["point"], g["mean"], g["sqrt_info"], g["R"][:2, :2], g["t"][:2],
                                    2, [1.0])
    np.testing.assert_allclose(H, g["H6"], rtol=1e-12)
    np.testing.assert_allclose(gr, g["g3"], rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(c, g["cost3"], rtol=1e-12)


def test_fixture_sizes(oracle):
    room = oracle.room_points()
    assert room.shape == (954605, 3)                 # results/maha_amd64_simple.txt:1
    np.testing.assert_array_equal(room, syn.room_points())
    grid = syn.room_ndt_grid(1.0)
    assert int(grid["valid"].sum()) == 96            # "Ndt map size: 96", :2


def test_voxel_key(oracle):
    # zig-zag fold + Cantor pairing, tests/simple_optimization_test.cc:282-294
    def ref(p, inv):
        k = [int(np.floor(v * inv)) for v in p]
        k = [2 * v if v >= 0 else -2 * v - 1 for v in k]
        xy = (k[0] + k[1]) * (k[0] + k[1] + 1) // 2 + k[1]
        return (xy + k[2]) * (xy + k[2] + 1) // 2 + k[2]
    rng = np.random.default_rng(0)
    for _ in range(200):
        p = rng.uniform(-4, 4, 3)
        assert oracle.voxel_key(p, 10.0) == ref(p, 10.0)
    assert oracle.voxel_key([0.0, 0.0, 0.0], 1.0) == 0
    # the 0.1 m voxel filter of the fixture keeps 9 356 points
    room = oracle.room_points()[::1]
    keys = {}
    kept = 0
    inv = 10.0
    k = np.floor(room * inv).astype(np.int64)
    k = np.where(k >= 0, 2 * k, -2 * k - 1)
    xy = (k[:, 0] + k[:, 1]) * (k[:, 0] + k[:, 1] + 1) // 2 + k[:, 1]
    key = (xy + k[:, 2]) * (xy + k[:, 2] + 1) // 2 + k[:, 2]
    assert len(np.unique(key)) == 9356


def _fd_jacobian(fn, x0, eps=1e-6):
    r0 = fn(x0)
    J = np.zeros((r0.size, x0.size))
    for k in range(x0.size):
        d = np.zeros_like(x0); d[k] = eps
        J[:, k] = (fn(x0 + d) - fn(x0 - d)) / (2 * eps)
    return J


def _exp_so3(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)


def test_ndt6_jacobian_is_the_derivative(oracle):
    rng = np.random.default_rng(1)
    for _ in range(10):
        R = syn.random_rotation(rng); t = rng.normal(size=3)
        p = rng.normal(size=3); mu = rng.normal(size=3); S = rng.normal(size=(3, 3))
        J, r = oracle.ndt6_jacobian_residual(R, t, p, mu, S)
        np.testing.assert_allclose(r, S @ (R @ p + t - mu), atol=1e-13)
        # update convention of Solve: t += d[:3]; R <- R * Exp(d[3:])  (..._analytic.cc:134-135)
        fn = lambda d: S @ (R @ _exp_so3(d[3:]) @ p + t + d[:3] - mu)
        np.testing.assert_allclose(J, _fd_jacobian(fn, np.zeros(6)), atol=1e-7)


def test_ndt3_jacobian_is_the_derivative(oracle):
    rng = np.random.default_rng(2)
    for _ in range(10):
        a = rng.uniform(-1, 1)
        R2 = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]); t2 = rng.normal(size=2)
        p = rng.normal(size=3); mu = rng.normal(size=3); S = rng.normal(size=(3, 3))
        J, r = oracle.ndt3_jacobian_residual(R2, t2, p, mu, S)

        def fn(d):
            c, s = np.cos(d[2]), np.sin(d[2])
            Rn = R2 @ np.array([[c, -s], [s, c]])     # Isometry2d::rotate, ..._analytic_3dof.cc:83
            u = Rn @ p[:2] + t2 + d[:2]
            return S @ (np.array([u[0], u[1], p[2]]) - mu)
        np.testing.assert_allclose(r, fn(np.zeros(3)), atol=1e-13)
        np.testing.assert_allclose(J, _fd_jacobian(fn, np.zeros(3)), atol=1e-7)


def test_reproj_jacobian_is_the_derivative_and_depth_gate(oracle):
    rng = np.random.default_rng(3)
    K = syn.PNP_INTRINSICS
    for _ in range(10):
        R = syn.random_rotation(rng, 0.2); t = rng.normal(size=3) * 0.1
        X = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(2, 4)])
        px = rng.uniform(0, 480, 2)
        J, r = oracle.reproj_jacobian_residual(R, t, X, px, K)

        def fn(d):
            Xw = R @ _exp_so3(d[3:]) @ X + t + d[:3]
            return np.array([Xw[0] / Xw[2] - K[4] * (px[0] - K[2]), Xw[1] / Xw[2] - K[5] * (px[1] - K[3])])
        np.testing.assert_allclose(r, fn(np.zeros(6)), atol=1e-13)
        np.testing.assert_allclose(J, _fd_jacobian(fn, np.zeros(6)), atol=1e-7)
    J, r = oracle.reproj_jacobian_residual(np.eye(3), np.zeros(3), [0.1, 0.2, 0.029], [1.0, 2.0], K)
    assert not J.any() and not r.any()               # kMinDepth = 0.03, ..._analytic.cc:111,119-123
    J, r = oracle.reproj_jacobian_residual(np.eye(3), np.zeros(3), [0.1, 0.2, 0.031], [1.0, 2.0], K)
    assert J.any()


def test_loss_functions(oracle):
    # Exponential: loss_function.h:28-33
    out = oracle.loss(oracle.LOSS_EXPONENTIAL, [2.0, 0.5], 3.0)
    e = np.exp(-1.5)
    np.testing.assert_allclose(out, [2.0 - 2.0 * e, 2.0 * e, -2.0 * 0.5 * 2.0 * e], rtol=1e-15)
    # Huber: :57-66 (strict > on the squared threshold)
    np.testing.assert_allclose(oracle.loss(oracle.LOSS_HUBER, [2.0], 4.0)[:2], [4.0, 1.0])
    np.testing.assert_allclose(oracle.loss(oracle.LOSS_HUBER, [2.0], 9.0)[:2], [2 * 2 * 3 - 4, 2.0 / 3.0])
    # weight = 2 rho'(s) for Exponential (c1 c2 e^{-c2 s} * 2), rho'(s) for Huber/Cauchy
    for kind, params, factor in [(1, [1.3, 0.7], 2.0), (2, [0.5], 1.0), (3, [0.8], 1.0)]:
        for s in [0.01, 0.3, 2.0, 7.0]:
            h = 1e-6 * max(s, 1.0)
            d = (oracle.loss(kind, params, s + h)[0] - oracle.loss(kind, params, s - h)[0]) / (2 * h)
            np.testing.assert_allclose(oracle.loss(kind, params, s)[1], factor * d, rtol=1e-6)
    assert oracle.loss(1, [1, 1], 0.0)[0] == 0.0 and oracle.loss(3, [1], 0.0)[0] == 0.0


def test_rotation_helpers(oracle):
    rng = np.random.default_rng(4)
    for _ in range(50):
        R = syn.random_rotation(rng, 3.1)
        q = oracle.rotmat_to_quat(R)
        np.testing.assert_allclose(np.linalg.norm(q), 1.0, atol=1e-14)
        np.testing.assert_allclose(oracle.quat_to_rotmat(q), R, atol=1e-14)


def test_long_double_and_thread_split(oracle):
    point, mean, S = syn.random_ndt_records(10007, seed=3)
    R = syn.random_rotation(np.random.default_rng(5)); t = np.array([0.1, -0.2, 0.05])
    a = oracle.ndt6_assemble(point, mean, S, R, t, 1, [1.0, 1.0])
    b = oracle.ndt6_assemble(point, mean, S, R, t, 1, [1.0, 1.0], long_double=True)
    np.testing.assert_allclose(a[0], b[0], rtol=1e-11)
    np.testing.assert_allclose(a[2], b[2], rtol=1e-12)
    # executor split (..._analytic.cc:59-73): chunks of floor(N/T), the N mod T tail is dropped
    T = 4
    nb = 10007 // T
    full = oracle.ndt6_assemble(point, mean, S, R, t, 1, [1.0, 1.0], end=T * nb, long_double=True)
    thr = oracle.ndt6_assemble_threads(point, mean, S, R, t, 1, [1.0, 1.0], num_threads=T)
    np.testing.assert_allclose(thr[0], full[0], rtol=1e-11)
    np.testing.assert_allclose(thr[2], full[2], rtol=1e-12)
    p1 = oracle.ndt6_solve(point[:2000], mean[:2000], S[:2000], _pose0(oracle), 1, [1.0, 1.0], num_threads=0)
    p2 = oracle.ndt6_solve(point[:2000], mean[:2000], S[:2000], _pose0(oracle), 1, [1.0, 1.0], num_threads=2)
    assert p1[1] == p2[1]
    np.testing.assert_allclose(p1[0], p2[0], atol=1e-9)


def test_simd_float_baseline_agrees_at_float_tolerance(oracle):
    point, mean, S = syn.ndt_problem(20000, 1001, syn.CFG1_TRUE)
    n = (len(point) // 8) * 8
    planes = oracle.simd_pack(point, mean, S)
    R = np.eye(3); t = np.zeros(3)
    ref = oracle.ndt6_assemble(point, mean, S, R, t, 1, [1.0, 1.0], end=n)
    for T in (1, 3):
        nT = 8 * ((len(point) // 8) // T) * T
        ref = oracle.ndt6_assemble(point, mean, S, R, t, 1, [1.0, 1.0], end=nT)
        got = oracle.simd_ndt6_assemble(planes, len(point), R, t, 1, [1.0, 1.0], num_threads=T)
        np.testing.assert_allclose(got[0], ref[0], rtol=2e-3, atol=1e-3 * np.abs(ref[0]).max())
        np.testing.assert_allclose(got[2], ref[2], rtol=2e-3)


def test_simd_planar_and_reprojection_baselines_agree_at_float_tolerance(oracle):
    """The float twins timed beside cfg2 / cfg3 compute what the double paths compute (up to float
    rounding and the reference's own quirks in the reprojection twin)."""
    point, mean, S = syn.ndt_problem(20000, 1002, syn.CFG2_TRUE)
    planes = oracle.simd_pack(point, mean, S)
    T = syn.yaw_pose([0.01, -0.02, 0.0], 0.03)
    for threads in (1, 3):
        nT = 8 * ((len(point) // 8) // threads) * threads
        ref = oracle.ndt3_assemble(point, mean, S, T[:2, :2], T[:2, 3], 2, [1.0], end=nT)
        got = oracle.simd_ndt3_assemble(planes, len(point), T[:2, :2], T[:2, 3], 2, [1.0], num_threads=threads)
        np.testing.assert_allclose(got[0], ref[0], rtol=2e-3, atol=1e-3 * np.abs(ref[0]).max())
        np.testing.assert_allclose(got[1], ref[1], rtol=0, atol=2e-3 * np.abs(ref[0]).max())
        np.testing.assert_allclose(got[2], ref[2], rtol=2e-3)
    X, px, K = syn.pnp_problem(4000, 1003)
    X[::9, 2] = -1.0                                  # behind the camera
    planes = oracle.simd_reproj_pack(X, px)
    R = syn.random_rotation(np.random.default_rng(6), 0.05); t = np.array([0.01, 0.02, -0.03])
    n8 = (len(X) // 8) * 8
    ref = oracle.reproj_assemble(X, px, K, R, t, 0, None, end=n8)
    got = oracle.simd_reproj_assemble(planes, len(X), K, R, t, 0, None, num_threads=1)
    np.testing.assert_allclose(got[0], ref[0], rtol=5e-3, atol=2e-3 * np.abs(ref[0]).max())
    np.testing.assert_allclose(got[1], ref[1], rtol=0, atol=5e-3 * np.abs(ref[1]).max())
    # the quirk: without a loss the twin's cost is sum ||r|| over ALL lanes, the gated ones included
    Rq = R
    Xw = X[:n8] @ Rq.T + t
    r = np.stack([Xw[:, 0] / Xw[:, 2] - K[4] * (px[:n8, 0] - K[2]), Xw[:, 1] / Xw[:, 2] - K[5] * (px[:n8, 1] - K[3])], 1)
    np.testing.assert_allclose(got[2], np.linalg.norm(r, axis=1).sum(), rtol=2e-3)


@pytest.mark.timeout(300)
def test_ndt_fixture_cost_band(oracle):
    """Reference fixture end to end (room -> 0.1 m voxel filter -> 1.0 m NDT map -> KD-tree
    radius match, <= 2 neighbours -> Solve, re-match up to 10 times), Exponential(1,1):
    results/maha_amd64_simple.txt:10-14  COST 17438.4/40, 17394.5/40, 17490.6/20, (17490.7/2)."""
    from scipy.spatial import cKDTree
    room = oracle.room_points()
    k = np.floor(room * 10.0).astype(np.int64)
    _, first = np.unique(k, axis=0, return_index=True)
    filtered = room[np.sort(first)]
    assert len(filtered) == 9356
    Tinv = np.linalg.inv(syn.CFG1_TRUE)
    local = filtered @ Tinv[:3, :3].T + Tinv[:3, 3]
    grid = syn.room_ndt_grid(1.0)
    vidx = np.nonzero(grid["valid"])[0]
    tree = cKDTree(grid["mean"][vidx])
    pose = _pose0(oracle)
    costs, iters = [], []
    for outer in range(10):
        R, t = oracle.pose_to_Rt(pose)
        w = local @ R.T + t
        d, idx = tree.query(w, k=2, distance_upper_bound=1.0)
        sel = np.isfinite(d)
        pi = np.repeat(np.arange(len(local)), 2).reshape(-1, 2)[sel]
        ci = vidx[idx[sel]]
        last = pose.copy()
        pose, it, cost, _ = oracle.ndt6_solve(local[pi], grid["mean"][ci], grid["sqrt_info"][ci], pose,
                                              1, [1.0, 1.0])
        costs.append(cost); iters.append(it)
        Ra, ta = oracle.pose_to_Rt(last); Rb, tb = oracle.pose_to_Rt(pose)
        dq = oracle.rotmat_to_quat(Ra.T @ Rb)
        if np.linalg.norm(ta - tb) < 1e-5 and np.linalg.norm(dq[:3]) < 1e-5:
            break
    # The fixture writes sqrt_information = diag * V (not V^T), so S^T S -- and with it the cost --
    # depends on the arbitrary SIGN of each eigenvector returned by the eigen-solver (Eigen there,
    # LAPACK here).  That makes the log a sanity band, not a known answer: 2 % on the cost, the
    # iteration caps of the first two solves, and the recovered translation to 5 mm.
    assert iters[0] == 40 and iters[1] == 40
    for got, ref in zip(costs[:3], [17438.4, 17394.5, 17490.6]):
        assert abs(got - ref) / ref < 0.02, (costs, iters)
    R, t = oracle.pose_to_Rt(pose)
    np.testing.assert_allclose(t, [-0.196416, 0.121469, 0.304836], atol=5e-3)  # :24


def _fixture_loop(oracle, grid, true_T, solve):
    """OptimizePoseAnalytic* of the reference's test mains (simple_optimization_test.cc:473-505,
    3dof_6dof_comparison_test.cc:382-413): <= 10 x {match <= 2 nearest means within 1 m, Solve}."""
    from scipy.spatial import cKDTree
    room = oracle.room_points()
    k = np.floor(room * 10.0).astype(np.int64)
    _, first = np.unique(k, axis=0, return_index=True)
    filtered = room[np.sort(first)]
    Tinv = np.linalg.inv(true_T)
    local = filtered @ Tinv[:3, :3].T + Tinv[:3, 3]
    vidx = np.nonzero(grid["valid"])[0]
    tree = cKDTree(grid["mean"][vidx])
    pose = _pose0(oracle)
    costs, iters = [], []
    for _ in range(10):
        R, t = oracle.pose_to_Rt(pose)
        d, idx = tree.query(local @ R.T + t, k=2, distance_upper_bound=1.0)
        sel = np.isfinite(d)
        pi = np.repeat(np.arange(len(local)), 2).reshape(-1, 2)[sel]
        ci = vidx[idx[sel]]
        last = pose.copy()
        pose, it, cost, _ = solve(local[pi], grid["mean"][ci], grid["sqrt_info"][ci], pose, 1, [1.0, 1.0])
        costs.append(cost); iters.append(it)
        Ra, ta = oracle.pose_to_Rt(last); Rb, tb = oracle.pose_to_Rt(pose)
        dq = oracle.rotmat_to_quat(Ra.T @ Rb)
        if np.linalg.norm(ta - tb) < 1e-5 and np.linalg.norm(dq[:3]) < 1e-5:
            break
    return costs, iters, pose


def test_eigen_solver_restatement(oracle):
    """oracle.eigen_selfadjoint3 (Eigen's SelfAdjointEigenSolver<Matrix3d>::compute restated) is an
    eigen-decomposition: A = V diag(w) V^T, V orthonormal, w ascending -- and on the fixture's map
    44 of the 96 cells have a two-fold DEGENERATE eigenvalue (a fully covered 1 m x 1 m patch of
    floor or wall: relative gap ~1e-13), i.e. an arbitrary basis of that eigenspace."""
    rng = np.random.default_rng(0)
    for _ in range(100):
        B = rng.normal(size=(3, 3)); A = B @ B.T
        w, V = oracle.eigen_selfadjoint3(A)
        assert np.all(np.diff(w) >= 0)
        np.testing.assert_allclose(V @ np.diag(w) @ V.T, A, atol=1e-12 * np.abs(A).max())
        np.testing.assert_allclose(V.T @ V, np.eye(3), atol=1e-12)
    grid = oracle.reference_ndt_grid(oracle.room_points(), 1.0)
    assert int(grid["valid"].sum()) == 96
    S = grid["sqrt_info"].reshape(-1, 3, 3)[grid["valid"] == 1]
    d = np.linalg.norm(S, axis=2)  # |row k of diag(d) V| = d_k
    degenerate = np.abs(d[:, 1] - d[:, 2]) / d[:, 2] < 1e-9
    assert int(degenerate.sum()) == 44


@pytest.mark.timeout(600)
def test_ndt_fixture_logs_with_eigen_restated(oracle):
    """The published NDT logs against the fixture built with Eigen's own eigen-solver algorithm
    (oracle.reference_ndt_grid): sqrt_information = diag * V as the reference writes it.
      results/maha_amd64_simple.txt:10-14   6-DoF  COST 17438.4/40, 17394.5/40, 17490.6/20, 17490.7/2
      results/maha_3_vs_6_amd64.txt:18-23   6-DoF  COST 17857.8/40, 17526.6/40, 17494.8/30, 17490.7/9, DBL_MAX/0
      results/maha_3_vs_6_amd64.txt:7-11    3-DoF  COST 17871.8/40, ...
    Costs agree to 0.1 %, the iteration caps, the number of outer rounds and even the last line of
    the second log (a Solve that starts converged: 0 iterations, cost DBL_MAX) are reproduced.  Not
    to the digit: with S = diag * V, S^T S = V^T D^2 V depends on which basis the solver happens to
    return for the degenerate eigenspace of 44 of the 96 cells (test above) -- rounding noise of
    the reference's build (-O3 -march=native, FMA contraction), which no restatement can pin."""
    grid = oracle.reference_ndt_grid(oracle.room_points(), 1.0)
    costs, iters, pose = _fixture_loop(oracle, grid, syn.CFG1_TRUE, oracle.ndt6_solve)
    assert iters[:2] == [40, 40] and len(iters) == 4 and abs(iters[2] - 20) <= 2 and iters[3] == 2
    for got, ref in zip(costs, [17438.4, 17394.5, 17490.6, 17490.7]):
        assert abs(got - ref) / ref < 1.5e-3, (costs, iters)
    _, t = oracle.pose_to_Rt(pose)
    np.testing.assert_allclose(t, [-0.196416, 0.121469, 0.304836], atol=2e-4)  # maha_amd64_simple.txt:24
    T2 = syn.yaw_pose([-0.15, 0.05, 0.0], 0.2)  # 3dof_6dof_comparison_test.cc:77-80
    costs, iters, pose = _fixture_loop(oracle, grid, T2, oracle.ndt6_solve)
    assert iters[:2] == [40, 40] and len(iters) == 5 and iters[4] == 0 and costs[4] > 1e300
    for got, ref in zip(costs[:4], [17857.8, 17526.6, 17494.8, 17490.7]):
        assert abs(got - ref) / ref < 1.5e-3, (costs, iters)
    R, t = oracle.pose_to_Rt(pose)
    np.testing.assert_allclose(t, [-0.145667, 0.0484111, 0.00497979], atol=2e-4)  # :33
    np.testing.assert_allclose(oracle.rotmat_to_quat(R), [-0.000218557, -0.001234, 0.099811, 0.995006], atol=1e-4)
    costs3, iters3, _ = _fixture_loop(oracle, grid, T2, oracle.ndt3_solve)
    assert iters3[0] == 40 and abs(costs3[0] - 17871.8) / 17871.8 < 1.5e-3  # :7


@pytest.mark.timeout(900)
def test_ndt3_fixture_cost_band(oracle):
    """results/maha_3_vs_6_amd64.txt:7-11,31: the planar minimizer on the same fixture,
    COST 17871.8/40, 17559.4/40, 17488.4/10, 17484.3/1, pose (-0.150014, 0.0478718), yaw 0.2000.
    The planar path is the one that feels the arbitrary basis of the degenerate eigenspaces (the
    6-DoF path can absorb the spurious y/z coupling in the three parameters the planar one lacks):
    a fixed basis choice lands anywhere in a band -- the restated Eigen arithmetic at y = -0.043,
    9 cm off.  So the pin is the band itself: over seeded random bases of the degenerate
    eigenspaces the reference's numbers must lie inside the ensemble (they do: x = -0.15000 +- 2e-5
    always, y in 0.045 .. 0.048, costs 17.8 - 18.0 k / 17.5 - 17.7 k / 17.5 - 17.6 k, iterations
    40 / 25 - 40 / 6 - 9 / 1 - 2).  No sign convention reproduces the log; see DESIGN.md."""
    base = oracle.reference_ndt_grid(oracle.room_points(), 1.0)
    T2 = syn.yaw_pose([-0.15, 0.05, 0.0], 0.2)
    S0 = base["sqrt_info"].reshape(-1, 3, 3)
    rng = np.random.default_rng(1)
    first, second, ys, xs, yaws = [], [], [], [], []
    for _ in range(5):
        S = S0.copy()
        for c in np.nonzero(base["valid"])[0]:
            d = np.linalg.norm(S[c], axis=1)
            if abs(d[1] - d[2]) / d[2] < 1e-9:
                V = S[c] / d[:, None]
                a = rng.uniform(0.0, 2.0 * np.pi)
                sgn = -1.0 if rng.random() < 0.5 else 1.0
                v1 = np.cos(a) * V[:, 1] + np.sin(a) * V[:, 2]
                v2 = sgn * (-np.sin(a) * V[:, 1] + np.cos(a) * V[:, 2])
                V[:, 1], V[:, 2] = v1, v2
                S[c] = d[:, None] * V
        grid = dict(base, sqrt_info=np.ascontiguousarray(S.reshape(-1, 9)))
        costs, iters, pose = _fixture_loop(oracle, grid, T2, oracle.ndt3_solve)
        assert iters[0] == 40 and 3 <= len(iters) <= 5
        R, t = oracle.pose_to_Rt(pose)
        first.append(costs[0]); second.append(costs[1]); xs.append(t[0]); ys.append(t[1])
        yaws.append(np.arctan2(R[1, 0], R[0, 0]))
    assert min(first) - 150 < 17871.8 < max(first) + 150, first
    assert min(second) - 80 < 17559.4 < max(second) + 80, second
    assert np.max(np.abs(np.array(xs) + 0.150014)) < 1e-4
    assert min(ys) - 2e-3 < 0.0478718 < max(ys) + 2e-3, ys
    assert np.max(np.abs(np.array(yaws) - 0.2)) < 5e-4
