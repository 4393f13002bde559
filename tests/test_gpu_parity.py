"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star): per-iteration H, g, cost within 1e-6 relative; converged
pose within 1e-6 m / 1e-6 rad at the same iteration count.  See tests/parity.py for the metric.
"""
import numpy as np
import pytest

from nonlinear_optimizer_for_slam_b200 import synthetic as syn
from parity import TOL, assert_sums_close, rotation_angle

pytestmark = pytest.mark.gpu

LOSSES = [(0, None), (1, [1.0, 1.0]), (1, [0.7, 2.5]), (2, [1.0]), (2, [0.05]), (3, [0.5])]


def _rand_pose(rng, nlo):
    R = syn.random_rotation(rng)
    t = rng.uniform(-0.5, 0.5, 3)
    return nlo.pose_from_Rt(R, t), R, t


@pytest.mark.parametrize("n", [1, 31, 256, 257, 5000, 70001])
@pytest.mark.parametrize("loss", LOSSES)
def test_ndt6_assemble_matches_oracle(ctx, nlo, oracle, n, loss):
    rng = np.random.default_rng(100 + n)
    point, mean, S = syn.random_ndt_records(n, seed=n)
    pose16, R, t = _rand_pose(rng, nlo)
    prob = nlo.NdtProblem(ctx, capacity=n)
    prob.upload(point, mean, S)
    ctx.set_loss(loss[0], loss[1])
    H, g, c = prob.assemble6(pose16)
    # the reference converts the pose through a quaternion first (..._analytic.cc:86-87,99)
    Rq = oracle.quat_to_rotmat(oracle.rotmat_to_quat(R))
    Hr, gr, cr = oracle.ndt6_assemble(point, mean, S, Rq, t, loss[0], loss[1], long_double=True)
    assert_sums_close(H, g, c, Hr, gr, cr)
    prob.close()


@pytest.mark.parametrize("n", [4, 255, 1024, 33333])
@pytest.mark.parametrize("loss", LOSSES)
def test_ndt3_assemble_matches_oracle(ctx, nlo, oracle, n, loss):
    rng = np.random.default_rng(200 + n)
    point, mean, S = syn.random_ndt_records(n, seed=1000 + n)
    yaw = rng.uniform(-0.4, 0.4)
    T = syn.yaw_pose([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), 0.0], yaw)
    prob = nlo.NdtProblem(ctx, capacity=n)
    prob.upload(point, mean, S)
    ctx.set_loss(loss[0], loss[1])
    H, g, c = prob.assemble3(syn.to_pose16(T))
    Hr, gr, cr = oracle.ndt3_assemble(point, mean, S, T[:2, :2], T[:2, 3], loss[0], loss[1],
                                      long_double=True)
    assert_sums_close(H, g, c, Hr, gr, cr)
    prob.close()


@pytest.mark.parametrize("n", [1, 630, 4097, 50000])
@pytest.mark.parametrize("loss", [(0, None), (1, [1.0, 1.0]), (2, [0.01]), (3, [0.01])])
def test_reproj_assemble_matches_oracle(ctx, nlo, oracle, n, loss):
    rng = np.random.default_rng(300 + n)
    X, px, K = syn.pnp_problem(n, seed=n)
    X[::7, 2] = -1.0  # behind the camera: depth gate (..._analytic.cc:119-123)
    pose16, R, t = _rand_pose(rng, nlo)
    prob = nlo.ReprojProblem(ctx, capacity=n)
    prob.upload(X, px, K)
    ctx.set_loss(loss[0], loss[1])
    H, g, c = prob.assemble(pose16)
    Rq = oracle.quat_to_rotmat(oracle.rotmat_to_quat(R))
    Hr, gr, cr = oracle.reproj_assemble(X, px, K, Rq, t, loss[0], loss[1], long_double=True)
    assert_sums_close(H, g, c, Hr, gr, cr)
    prob.close()


def test_assemble_ranges_and_empty(ctx, nlo, oracle):
    n = 10000
    point, mean, S = syn.random_ndt_records(n, seed=5)
    prob = nlo.NdtProblem(ctx, capacity=n + 100)
    prob.upload(point, mean, S)
    ctx.set_loss(1, [1.0, 1.0])
    pose16 = nlo.identity_pose()
    total = np.zeros(28)
    for b, e in [(0, 0), (0, 1), (1, 300), (300, 4097), (4097, 9999), (9999, 10000)]:
        H, g, c = prob.assemble6(pose16, b, e)
        Hr, gr, cr = oracle.ndt6_assemble(point, mean, S, np.eye(3), np.zeros(3), 1, [1.0, 1.0],
                                          begin=b, end=e, long_double=True)
        if e > b:
            assert_sums_close(H, g, c, Hr, gr, cr)
        else:
            assert not H.any() and not g.any() and c == 0.0
        total += np.concatenate([H, g, [c]])
    H, g, c = prob.assemble6(pose16)
    # linearity: shards sum to the whole (what the multi-GPU all-reduce relies on)
    assert_sums_close(total[:21], total[21:27], total[27], H, g, c, tol=1e-12)
    prob.close()


def _check_trajectory(res, ref, width, nlo, three_dof=False):
    pose_r, it_r, cost_r, trace_r = ref
    assert res["iterations"] == it_r
    assert res["trace"].shape == trace_r.shape
    nh = 6 if three_dof else 21
    ng = 3 if three_dof else 6
    for k in range(trace_r.shape[0]):
        a, b = res["trace"][k], trace_r[k]
        assert_sums_close(a[:nh], a[nh:nh + ng], a[nh + ng], b[:nh], b[nh:nh + ng], b[nh + ng])
        np.testing.assert_allclose(a[nh + ng + 1:], b[nh + ng + 1:], rtol=0, atol=1e-6)
    Ra, ta = nlo.pose_to_Rt(res["pose"])
    Rb, tb = nlo.pose_to_Rt(pose_r)
    assert np.max(np.abs(ta - tb)) < 1e-6
    assert rotation_angle(Ra, Rb) < 1e-6
    assert abs(res["final_cost"] - cost_r) <= TOL * abs(cost_r)


def test_pnp_known_answer(ctx, nlo, oracle):
    """results/reproj_amd64.txt:5,10 -- COST: 2.33228e-11, iter: 6, pose^-1 = (-0.1, 0.123, -0.5)."""
    X, px, K = syn.pnp_fixture()
    prob = nlo.ReprojProblem(ctx, capacity=len(X))
    prob.upload(X, px, K)
    ctx.set_loss(1, [1.0, 1.0])
    res = prob.solve(nlo.identity_pose(), trace=True)
    assert res["iterations"] == 6
    assert "%.5e" % res["final_cost"] == "2.33228e-11"
    R, t = nlo.pose_to_Rt(res["pose"])
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
    Ti = np.linalg.inv(T)
    np.testing.assert_allclose(Ti[:3, 3], [-0.1, 0.123, -0.5], atol=5e-7)
    q = oracle.rotmat_to_quat(Ti[:3, :3])
    np.testing.assert_allclose(q, [0, 0, 0.0499792, 0.99875], atol=5e-7)
    ref = oracle.reproj_solve(X, px, K, nlo.identity_pose(), 1, [1.0, 1.0])
    _check_trajectory(res, ref, 36, nlo)
    prob.close()


@pytest.mark.parametrize("n,loss", [(3000, (1, [1.0, 1.0])), (100000, (1, [1.0, 1.0])),
                                    (20000, (0, None)), (20000, (2, [1.0]))])
def test_ndt6_solve_trajectory(ctx, nlo, oracle, n, loss):
    point, mean, S = syn.ndt_problem(n, 1001, syn.CFG1_TRUE)
    prob = nlo.NdtProblem(ctx, capacity=len(point))
    prob.upload(point, mean, S)
    ctx.set_loss(loss[0], loss[1])
    res = prob.solve6(nlo.identity_pose(), trace=True)
    ref = oracle.ndt6_solve(point, mean, S, nlo.identity_pose(), loss[0], loss[1])
    _check_trajectory(res, ref, 36, nlo)
    prob.close()


@pytest.mark.parametrize("n,loss", [(2001, (2, [1.0])), (200000, (2, [1.0])), (30000, (1, [1.0, 1.0]))])
def test_ndt3_solve_trajectory(ctx, nlo, oracle, n, loss):
    point, mean, S = syn.ndt_problem(n, 1002, syn.CFG2_TRUE)
    prob = nlo.NdtProblem(ctx, capacity=len(point))
    prob.upload(point, mean, S)
    ctx.set_loss(loss[0], loss[1])
    init = syn.to_pose16(syn.yaw_pose([0.02, -0.01, 0.3], 0.03))  # z and 3-D part must survive
    res = prob.solve3(init, trace=True)
    ref = oracle.ndt3_solve(point, mean, S, init, loss[0], loss[1])
    _check_trajectory(res, ref, 17, nlo, three_dof=True)
    np.testing.assert_array_equal(res["pose"][[2, 6, 8, 9, 10, 14]], init[[2, 6, 8, 9, 10, 14]])
    prob.close()


# Launch-shape boundaries of the device-resident loop (nlo_api.cu::EnqueueLoop): <= 3 tiles run inside one
# CTA; up to 8 tiles per CTA of one CTA per SM run in the resident kernel (clusters of 8 CTAs: partly
# idle clusters at 4 - 9 tiles, an odd tile count per CTA pairs the last tile with a masked one, the
# last size that fits is ~132 x 8 tiles); anything larger streams.
@pytest.mark.parametrize("n", [769, 1025, 2049, 2305, 33793, 68000, 270000, 271000, 303105])
def test_ndt6_solve_trajectory_at_the_launch_shape_boundaries(ctx, nlo, oracle, n):
    point, mean, S = (a[:n] for a in syn.ndt_problem(n + n // 8 + 64, 1007, syn.CFG1_TRUE))
    assert len(point) == n   # the exact tile count matters here
    prob = nlo.NdtProblem(ctx, capacity=len(point))
    prob.upload(point, mean, S)
    ctx.set_loss(1, [1.0, 1.0])
    opts = nlo.Options(max_iterations=12)
    res = prob.solve6(nlo.identity_pose(), opts, trace=True)
    ref = oracle.ndt6_solve(point, mean, S, nlo.identity_pose(), 1, [1.0, 1.0], max_iterations=12)
    _check_trajectory(res, ref, 36, nlo)
    again = prob.solve6(nlo.identity_pose(), opts, trace=True)   # bitwise repeatable in every shape
    assert np.array_equal(res["trace"], again["trace"]) and np.array_equal(res["pose"], again["pose"])
    prob.close()


@pytest.mark.parametrize("n", [1025, 9000, 200001, 800000])
def test_planar_and_reprojection_at_the_launch_shape_boundaries(ctx, nlo, oracle, n):
    point, mean, S = syn.ndt_problem(min(n, 300000), 1008, syn.CFG2_TRUE)
    prob = nlo.NdtProblem(ctx, capacity=len(point))
    prob.upload(point, mean, S)
    ctx.set_loss(2, [1.0])
    opts = nlo.Options(max_iterations=10)
    res = prob.solve3(nlo.identity_pose(), opts, trace=True)
    ref = oracle.ndt3_solve(point, mean, S, nlo.identity_pose(), 2, [1.0], max_iterations=10)
    _check_trajectory(res, ref, 17, nlo, three_dof=True)
    prob.close()
    X, px, K = syn.pnp_problem(n, 1009)   # 800 000 is past what the resident kernel holds (10 KB tiles, 22 per CTA)
    rp = nlo.ReprojProblem(ctx, capacity=len(X))
    rp.upload(X, px, K)
    ctx.set_loss(3, [1e-2])
    res = rp.solve(nlo.identity_pose(), opts, trace=True)
    ref = oracle.reproj_solve(X, px, K, nlo.identity_pose(), 3, [1e-2], max_iterations=10)
    _check_trajectory(res, ref, 36, nlo)
    rp.close()


def test_reproj_solve_cauchy(ctx, nlo, oracle):
    X, px, K = syn.pnp_problem(50000, 1003)
    prob = nlo.ReprojProblem(ctx, capacity=len(X))
    prob.upload(X, px, K)
    ctx.set_loss(3, [1e-2])
    res = prob.solve(nlo.identity_pose(), trace=True)
    ref = oracle.reproj_solve(X, px, K, nlo.identity_pose(), 3, [1e-2])
    _check_trajectory(res, ref, 36, nlo)
    prob.close()


def test_batched_matches_single(ctx, nlo, oracle):
    rng = np.random.default_rng(7)
    counts = [2000, 777, 5000, 256, 1]
    pts, mus, Ss, poses, refs = [], [], [], [], []
    ctx.set_loss(1, [1.0, 1.0])
    for k, c in enumerate(counts):
        T = syn.yaw_pose(rng.uniform(-0.2, 0.2, 3), rng.uniform(-0.1, 0.1))
        p, m, s = syn.ndt_problem(c, 2000 + k, T)
        pts.append(p); mus.append(m); Ss.append(s)
        poses.append(nlo.identity_pose())
        refs.append(oracle.ndt6_solve(p, m, s, nlo.identity_pose(), 1, [1.0, 1.0]))
    prob = nlo.NdtProblem(ctx, counts=[len(p) for p in pts])
    prob.upload(np.concatenate(pts), np.concatenate(mus), np.concatenate(Ss))
    out = prob.solve6_batched(np.stack(poses))
    for k in range(len(counts)):
        pose_r, it_r, cost_r, _ = refs[k]
        assert out["iterations"][k] == it_r
        Ra, ta = nlo.pose_to_Rt(out["poses"][k]); Rb, tb = nlo.pose_to_Rt(pose_r)
        assert np.max(np.abs(ta - tb)) < 1e-6 and rotation_angle(Ra, Rb) < 1e-6
        assert abs(out["final_cost"][k] - cost_r) <= TOL * abs(cost_r)
        H, g, c = prob.assemble6(nlo.identity_pose(), problem_index=k)
        Hr, gr, cr = oracle.ndt6_assemble(pts[k], mus[k], Ss[k], np.eye(3), np.zeros(3), 1, [1.0, 1.0],
                                          long_double=True)
        assert_sums_close(H, g, c, Hr, gr, cr)
    prob.close()


def test_upload_aos_reference_layout(ctx, nlo, oracle):
    """std::vector<Correspondence> layout: point(24) | NDT{count(4)+pad, sum, moment, mean,
    information, sqrt_information (Eigen column-major), is_valid, is_planar} = 304 bytes."""
    n = 3000
    point, mean, S = syn.random_ndt_records(n, seed=11)
    stride = 304
    off_point, off_mean, off_sqrt = 0, 24 + 8 + 24 + 72, 24 + 8 + 24 + 72 + 24 + 72
    rec = np.zeros((n, stride), dtype=np.uint8)
    rec[:, off_point:off_point + 24] = point.view(np.uint8).reshape(n, 24)
    rec[:, off_mean:off_mean + 24] = mean.view(np.uint8).reshape(n, 24)
    S_col = np.ascontiguousarray(S.reshape(n, 3, 3).transpose(0, 2, 1)).reshape(n, 9)
    rec[:, off_sqrt:off_sqrt + 72] = S_col.view(np.uint8).reshape(n, 72)
    prob = nlo.NdtProblem(ctx, capacity=n)
    prob.upload_aos(rec, n, stride, off_point, off_mean, off_sqrt, True)
    p2, m2, info = prob.download(0, n)
    np.testing.assert_array_equal(p2, point)
    np.testing.assert_array_equal(m2, mean)
    # the device keeps S only through the information matrix S^T S (formed once at ingest)
    np.testing.assert_allclose(info, syn.information6(S), rtol=1e-13, atol=1e-13 * np.abs(S).max() ** 2)
    ctx.set_loss(1, [1.0, 1.0])
    H, g, c = prob.assemble6(nlo.identity_pose())
    Hr, gr, cr = oracle.ndt6_assemble(point, mean, S, np.eye(3), np.zeros(3), 1, [1.0, 1.0], long_double=True)
    assert_sums_close(H, g, c, Hr, gr, cr)
    prob.close()


def test_generate_matches_host_association(ctx, nlo):
    n = 200000
    grid = syn.room_ndt_grid(0.5)
    prob = nlo.NdtProblem(ctx, capacity=n)
    true16 = syn.to_pose16(syn.CFG1_TRUE)
    prob.generate(n, 1004, 0, 0.01, true16, nlo.identity_pose(), grid)
    point, mean, info = prob.download(0, n)
    # points lie on the room surfaces (in the world frame) up to the noise
    w = point @ syn.CFG1_TRUE[:3, :3].T + syn.CFG1_TRUE[:3, 3]
    d = np.minimum.reduce([np.abs(w[:, 2]), np.abs(w[:, 1] + 2.5), np.abs(w[:, 1] - 2.5),
                           np.abs(w[:, 0] + 3.5), np.abs(w[:, 0] - 3.5)])
    assert d.max() < 0.08 and 0.004 < d.std() < 0.02
    _, m_ref, s_ref = syn.associate_dense(point, np.eye(4), grid, keep_unmatched=True)
    info_ref = syn.information6(s_ref)
    mismatch = np.any(mean != m_ref, axis=1) | np.any(np.abs(info - info_ref) > 1e-12 * (1 + np.abs(info_ref)), axis=1)
    assert mismatch.mean() < 1e-4
    assert (np.abs(info).sum(1) > 0).mean() > 0.99
    # a different offset continues the same stream
    prob2 = nlo.NdtProblem(ctx, capacity=1000)
    prob2.generate(1000, 1004, 5000, 0.01, true16, nlo.identity_pose(), grid)
    p2, _, _ = prob2.download(0, 1000)
    np.testing.assert_array_equal(p2, point[5000:6000])
    prob.close(); prob2.close()


@pytest.mark.parametrize("n", [4_000_000, 64 * 1024 * 1024])
def test_large_generated_shards_sum_and_repeat(ctx, nlo, n):
    """Size-independent properties up to BASELINE's full cfg4 size (64M points, 8 GB of planes):
    point-range shards sum to the whole (linearity, what the multi-GPU all-reduce relies on), the
    reduction is deterministic (bitwise repeatable), and a solve is repeatable bit for bit."""
    grid = syn.room_ndt_grid(0.5)
    prob = nlo.NdtProblem(ctx, capacity=n)
    prob.generate(n, 1004, 0, 0.01, syn.to_pose16(syn.CFG1_TRUE), nlo.identity_pose(), grid)
    ctx.set_loss(1, [1.0, 1.0])
    pose = nlo.identity_pose()
    H, g, c = prob.assemble6(pose)
    H2, g2, c2 = prob.assemble6(pose)
    assert np.array_equal(H, H2) and np.array_equal(g, g2) and c == c2
    acc = np.zeros(28)
    cuts = sorted({0, 1_000_003, 2_500_000, 2_500_001, n // 2 + 17, n})
    for b, e in zip(cuts[:-1], cuts[1:]):
        Hs, gs, cs = prob.assemble6(pose, b, e)
        acc += np.concatenate([Hs, gs, [cs]])
    assert_sums_close(acc[:21], acc[21:27], acc[27], H, g, c, tol=1e-11)
    r1 = prob.solve6(pose, nlo.Options(max_iterations=5), trace=True)
    r2 = prob.solve6(pose, nlo.Options(max_iterations=5), trace=True)
    assert np.array_equal(r1["trace"], r2["trace"]) and np.array_equal(r1["pose"], r2["pose"])
    # the first trace row is the assembly at the initial pose (the persistent loop pre-reduces the
    # per-CTA sums inside thread-block clusters, the single assembly adds them flat: same numbers
    # up to the order of a few hundred fp64 additions)
    assert_sums_close(r1["trace"][0, :21], r1["trace"][0, 21:27], r1["trace"][0, 27], H, g, c, tol=1e-12)
    prob.close()


def test_batched_copies_are_identical_and_match_single(ctx, nlo):
    """cfg5-shaped property: B copies of one registration solved in one batched launch give B
    bit-identical results, equal to the single-problem path within the parity tolerance."""
    grid = syn.room_ndt_grid(0.5)
    B, n = 300, 20000
    true = np.tile(syn.to_pose16(syn.CFG1_TRUE), (B, 1))
    ctx.set_loss(1, [1.0, 1.0])
    prob = nlo.NdtProblem(ctx, counts=[n] * B)
    # every registration from the same stream: seeds differ by +k, so rebuild with equal seeds
    single = nlo.NdtProblem(ctx, capacity=n)
    single.generate(n, 77, 0, 0.01, true[0], nlo.identity_pose(), grid)
    p, m, info = single.download(0, n)
    s = syn.sqrt_info_from_information6(info)
    single.upload(p, m, s)                      # both sides from the same host arrays
    prob.upload(np.tile(p, (B, 1)), np.tile(m, (B, 1)), np.tile(s, (B, 1)))
    out = prob.solve6_batched(np.tile(nlo.identity_pose(), (B, 1)))
    assert np.all(out["iterations"] == out["iterations"][0])
    assert np.all(out["poses"] == out["poses"][0])
    ref = single.solve6(nlo.identity_pose())
    assert ref["iterations"] == out["iterations"][0]
    assert np.max(np.abs(ref["pose"] - out["poses"][0])) < 1e-9
    prob.close(); single.close()


def test_batched_generate_matches_single_generate(ctx, nlo, oracle):
    """cfg5 generator: registration k of a batched problem = a single problem generated from
    stream seed + k with the same true pose; the batched solve equals the oracle on those points."""
    grid = syn.room_ndt_grid(0.5)
    rng = np.random.default_rng(3)
    counts = [20000, 3000, 511]
    true = np.stack([syn.to_pose16(syn.yaw_pose(rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.15, 0.15)))
                     for _ in counts])
    ctx.set_loss(1, [1.0, 1.0])
    prob = nlo.NdtProblem(ctx, counts=counts)
    prob.generate_batched(2000, 0.01, true, nlo.identity_pose(), grid)
    out = prob.solve6_batched(np.tile(nlo.identity_pose(), (len(counts), 1)))
    for k, c in enumerate(counts):
        single = nlo.NdtProblem(ctx, capacity=c)
        single.generate(c, 2000 + k, 0, 0.01, true[k], nlo.identity_pose(), grid)
        p, m, info = single.download(0, c)
        s = syn.sqrt_info_from_information6(info)   # any S with S^T S = information is equivalent
        pose_r, it_r, cost_r, _ = oracle.ndt6_solve(p, m, s, nlo.identity_pose(), 1, [1.0, 1.0])
        assert out["iterations"][k] == it_r
        Ra, ta = nlo.pose_to_Rt(out["poses"][k]); Rb, tb = nlo.pose_to_Rt(pose_r)
        assert np.max(np.abs(ta - tb)) < 1e-6 and rotation_angle(Ra, Rb) < 1e-6
        assert abs(out["final_cost"][k] - cost_r) <= TOL * abs(cost_r)
        single.close()
    prob.close()


def test_batched_3dof_and_reprojection(ctx, nlo, oracle):
    """Batched twins of the planar and PnP minimizers: every registration equals its own oracle
    solve (the planar one with the reference's floor(n/4)*4 truncation per registration)."""
    rng = np.random.default_rng(21)
    # --- 3-DoF, Huber
    counts = [1003, 4000, 258]
    pts, mus, Ss, refs = [], [], [], []
    init = syn.to_pose16(syn.yaw_pose([0.01, 0.02, 0.1], 0.01))
    for k, c in enumerate(counts):
        T = syn.yaw_pose([rng.uniform(-0.2, 0.2), rng.uniform(-0.2, 0.2), 0.0], rng.uniform(-0.1, 0.1))
        p, m, s = syn.ndt_problem(c, 3000 + k, T)
        pts.append(p); mus.append(m); Ss.append(s)
        refs.append(oracle.ndt3_solve(p, m, s, init, 2, [1.0]))
    ctx.set_loss(2, [1.0])
    prob = nlo.NdtProblem(ctx, counts=[len(p) for p in pts])
    prob.upload(np.concatenate(pts), np.concatenate(mus), np.concatenate(Ss))
    out = prob.solve3_batched(np.tile(init, (len(counts), 1)))
    for k in range(len(counts)):
        pose_r, it_r, cost_r, _ = refs[k]
        assert out["iterations"][k] == it_r
        np.testing.assert_allclose(out["poses"][k], pose_r, rtol=0, atol=1e-6)
        assert abs(out["final_cost"][k] - cost_r) <= TOL * abs(cost_r)
    # the same problem object still serves the 6-DoF path with full ranges
    ctx.set_loss(1, [1.0, 1.0])
    out6 = prob.solve6_batched(np.tile(nlo.identity_pose(), (len(counts), 1)))
    ref6 = oracle.ndt6_solve(pts[0], mus[0], Ss[0], nlo.identity_pose(), 1, [1.0, 1.0])
    assert out6["iterations"][0] == ref6[1]
    np.testing.assert_allclose(out6["poses"][0], ref6[0], rtol=0, atol=1e-6)
    prob.close()
    # --- reprojection, Cauchy
    counts = [630, 5000, 300]
    Xs, pxs, refs = [], [], []
    for k, c in enumerate(counts):
        T = syn.yaw_pose([rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), rng.uniform(-0.3, 0.0)],
                         rng.uniform(-0.1, 0.1))
        X, px, K = syn.pnp_problem(c, 4000 + k, true_T=T)
        Xs.append(X); pxs.append(px)
        refs.append(oracle.reproj_solve(X, px, K, nlo.identity_pose(), 3, [1e-2]))
    ctx.set_loss(3, [1e-2])
    rp = nlo.ReprojProblem(ctx, counts=counts)
    rp.upload(np.concatenate(Xs), np.concatenate(pxs), K)
    out = rp.solve_batched(np.tile(nlo.identity_pose(), (len(counts), 1)))
    for k in range(len(counts)):
        pose_r, it_r, cost_r, _ = refs[k]
        assert out["iterations"][k] == it_r
        np.testing.assert_allclose(out["poses"][k], pose_r, rtol=0, atol=1e-6)
        assert abs(out["final_cost"][k] - cost_r) <= TOL * abs(cost_r)
    rp.close()


@pytest.mark.parametrize("axis,angle", [((0, 0, 1), np.pi), ((1, 0, 0), np.pi), ((0, 1, 0), 3.0),
                                        ((1, 1, 0), np.pi), ((0.3, -0.5, 0.8), 2.9)])
def test_initial_pose_large_rotations(ctx, nlo, oracle, axis, angle):
    """Initial poses with trace(R) <= 0 exercise the three largest-diagonal branches of
    Eigen's Quaterniond(Matrix3d) (..._analytic.cc:87), restated on the device in init_states."""
    axis = np.asarray(axis, dtype=np.float64); axis /= np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)
    assert np.trace(R) <= 0.0
    pose16 = nlo.pose_from_Rt(R, [0.1, -0.2, 0.3])
    point, mean, S = syn.random_ndt_records(3000, seed=31)
    prob = nlo.NdtProblem(ctx, capacity=3000)
    prob.upload(point, mean, S)
    ctx.set_loss(1, [1.0, 1.0])
    H, g, c = prob.assemble6(pose16)
    Rq = oracle.quat_to_rotmat(oracle.rotmat_to_quat(R))
    Hr, gr, cr = oracle.ndt6_assemble(point, mean, S, Rq, [0.1, -0.2, 0.3], 1, [1.0, 1.0], long_double=True)
    assert_sums_close(H, g, c, Hr, gr, cr)
    res = prob.solve6(pose16, nlo.Options(max_iterations=3), trace=True)
    ref = oracle.ndt6_solve(point, mean, S, pose16, 1, [1.0, 1.0], max_iterations=3)
    _check_trajectory(res, ref, 36, nlo)
    prob.close()


def test_convergence_by_gradient_and_by_step(ctx, nlo, oracle):
    """Both break conditions (..._analytic.cc:139-144) with the iteration count they leave behind."""
    rng = np.random.default_rng(8)
    point = rng.uniform(-2, 2, (4000, 3))
    T = syn.yaw_pose([0.05, -0.03, 0.02], 0.01)
    mean = point @ T[:3, :3].T + T[:3, 3]             # exact correspondences: cost -> 0
    S = np.tile(np.eye(3).reshape(9), (4000, 1))
    prob = nlo.NdtProblem(ctx, capacity=4000)
    prob.upload(point, mean, S)
    ctx.set_loss(0)
    for opts in (dict(parameter_tolerance=1e-6, gradient_tolerance=1e-300),   # stops on the step norm
                 dict(parameter_tolerance=1e-300, gradient_tolerance=1e-3),   # stops on the gradient norm
                 dict(parameter_tolerance=1e-300, gradient_tolerance=1e-300)):  # runs to the cap
        res = prob.solve6(nlo.identity_pose(), nlo.Options(max_iterations=25, **opts), trace=True)
        ref = oracle.ndt6_solve(point, mean, S, nlo.identity_pose(), 0, None, max_iterations=25, **opts)
        assert res["iterations"] == ref[1]
        np.testing.assert_allclose(res["pose"], ref[0], rtol=0, atol=1e-6)
    assert res["iterations"] == 25
    Rr, tr = nlo.pose_to_Rt(res["pose"])
    np.testing.assert_allclose(tr, T[:3, 3], atol=1e-9)
    prob.close()


def test_loss_edge_values(ctx, nlo, oracle):
    """Huber exactly at / around its threshold (strict >, loss_function.h:58), exponential weights
    that underflow to zero, Cauchy with huge residuals -- sums stay finite and match the oracle."""
    n = 2048
    point = np.zeros((n, 3)); mean = np.zeros((n, 3)); S = np.zeros((n, 9))
    S[:, 0] = 1.0; S[:, 4] = 1.0; S[:, 8] = 1.0
    r = np.linspace(0.0, 4.0, n)
    r[100] = 2.0                                           # s == threshold^2 exactly
    r[101] = np.nextafter(2.0, 3.0); r[102] = np.nextafter(2.0, 1.0)
    r[-1] = 40.0; r[-2] = 1e3                              # exp(-s) underflows
    mean[:, 0] = -r                                        # e = (r, 0, 0) at the identity pose
    prob = nlo.NdtProblem(ctx, capacity=n)
    prob.upload(point, mean, S)
    for kind, params in ((2, [2.0]), (1, [1.0, 1.0]), (3, [0.1]), (1, [3.0, 50.0])):
        ctx.set_loss(kind, params)
        H, g, c = prob.assemble6(nlo.identity_pose())
        Hr, gr, cr = oracle.ndt6_assemble(point, mean, S, np.eye(3), np.zeros(3), kind, params, long_double=True)
        assert np.isfinite(H).all() and np.isfinite(g).all() and np.isfinite(c)
        assert_sums_close(H, g, c, Hr, gr, cr)
    prob.close()


def test_reprojection_all_points_behind_camera(ctx, nlo):
    """Every correspondence gated out (z < 0.03): H = 0, the damped system is singular, the solve
    reports NLO_ENUMERIC instead of returning garbage (the reference would write NaNs and return true)."""
    X, px, K = syn.pnp_fixture()
    X = X.copy(); X[:, 2] = -1.0
    rp = nlo.ReprojProblem(ctx, capacity=len(X))
    rp.upload(X, px, K)
    ctx.set_loss(0)
    H, g, c = rp.assemble(nlo.identity_pose())
    assert not H.any() and not g.any() and c == 0.0
    with pytest.raises(nlo.NloError) as e:
        rp.solve(nlo.identity_pose())
    assert e.value.code == -5
    rp.close()


def test_point_to_plane_is_a_rank_one_ndt_record(ctx, nlo):
    """the point-to-plane residual r = n^T (R p + t - q) of the reference's unfinished
    pose_optimizer/cost_functors.h is the NDT residual with sqrt_information = [n^T; 0; 0] and
    mean = q, so the same kernel assembles it.  Checked against normal equations built in numpy
    straight from the point-to-plane definition."""
    rng = np.random.default_rng(12)
    n = 5000
    p = rng.uniform(-2, 2, (n, 3))
    normal = rng.normal(size=(n, 3)); normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    T = syn.yaw_pose([0.04, -0.02, 0.03], 0.02)
    q = p @ T[:3, :3].T + T[:3, 3] + rng.normal(0, 0.01, (n, 3))
    S = np.zeros((n, 9)); S[:, 0:3] = normal
    prob = nlo.NdtProblem(ctx, capacity=n)
    prob.upload(p, q, S)
    ctx.set_loss(0)
    R0 = syn.random_rotation(rng, 0.05); t0 = np.array([0.01, 0.0, -0.01])
    Rq = R0  # already a rotation; the quaternion round trip only re-normalises it
    H, g, c = prob.assemble6(nlo.pose_from_Rt(R0, t0))
    r = np.einsum("ni,ni->n", normal, p @ Rq.T + t0 - q)
    J = np.zeros((n, 6))
    J[:, :3] = normal
    skew = np.zeros((n, 3, 3))
    skew[:, 0, 1] = -p[:, 2]; skew[:, 0, 2] = p[:, 1]; skew[:, 1, 0] = p[:, 2]
    skew[:, 1, 2] = -p[:, 0]; skew[:, 2, 0] = -p[:, 1]; skew[:, 2, 1] = p[:, 0]
    J[:, 3:] = -np.einsum("ni,nij->nj", normal @ Rq, skew)
    Href = J.T @ J
    iu = np.triu_indices(6)
    assert_sums_close(H, g, c, Href[iu], J.T @ r, float(r @ r), tol=1e-9)
    res = prob.solve6(nlo.identity_pose(), nlo.Options(max_iterations=60))
    Rr, tr = nlo.pose_to_Rt(res["pose"])
    np.testing.assert_allclose(tr, T[:3, 3], atol=2e-3)     # point-to-plane ICP recovers the motion
    assert rotation_angle(Rr, T[:3, :3]) < 2e-3
    prob.close()


def test_small_batch_uses_many_ctas_per_registration(ctx, nlo, oracle):
    """A batch with fewer registrations than SMs runs as one persistent grid with several CTAs per
    registration (own leader, counter and published state each); results must equal the oracle and
    the one-CTA-per-registration path (a batch too large for that shape)."""
    rng = np.random.default_rng(17)
    counts = [30000, 5000, 12345, 800, 20001]
    pts, mus, Ss, refs = [], [], [], []
    ctx.set_loss(1, [1.0, 1.0])
    for k, c in enumerate(counts):
        T = syn.yaw_pose(rng.uniform(-0.2, 0.2, 3), rng.uniform(-0.1, 0.1))
        p, m, s = syn.ndt_problem(c, 6000 + k, T)
        pts.append(p); mus.append(m); Ss.append(s)
        refs.append(oracle.ndt6_solve(p, m, s, nlo.identity_pose(), 1, [1.0, 1.0]))
    prob = nlo.NdtProblem(ctx, counts=[len(p) for p in pts])
    prob.upload(np.concatenate(pts), np.concatenate(mus), np.concatenate(Ss))
    out = prob.solve6_batched(np.tile(nlo.identity_pose(), (len(counts), 1)))
    for k in range(len(counts)):
        pose_r, it_r, cost_r, _ = refs[k]
        assert out["iterations"][k] == it_r
        np.testing.assert_allclose(out["poses"][k], pose_r, rtol=0, atol=1e-6)
        assert abs(out["final_cost"][k] - cost_r) <= TOL * abs(cost_r)
    # 3-DoF twin through the same shape
    ctx.set_loss(2, [1.0])
    out3 = prob.solve3_batched(np.tile(nlo.identity_pose(), (len(counts), 1)))
    ref3 = oracle.ndt3_solve(pts[0], mus[0], Ss[0], nlo.identity_pose(), 2, [1.0])
    assert out3["iterations"][0] == ref3[1]
    np.testing.assert_allclose(out3["poses"][0], ref3[0], rtol=0, atol=1e-6)
    prob.close()


def _f32_round(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def _well_conditioned(S, floor=0.05):
    """The same records with the singular values of every S raised to >= floor * the largest, so that
    S^T S stays positive definite after rounding to float (the test rebuilds an S from it)."""
    U, d, Vt = np.linalg.svd(np.asarray(S).reshape(-1, 3, 3))
    d = np.maximum(d, floor * d[:, :1])
    return np.einsum("nij,nj,njk->nik", U, d, Vt).reshape(-1, 9)


@pytest.mark.parametrize("n", [1, 255, 4097, 70001])
def test_f32_storage_equals_oracle_on_float_rounded_records(ctx, nlo, oracle, n):
    """Opt-in throughput mode: records (point, mean, S^T S) stored as float, arithmetic in fp64.  It
    must equal the double oracle evaluated on the float-rounded records to the normal parity bar."""
    rng = np.random.default_rng(500 + n)
    point, mean, S = syn.random_ndt_records(n, seed=900 + n)
    S = _well_conditioned(S)
    pose16, R, t = _rand_pose(rng, nlo)
    prob = nlo.NdtProblem(ctx, capacity=n, storage="f32")
    p64 = nlo.NdtProblem(ctx, capacity=n)
    p64.upload(point, mean, S)
    i64 = p64.download(0, n)[2]
    p64.close()
    prob.upload_f32(point, mean, S)           # float host arrays: S is rounded before S^T S is formed
    p3, m3, i3 = prob.download(0, n)
    np.testing.assert_array_equal(p3, _f32_round(point))
    np.testing.assert_array_equal(m3, _f32_round(mean))
    ref3 = syn.information6(_f32_round(S))
    np.testing.assert_allclose(i3, ref3, rtol=2e-7, atol=2e-7 * np.abs(ref3).max())
    prob.upload(point, mean, S)               # double host arrays: S^T S in fp64, rounded once
    p2, m2, i2 = prob.download(0, n)
    np.testing.assert_array_equal(p2, _f32_round(point))
    np.testing.assert_array_equal(m2, _f32_round(mean))
    np.testing.assert_array_equal(i2, _f32_round(i64))
    S2 = syn.sqrt_info_from_information6(i2)  # an S whose S^T S is the stored information
    np.testing.assert_allclose(syn.information6(S2), i2, rtol=1e-13, atol=1e-13 * np.abs(i2).max())
    Rq = oracle.quat_to_rotmat(oracle.rotmat_to_quat(R))
    for loss in [(0, None), (1, [1.0, 1.0]), (2, [1.0])]:
        ctx.set_loss(loss[0], loss[1])
        H, g, c = prob.assemble6(pose16)
        Hr, gr, cr = oracle.ndt6_assemble(p2, m2, S2, Rq, t, loss[0], loss[1], long_double=True)
        assert_sums_close(H, g, c, Hr, gr, cr)
        H3, g3, c3 = prob.assemble3(syn.to_pose16(syn.yaw_pose([0.1, -0.1, 0.0], 0.05)))
        T = syn.yaw_pose([0.1, -0.1, 0.0], 0.05)
        Hr3, gr3, cr3 = oracle.ndt3_assemble(p2, m2, S2, T[:2, :2], T[:2, 3], loss[0], loss[1], long_double=True)
        assert_sums_close(H3, g3, c3, Hr3, gr3, cr3)
    prob.close()


def test_f32_storage_solve_trajectory_and_quantisation_error(ctx, nlo, oracle):
    point, mean, S = syn.ndt_problem(100000, 1001, syn.CFG1_TRUE)
    ctx.set_loss(1, [1.0, 1.0])
    p32 = nlo.NdtProblem(ctx, capacity=len(point), storage="f32")
    p32.upload(point, mean, S)
    res = p32.solve6(nlo.identity_pose(), trace=True)
    pf, mf, inf = p32.download(0, len(point))
    ref = oracle.ndt6_solve(pf, mf, syn.sqrt_info_from_information6(inf), nlo.identity_pose(), 1, [1.0, 1.0])
    _check_trajectory(res, ref, 36, nlo)
    # against the unrounded double problem: only the input quantisation (float eps ~ 6e-8) shows
    p64 = nlo.NdtProblem(ctx, capacity=len(point))
    p64.upload(point, mean, S)
    H32, g32, c32 = p32.assemble6(nlo.identity_pose())
    H64, g64, c64 = p64.assemble6(nlo.identity_pose())
    from parity import rel_errors
    eh, eg, ec = rel_errors(H32, g32, c32, H64, g64, c64)
    assert eh < 1e-6 and ec < 1e-6 and eg < 1e-4, (eh, eg, ec)
    res64 = p64.solve6(nlo.identity_pose())
    Ra, ta = nlo.pose_to_Rt(res["pose"]); Rb, tb = nlo.pose_to_Rt(res64["pose"])
    assert np.max(np.abs(ta - tb)) < 1e-5 and rotation_angle(Ra, Rb) < 1e-5
    # generated problems agree between the storage types up to the rounding of the stored values
    grid = syn.room_ndt_grid(0.5)
    g32p = nlo.NdtProblem(ctx, capacity=5000, storage="f32")
    g64p = nlo.NdtProblem(ctx, capacity=5000)
    for pr in (g32p, g64p):
        pr.generate(5000, 9, 0, 0.01, syn.to_pose16(syn.CFG1_TRUE), nlo.identity_pose(), grid)
    a = g32p.download(0, 5000); b = g64p.download(0, 5000)
    np.testing.assert_array_equal(a[0], _f32_round(b[0]))
    np.testing.assert_array_equal(a[1], _f32_round(b[1]))
    np.testing.assert_array_equal(a[2], _f32_round(b[2]))
    for pr in (p32, p64, g32p, g64p):
        pr.close()
