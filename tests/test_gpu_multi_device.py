"""GPU: ONE context over several devices of this process (nlo_context_create_multi) -- the sharded
Solve behind the drop-in API.  With >= 2 GPUs the devices are distinct; on a one-GPU box the list
names device 0 twice, which runs the same code (two shards, two persistent iteration kernels, the
peer-memory all-reduce between them) on one device, so the sharded path is exercised everywhere.
Every result is compared with the CPU oracle run on the UNSHARDED problem."""
import numpy as np
import pytest

from nonlinear_optimizer_for_slam_b200 import synthetic as syn
from parity import TOL, assert_sums_close, rotation_angle

pytestmark = pytest.mark.gpu


def _device_lists():
    import torch
    n = torch.cuda.device_count()
    if n >= 2:
        lists = [list(range(2))]
        if n >= 4:
            lists.append(list(range(4)))
        if n >= 8:
            lists.append(list(range(8)))
        return lists
    return [[0, 0], [0, 0, 0]]


@pytest.fixture(scope="module", params=range(3))
def mctx(request, nlo):
    lists = _device_lists()
    if request.param >= len(lists):
        pytest.skip("fewer device lists on this box")
    c = nlo.Context(devices=lists[request.param])
    assert c.device_count == len(lists[request.param])
    yield c
    c.close()


def _check_solve(res, ref, nlo, nh, ng):
    pose_r, it_r, cost_r, trace_r = ref
    assert res["iterations"] == it_r, (res["iterations"], it_r)
    assert res["trace"].shape == trace_r.shape
    for k in range(trace_r.shape[0]):
        a, r = res["trace"][k], trace_r[k]
        assert_sums_close(a[:nh], a[nh:nh + ng], a[nh + ng], r[:nh], r[nh:nh + ng], r[nh + ng])
    Ra, ta = nlo.pose_to_Rt(res["pose"]); Rb, tb = nlo.pose_to_Rt(pose_r)
    assert np.max(np.abs(ta - tb)) < 1e-6 and rotation_angle(Ra, Rb) < 1e-6
    assert abs(res["final_cost"] - cost_r) <= TOL * abs(cost_r)


@pytest.mark.parametrize("n", [1, 1000, 300001])
def test_sharded_ndt6_assemble_solve_match_oracle(mctx, nlo, oracle, n):
    point, mean, S = syn.ndt_problem(n, 1004, syn.CFG1_TRUE)
    n = len(point)
    prob = nlo.NdtProblem(mctx, capacity=n)
    prob.upload(point, mean, S)
    assert prob.size == n
    mctx.set_loss(1, [1.0, 1.0])
    pose0 = nlo.identity_pose()
    H, g, c = prob.assemble6(pose0)
    Hr, gr, cr = oracle.ndt6_assemble(point, mean, S, np.eye(3), np.zeros(3), 1, [1.0, 1.0], long_double=True)
    assert_sums_close(H, g, c, Hr, gr, cr)
    if n > 10:   # a sub-range that cuts through the shards
        b, e = n // 7, n - n // 5
        H, g, c = prob.assemble6(pose0, b, e)
        Hr, gr, cr = oracle.ndt6_assemble(point, mean, S, np.eye(3), np.zeros(3), 1, [1.0, 1.0], begin=b, end=e,
                                          long_double=True)
        assert_sums_close(H, g, c, Hr, gr, cr)
        res = prob.solve6(pose0, trace=True)
        _check_solve(res, oracle.ndt6_solve(point, mean, S, pose0, 1, [1.0, 1.0]), nlo, 21, 6)
    # download walks the shards
    p2, m2, info = prob.download(0, n)
    np.testing.assert_array_equal(p2, point)
    np.testing.assert_array_equal(m2, mean)
    prob.close()


@pytest.mark.parametrize("n", [1203, 120003])
def test_sharded_planar_solve_drops_the_global_tail(mctx, nlo, oracle, n):
    """floor(n / 4) * 4 applies to the WHOLE correspondence list (..._analytic_3dof.cc:33-36), not
    to every shard: the sharded solve equals the oracle on the unsharded list."""
    point, mean, S = syn.ndt_problem(n, 1002, syn.CFG2_TRUE)
    keep = len(point) - (len(point) % 4) + 3        # force n mod 4 == 3
    point, mean, S = point[:keep], mean[:keep], S[:keep]
    assert len(point) % 4 == 3
    prob = nlo.NdtProblem(mctx, capacity=len(point))
    prob.upload(point, mean, S)
    mctx.set_loss(2, [1.0])
    init = syn.to_pose16(syn.yaw_pose([0.02, -0.01, 0.3], 0.03))
    res = prob.solve3(init, trace=True)
    _check_solve(res, oracle.ndt3_solve(point, mean, S, init, 2, [1.0]), nlo, 6, 3)
    np.testing.assert_array_equal(res["pose"][[2, 6, 8, 9, 10, 14]], init[[2, 6, 8, 9, 10, 14]])
    prob.close()


def test_sharded_reprojection_matches_oracle(mctx, nlo, oracle):
    X, px, K = syn.pnp_problem(50000, 1003)
    prob = nlo.ReprojProblem(mctx, capacity=len(X))
    prob.upload(X, px, K)
    mctx.set_loss(3, [1e-2])
    res = prob.solve(nlo.identity_pose(), trace=True)
    _check_solve(res, oracle.reproj_solve(X, px, K, nlo.identity_pose(), 3, [1e-2]), nlo, 21, 6)
    # the reference's known answer through the sharded path (results/reproj_amd64.txt:5)
    Xf, pxf, Kf = syn.pnp_fixture()
    fix = nlo.ReprojProblem(mctx, capacity=len(Xf))
    fix.upload(Xf, pxf, Kf)
    mctx.set_loss(1, [1.0, 1.0])
    r = fix.solve(nlo.identity_pose())
    assert r["iterations"] == 6 and "%.5e" % r["final_cost"] == "2.33228e-11"
    prob.close(); fix.close()


def test_sharded_aos_ingest_and_generate(mctx, nlo, oracle):
    n = 50001
    point, mean, S = syn.random_ndt_records(n, seed=12)
    stride, off_mean, off_sqrt = 304, 24 + 8 + 24 + 72, 24 + 8 + 24 + 72 + 24 + 72
    rec = np.zeros((n, stride), dtype=np.uint8)
    rec[:, 0:24] = point.view(np.uint8).reshape(n, 24)
    rec[:, off_mean:off_mean + 24] = mean.view(np.uint8).reshape(n, 24)
    S_col = np.ascontiguousarray(S.reshape(n, 3, 3).transpose(0, 2, 1)).reshape(n, 9)
    rec[:, off_sqrt:off_sqrt + 72] = S_col.view(np.uint8).reshape(n, 72)
    prob = nlo.NdtProblem(mctx, capacity=n)
    prob.upload_aos(rec, n, stride, 0, off_mean, off_sqrt, True)
    p2, m2, info = prob.download(0, n)
    np.testing.assert_array_equal(p2, point)
    np.testing.assert_array_equal(m2, mean)
    np.testing.assert_allclose(info, syn.information6(S), rtol=1e-13, atol=1e-13 * np.abs(S).max() ** 2)
    # the device generator yields the same stream whatever the number of shards
    grid = syn.room_ndt_grid(0.5)
    m = 40001
    prob.generate(m, 1004, 77, 0.01, syn.to_pose16(syn.CFG1_TRUE), nlo.identity_pose(), grid)
    pa, ma, ia = prob.download(0, m)
    import torch
    one = nlo.Context(0)
    single = nlo.NdtProblem(one, capacity=m)
    single.generate(m, 1004, 77, 0.01, syn.to_pose16(syn.CFG1_TRUE), nlo.identity_pose(), grid)
    pb, mb, ib = single.download(0, m)
    np.testing.assert_array_equal(pa, pb)
    np.testing.assert_array_equal(ma, mb)
    np.testing.assert_array_equal(ia, ib)
    single.close(); one.close(); prob.close()


def test_batched_problem_is_partitioned_by_registration(mctx, nlo, oracle):
    rng = np.random.default_rng(17)
    counts = [2000, 777, 5000, 256, 1, 3001, 1500]
    pts, mus, Ss, refs = [], [], [], []
    mctx.set_loss(1, [1.0, 1.0])
    for k, c in enumerate(counts):
        T = syn.yaw_pose(rng.uniform(-0.2, 0.2, 3), rng.uniform(-0.1, 0.1))
        p, m, s = syn.ndt_problem(c, 2000 + k, T)
        pts.append(p); mus.append(m); Ss.append(s)
        refs.append(oracle.ndt6_solve(p, m, s, nlo.identity_pose(), 1, [1.0, 1.0]))
    prob = nlo.NdtProblem(mctx, counts=[len(p) for p in pts])
    prob.upload(np.concatenate(pts), np.concatenate(mus), np.concatenate(Ss))
    out = prob.solve6_batched(np.tile(nlo.identity_pose(), (len(counts), 1)))
    for k in range(len(counts)):
        pose_r, it_r, cost_r, _ = refs[k]
        assert out["iterations"][k] == it_r
        Ra, ta = nlo.pose_to_Rt(out["poses"][k]); Rb, tb = nlo.pose_to_Rt(pose_r)
        assert np.max(np.abs(ta - tb)) < 1e-6 and rotation_angle(Ra, Rb) < 1e-6
        H, g, c = prob.assemble6(nlo.identity_pose(), problem_index=k)
        Hr, gr, cr = oracle.ndt6_assemble(pts[k], mus[k], Ss[k], np.eye(3), np.zeros(3), 1, [1.0, 1.0],
                                          long_double=True)
        assert_sums_close(H, g, c, Hr, gr, cr)
        p2, _, _ = prob.download(0, len(pts[k]), problem_index=k)
        np.testing.assert_array_equal(p2, pts[k])
    prob.close()


def test_registration_runs_on_the_first_device(mctx, nlo):
    grid = syn.room_ndt_grid(1.0)
    rng = np.random.default_rng(9)
    world = syn.room_surface_samples(5000, rng, 0.01)
    Tinv = np.linalg.inv(syn.CFG1_TRUE)
    local = world @ Tinv[:3, :3].T + Tinv[:3, 3]
    mctx.set_loss(1, [1.0, 1.0])
    ndt_map = nlo.NdtMap(mctx, grid=grid)
    scan = nlo.Scan(mctx, local)
    res = scan.register(ndt_map, nlo.identity_pose())
    _, t = nlo.pose_to_Rt(res["pose"])
    np.testing.assert_allclose(t, syn.CFG1_TRUE[:3, 3], atol=2e-2)
    scan.close(); ndt_map.close()


def test_communicator_calls_are_refused(mctx, nlo):
    with pytest.raises(nlo.NloError):
        mctx.comm_peer_export()
    with pytest.raises(nlo.NloError):
        mctx.comm_suspend(True)
