"""GPU: the widened rows of the scope table -- device NDT map builder, device
matcher and the outer registration loop -- against numpy restatements of the reference's test-main
helpers (mahalanobis_distance_minimizer/tests/simple_optimization_test.cc:236-342,473-505)."""
import numpy as np
import pytest

from nonlinear_optimizer_for_slam_b200 import synthetic as syn
from parity import rotation_angle

pytestmark = pytest.mark.gpu


def brute_force_match(points, T, grid, radius=1.0, max_neighbors=2):
    """MatchPointCloud with an exhaustive search: per point the <= 2 nearest valid means with
    squared distance < radius^2; returns index arrays [n, max_neighbors] (-1 = none)."""
    w = points @ T[:3, :3].T + T[:3, 3]
    vidx = np.nonzero(grid["valid"])[0]
    means = grid["mean"][vidx]
    out = np.full((len(points), max_neighbors), -1, dtype=np.int64)
    for b in range(0, len(points), 20000):
        e = w[b:b + 20000, None, :] - means[None, :, :]
        d2 = (e[..., 0] * e[..., 0] + e[..., 1] * e[..., 1]) + e[..., 2] * e[..., 2]
        order = np.argsort(d2, axis=1, kind="stable")[:, :max_neighbors]
        dsel = np.take_along_axis(d2, order, axis=1)
        out[b:b + 20000] = np.where(dsel < radius * radius, vidx[order], -1)
    return out


@pytest.mark.parametrize("voxel", [1.0, 0.5])
def test_match_equals_exhaustive_search(ctx, nlo, voxel):
    grid = syn.room_ndt_grid(voxel)
    rng = np.random.default_rng(5)
    world = syn.room_surface_samples(30000, rng, 0.02)
    Tinv = np.linalg.inv(syn.CFG1_TRUE)
    local = world @ Tinv[:3, :3].T + Tinv[:3, 3]
    T = syn.yaw_pose([0.03, -0.02, 0.05], 0.02)
    ndt_map = nlo.NdtMap(ctx, grid=grid)
    scan = nlo.Scan(ctx, local)
    prob = nlo.NdtProblem(ctx, capacity=2 * len(local))
    matched = scan.match(ndt_map, syn.to_pose16(T), prob)
    ref = brute_force_match(local, T, grid)
    assert matched == int((ref >= 0).sum())
    point, mean, info = prob.download(0, 2 * len(local))
    n = len(local)
    cell_info = syn.information6(grid["sqrt_info"])
    for j in range(2):
        has = ref[:, j] >= 0
        np.testing.assert_array_equal(point[j * n:(j + 1) * n], local)
        np.testing.assert_array_equal(mean[j * n:(j + 1) * n][has], grid["mean"][ref[has, j]])
        # entries that cancel to ~0 differ in the last bits between numpy and the device (FMA)
        np.testing.assert_allclose(info[j * n:(j + 1) * n][has], cell_info[ref[has, j]], rtol=1e-13,
                                   atol=1e-13 * np.abs(cell_info).max())
        assert not info[j * n:(j + 1) * n][~has].any()
    prob.close(); scan.close(); ndt_map.close()


def test_register_matches_oracle_outer_loop(ctx, nlo, oracle):
    """Whole registration: the device outer loop vs the same loop with a numpy matcher + the oracle
    solver; same number of rounds and inner iterations, pose within 1e-6."""
    grid = syn.room_ndt_grid(1.0)
    rng = np.random.default_rng(9)
    world = syn.room_surface_samples(20000, rng, 0.01)
    Tinv = np.linalg.inv(syn.CFG1_TRUE)
    local = world @ Tinv[:3, :3].T + Tinv[:3, 3]
    ctx.set_loss(1, [1.0, 1.0])
    ndt_map = nlo.NdtMap(ctx, grid=grid)
    scan = nlo.Scan(ctx, local)
    res = scan.register(ndt_map, nlo.identity_pose())
    # reference loop on the host
    pose = nlo.identity_pose()
    outer_done, inner = 0, 0
    for outer in range(10):
        R, t = nlo.pose_to_Rt(pose)
        T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
        ref = brute_force_match(local, T, grid)
        pts, mus, Ss = [], [], []
        for j in range(2):
            has = ref[:, j] >= 0
            pts.append(local[has]); mus.append(grid["mean"][ref[has, j]]); Ss.append(grid["sqrt_info"][ref[has, j]])
        last = pose.copy()
        pose, it, cost, _ = oracle.ndt6_solve(np.concatenate(pts), np.concatenate(mus), np.concatenate(Ss),
                                              pose, 1, [1.0, 1.0])
        outer_done += 1; inner += it
        Ra, ta = nlo.pose_to_Rt(last); Rb, tb = nlo.pose_to_Rt(pose)
        dq = oracle.rotmat_to_quat(Ra.T @ Rb)
        if np.linalg.norm(Ra.T @ (tb - ta)) < 1e-5 and np.linalg.norm(dq[:3]) < 1e-5:
            break
    assert res["outer_iterations"] == outer_done
    assert res["inner_iterations"] == inner
    Ra, ta = nlo.pose_to_Rt(res["pose"]); Rb, tb = nlo.pose_to_Rt(pose)
    assert np.max(np.abs(ta - tb)) < 1e-6 and rotation_angle(Ra, Rb) < 1e-6
    # and it actually registers the scan (the reference's own runs end a few mm from the truth)
    np.testing.assert_allclose(tb, syn.CFG1_TRUE[:3, 3], atol=1e-2)
    scan.close(); ndt_map.close()


@pytest.mark.parametrize("n_points", [20001, 20002, 20003])
def test_register_3dof_matches_oracle_outer_loop(ctx, nlo, oracle, n_points):
    """The planar registration loop of 3dof_6dof_comparison_test.cc: the reference hands the planar
    minimizer MatchPointCloud's POINT-MAJOR list of real hits and the minimizer drops the last
    M mod 4 of them (..._analytic_3dof.cc:33-36).  The device loop must drop exactly those: same
    rounds, same inner iterations, same pose as the host loop with the numpy matcher + oracle."""
    grid = syn.room_ndt_grid(1.0)
    rng = np.random.default_rng(n_points)
    world = syn.room_surface_samples(n_points, rng, 0.01)
    Tinv = np.linalg.inv(syn.CFG2_TRUE)
    local = world @ Tinv[:3, :3].T + Tinv[:3, 3]
    ctx.set_loss(1, [1.0, 1.0])
    ndt_map = nlo.NdtMap(ctx, grid=grid)
    scan = nlo.Scan(ctx, local)
    res = scan.register(ndt_map, nlo.identity_pose(), three_dof=True)
    pose = nlo.identity_pose()
    outer_done, inner, remainders = 0, 0, set()
    for outer in range(10):
        R, t = nlo.pose_to_Rt(pose)
        T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
        ref = brute_force_match(local, T, grid)
        has = ref >= 0                                   # [n, 2], point-major order when flattened
        pidx = np.repeat(np.arange(len(local)), 2).reshape(-1, 2)[has]
        cidx = ref[has]
        remainders.add(len(pidx) % 4)
        last = pose.copy()
        pose, it, cost, _ = oracle.ndt3_solve(local[pidx], grid["mean"][cidx], grid["sqrt_info"][cidx], pose,
                                              1, [1.0, 1.0])
        outer_done += 1; inner += it
        Ra, ta = nlo.pose_to_Rt(last); Rb, tb = nlo.pose_to_Rt(pose)
        dq = oracle.rotmat_to_quat(Ra.T @ Rb)
        if np.linalg.norm(Ra.T @ (tb - ta)) < 1e-5 and np.linalg.norm(dq[:3]) < 1e-5:
            break
    assert res["outer_iterations"] == outer_done
    assert res["inner_iterations"] == inner
    Ra, ta = nlo.pose_to_Rt(res["pose"]); Rb, tb = nlo.pose_to_Rt(pose)
    assert np.max(np.abs(ta - tb)) < 1e-6 and rotation_angle(Ra, Rb) < 1e-6
    assert res["pose"][14] == 0.0                        # z untouched (..._analytic_3dof.cc:104-105)
    scan.close(); ndt_map.close()


@pytest.mark.parametrize("voxel", [1.0, 0.5])
def test_map_build_matches_numpy(ctx, nlo, voxel):
    points = syn.room_points()
    ndt_map = nlo.NdtMap(ctx, points=points, voxel=voxel, v_not_transposed=False)
    got = ndt_map.to_grid()
    ref = syn.build_ndt_grid(points, voxel, v_not_transposed=False)
    np.testing.assert_array_equal(got["dims"], ref["dims"])
    np.testing.assert_allclose(got["origin"], ref["origin"])
    np.testing.assert_array_equal(got["valid"], ref["valid"])
    if voxel == 1.0:
        assert got["valid_cells"] == 96          # results/maha_amd64_simple.txt:2
    v = ref["valid"] != 0
    np.testing.assert_allclose(got["mean"][v], ref["mean"][v], rtol=0, atol=1e-10)
    # S is unique only up to the eigenvector basis; the information matrix S^T S is not
    Sg = got["sqrt_info"][v].reshape(-1, 3, 3); Sr = ref["sqrt_info"][v].reshape(-1, 3, 3)
    Ig = np.einsum("nki,nkj->nij", Sg, Sg); Ir = np.einsum("nki,nkj->nij", Sr, Sr)
    np.testing.assert_allclose(Ig, Ir, rtol=1e-7, atol=1e-7 * np.abs(Ir).max())
    assert not got["sqrt_info"][~v].any()
    # the reference's literal form (diag * V) is also available; same eigenvalues => same row norms
    quirk = nlo.NdtMap(ctx, points=points, voxel=voxel, v_not_transposed=True).to_grid()
    np.testing.assert_allclose(np.linalg.norm(quirk["sqrt_info"][v].reshape(-1, 3, 3), axis=2),
                               np.linalg.norm(Sg, axis=2), rtol=1e-7)
    ndt_map.close()


def decode_voxel_keys(keys):
    """x | y << 21 | z << 42 (include/nlo_cuda.h, nlo_ndt_map_download_keys)."""
    mask = np.uint64((1 << 21) - 1)
    return ((keys & mask).astype(np.int64), ((keys >> np.uint64(21)) & mask).astype(np.int64),
            ((keys >> np.uint64(42)) & mask).astype(np.int64))


@pytest.mark.parametrize("voxel", [1.0, 0.5])
def test_hashed_map_holds_the_dense_map(ctx, nlo, voxel):
    """The voxel hash (the reference's unordered_map, simple_optimization_test.cc:282-294) holds
    exactly the occupied voxels of the dense grid, with the same NDT per voxel."""
    points = syn.room_points()
    dense = nlo.NdtMap(ctx, points=points, voxel=voxel).to_grid()
    sparse_map = nlo.NdtMap(ctx, points=points, voxel=voxel, hashed=True)
    sparse = sparse_map.to_grid()
    assert sparse["hashed"] and not dense["hashed"]
    np.testing.assert_array_equal(sparse["dims"], dense["dims"])
    np.testing.assert_allclose(sparse["origin"], dense["origin"])
    assert sparse["valid_cells"] == dense["valid_cells"]
    used = sparse["keys"] != np.uint64(0xFFFFFFFFFFFFFFFF)
    slots = len(sparse["keys"])
    assert slots & (slots - 1) == 0 and 2 * used.sum() <= slots
    # occupied voxels = voxels that hold at least one point
    k = np.floor(points / voxel).astype(np.int64)
    k -= k.min(axis=0)
    dx, dy, dz = (int(v) for v in dense["dims"])
    occupied = np.unique((k[:, 2] * dy + k[:, 1]) * dx + k[:, 0])
    x, y, z = decode_voxel_keys(sparse["keys"][used])
    cell = (z * dy + y) * dx + x
    np.testing.assert_array_equal(np.sort(cell), occupied)
    np.testing.assert_array_equal(sparse["valid"][used], dense["valid"][cell])
    assert not sparse["valid"][~used].any()
    # the per-voxel sums are integers (fixed point), so both builds hold the same bits per voxel
    np.testing.assert_array_equal(sparse["mean"][used], dense["mean"][cell])
    np.testing.assert_array_equal(sparse["sqrt_info"][used], dense["sqrt_info"][cell])
    sparse_map.close()


@pytest.mark.parametrize("hashed", [False, True])
def test_map_build_and_registration_are_bitwise_repeatable(ctx, nlo, hashed):
    """The map accumulates 64-bit fixed-point sums (integer atomics: associative), so two builds of
    the same cloud are identical to the bit whatever order the atomics land in, and so is a whole
    registration (map build -> <= 10 x {match, Solve}) run twice from scratch."""
    points = syn.room_points()
    rng = np.random.default_rng(5)
    shuffled = points[rng.permutation(len(points))]      # a different arrival order, the same cloud
    grids, poses = [], []
    for cloud in (points, points, shuffled):
        ndt_map = nlo.NdtMap(ctx, points=cloud, voxel=1.0, hashed=hashed)
        grids.append(ndt_map.to_grid())
        filt = points[::97]
        Tinv = np.linalg.inv(syn.CFG1_TRUE)
        scan = nlo.Scan(ctx, filt @ Tinv[:3, :3].T + Tinv[:3, 3])
        ctx.set_loss(nlo.LOSS_EXPONENTIAL, [1.0, 1.0])
        poses.append(scan.register(ndt_map, nlo.identity_pose())["pose"].copy())
        scan.close(); ndt_map.close()
    def by_voxel(g):
        # a hashed map may seat a voxel in another slot from build to build (the slots are claimed by
        # racing atomicCAS): compare voxel by voxel, i.e. in key order
        if not hashed:
            return g["mean"], g["sqrt_info"]
        order = np.argsort(g["keys"], kind="stable")
        return g["keys"][order], g["mean"][order], g["sqrt_info"][order]
    for other in (1, 2):
        for a, b in zip(by_voxel(grids[0]), by_voxel(grids[other])):
            np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(poses[0], poses[other])


def test_hashed_match_equals_dense_match(ctx, nlo):
    points = syn.room_points()
    rng = np.random.default_rng(17)
    world = syn.room_surface_samples(30000, rng, 0.02)
    Tinv = np.linalg.inv(syn.CFG1_TRUE)
    local = world @ Tinv[:3, :3].T + Tinv[:3, 3]
    # a few points far outside the map and one absurdly far away: no neighbours, no overflow
    local[:3] = [[500.0, 0, 0], [0, -700.0, 3], [1e15, -1e15, 1e15]]
    T = syn.yaw_pose([0.03, -0.02, 0.05], 0.02)
    scan = nlo.Scan(ctx, local)
    n = len(local)
    got = []
    for hashed in (False, True):
        ndt_map = nlo.NdtMap(ctx, points=points, voxel=1.0, hashed=hashed)
        prob = nlo.NdtProblem(ctx, capacity=2 * n)
        matched = scan.match(ndt_map, syn.to_pose16(T), prob)
        got.append((matched,) + tuple(prob.download(0, 2 * n)))
        prob.close(); ndt_map.close()
    assert got[0][0] == got[1][0] > n
    np.testing.assert_array_equal(got[0][1], got[1][1])
    np.testing.assert_allclose(got[0][2], got[1][2], rtol=0, atol=1e-11)
    np.testing.assert_allclose(got[0][3], got[1][3], rtol=1e-8, atol=1e-8 * np.abs(got[0][3]).max())
    for j in range(2):
        assert not got[1][3][j * n:j * n + 3].any()
    scan.close()


def test_sparse_map_past_the_dense_limit_registers(ctx, nlo):
    """Two rooms 30 km apart: the bounding box has ~2e9 voxels (a dense grid is refused above
    2^28), the hash holds ~2 x the occupied ones; registration in the far room still lands."""
    shift = np.array([30000.0, 12000.0, 0.0])
    far_room = syn.room_points() + shift
    near_room = far_room - shift     # exact, so both rooms fall into voxels the same way
    one = nlo.NdtMap(ctx, points=near_room, voxel=1.0)
    valid_one = one.to_grid()["valid_cells"]
    one.close()
    ndt_map = nlo.NdtMap(ctx, points=np.concatenate([near_room, far_room]), voxel=1.0)   # plain build
    hashed, slots = ndt_map.layout()
    assert hashed and slots <= 4096
    g = ndt_map.to_grid()
    assert np.prod(g["dims"].astype(np.float64)) > 2 ** 28
    assert g["valid_cells"] == 2 * valid_one
    rng = np.random.default_rng(9)
    world = syn.room_surface_samples(20000, rng, 0.01)
    Tinv = np.linalg.inv(syn.CFG1_TRUE)
    local = world @ Tinv[:3, :3].T + Tinv[:3, 3]
    ctx.set_loss(1, [1.0, 1.0])
    scan = nlo.Scan(ctx, local)
    near = scan.register(ndt_map, nlo.identity_pose())
    T0 = np.eye(4); T0[:3, 3] = shift
    far = scan.register(ndt_map, syn.to_pose16(T0))
    Rn, tn = nlo.pose_to_Rt(near["pose"]); Rf, tf = nlo.pose_to_Rt(far["pose"])
    np.testing.assert_allclose(tn, syn.CFG1_TRUE[:3, 3], atol=1e-2)
    np.testing.assert_allclose(tf - shift, tn, atol=1e-6)
    assert rotation_angle(Rn, Rf) < 1e-6
    assert far["matched"] == near["matched"]
    scan.close(); ndt_map.close()
