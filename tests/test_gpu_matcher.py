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
    # planar variant runs through the same loop
    res3 = scan.register(ndt_map, nlo.identity_pose(), three_dof=True)
    assert res3["outer_iterations"] >= 1
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
