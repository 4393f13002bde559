"""GPU parity at BASELINE.json's FULL sizes, against the CPU oracle (not only self-consistency):

  cfg4  64M-point NDT 6-DoF scan: H, g, cost of one assembly pass vs the threaded oracle, the scan
        read back from the device in 2M-point chunks;
  cfg2  1M-point planar registration, Huber: every iteration of the trajectory vs the oracle;
  cfg5  4096 x 20k batched registrations: 8 sampled registrations vs the oracle.

Tolerance: 1e-6 relative on H, g, cost; 1e-6 m / 1e-6 rad on the pose at equal iteration counts
(BASELINE.json north_star).
"""
import numpy as np
import pytest

from nonlinear_optimizer_for_slam_b200 import synthetic as syn
from parity import TOL, assert_sums_close, rotation_angle

pytestmark = pytest.mark.gpu


def chol_sqrt_info(info6):
    """Row-major S[n,9] (upper triangular) with S^T S = the given information matrices
    (00 01 02 11 12 22); closed form, vectorised.  Zero records map to zero rows.  The minimizers
    see sqrt_information only through S^T S, so any such S is equivalent input for the oracle."""
    a, b, c, d, e, f = (np.asarray(info6, dtype=np.float64)[:, k] for k in range(6))
    with np.errstate(divide="ignore", invalid="ignore"):
        u00 = np.sqrt(np.maximum(a, 0.0))
        u01 = np.where(u00 > 0, b / u00, 0.0)
        u02 = np.where(u00 > 0, c / u00, 0.0)
        u11 = np.sqrt(np.maximum(d - u01 * u01, 0.0))
        u12 = np.where(u11 > 0, (e - u01 * u02) / u11, 0.0)
        u22 = np.sqrt(np.maximum(f - u02 * u02 - u12 * u12, 0.0))
    S = np.zeros((len(a), 9))
    S[:, 0] = u00; S[:, 1] = u01; S[:, 2] = u02; S[:, 4] = u11; S[:, 5] = u12; S[:, 8] = u22
    return S


@pytest.mark.timeout(900)
def test_cfg4_64m_assembly_matches_oracle(ctx, nlo, oracle):
    import os
    n = 64 * 1024 * 1024
    grid = syn.room_ndt_grid(0.5)
    prob = nlo.NdtProblem(ctx, capacity=n)
    prob.generate(n, 1004, 0, 0.01, syn.to_pose16(syn.CFG1_TRUE), nlo.identity_pose(), grid)
    ctx.set_loss(1, [1.0, 1.0])
    T = syn.yaw_pose([0.05, -0.02, 0.03], 0.02)      # not the identity: R enters the sums
    pose = syn.to_pose16(T)
    H, g, c = prob.assemble6(pose)
    Rq = oracle.quat_to_rotmat(oracle.rotmat_to_quat(T[:3, :3]))
    threads = max(1, min(os.cpu_count() or 1, 32))
    acc = np.zeros(28, dtype=np.longdouble)
    chunk = 2 * 1024 * 1024
    for b in range(0, n, chunk):
        p, m, info = prob.download(b, b + chunk)
        S = chol_sqrt_info(info)
        # chunk of T * floor(chunk / T) points on T threads, the remainder single-threaded
        per = (chunk // threads) * threads
        Hc, gc, cc = oracle.ndt6_assemble_threads(p[:per], m[:per], S[:per], Rq, T[:3, 3], 1, [1.0, 1.0], threads)
        acc += np.concatenate([Hc, gc, [cc]])
        if per < chunk:
            Hc, gc, cc = oracle.ndt6_assemble(p[per:], m[per:], S[per:], Rq, T[:3, 3], 1, [1.0, 1.0])
            acc += np.concatenate([Hc, gc, [cc]])
    acc = acc.astype(np.float64)
    assert_sums_close(H, g, c, acc[:21], acc[21:27], acc[27])
    # and the first trace row of a solve from that pose is the same assembly
    res = prob.solve6(pose, nlo.Options(max_iterations=2), trace=True)
    assert_sums_close(res["trace"][0, :21], res["trace"][0, 21:27], res["trace"][0, 27], acc[:21], acc[21:27], acc[27])
    prob.close()


@pytest.mark.timeout(600)
def test_cfg2_1m_planar_trajectory_matches_oracle(ctx, nlo, oracle):
    point, mean, S = syn.ndt_problem(1_000_000, 1002, syn.CFG2_TRUE)
    assert len(point) > 990_000
    prob = nlo.NdtProblem(ctx, capacity=len(point))
    prob.upload(point, mean, S)
    ctx.set_loss(2, [1.0])
    init = nlo.identity_pose()
    res = prob.solve3(init, trace=True)
    pose_r, it_r, cost_r, trace_r = oracle.ndt3_solve(point, mean, S, init, 2, [1.0])
    assert res["iterations"] == it_r
    assert res["trace"].shape == trace_r.shape
    for k in range(trace_r.shape[0]):
        a, r = res["trace"][k], trace_r[k]
        assert_sums_close(a[:6], a[6:9], a[9], r[:6], r[6:9], r[9])
        np.testing.assert_allclose(a[10:], r[10:], rtol=0, atol=1e-6)
    Ra, ta = nlo.pose_to_Rt(res["pose"]); Rb, tb = nlo.pose_to_Rt(pose_r)
    assert np.max(np.abs(ta - tb)) < 1e-6 and rotation_angle(Ra, Rb) < 1e-6
    assert abs(res["final_cost"] - cost_r) <= TOL * abs(cost_r)
    prob.close()


@pytest.mark.timeout(900)
def test_cfg5_batched_4096x20k_sampled_registrations_match_oracle(ctx, nlo, oracle):
    B, n = 4096, 20000
    grid = syn.room_ndt_grid(0.5)
    rng = np.random.default_rng(2000)
    true = np.stack([syn.to_pose16(syn.yaw_pose(rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.15, 0.15)))
                     for _ in range(B)])
    ctx.set_loss(1, [1.0, 1.0])
    prob = nlo.NdtProblem(ctx, counts=[n] * B)
    prob.generate_batched(2000, 0.01, true, nlo.identity_pose(), grid)
    out = prob.solve6_batched(np.tile(nlo.identity_pose(), (B, 1)))
    assert out["iterations"].min() >= 1
    for k in (0, 1, 511, 1024, 2047, 3000, 4094, 4095):
        p, m, info = prob.download(0, n, problem_index=k)
        S = chol_sqrt_info(info)
        pose_r, it_r, cost_r, _ = oracle.ndt6_solve(p, m, S, nlo.identity_pose(), 1, [1.0, 1.0])
        assert out["iterations"][k] == it_r, (k, out["iterations"][k], it_r)
        Ra, ta = nlo.pose_to_Rt(out["poses"][k]); Rb, tb = nlo.pose_to_Rt(pose_r)
        assert np.max(np.abs(ta - tb)) < 1e-6 and rotation_angle(Ra, Rb) < 1e-6, k
        assert abs(out["final_cost"][k] - cost_r) <= TOL * abs(cost_r)
        H, g, c = prob.assemble6(nlo.identity_pose(), problem_index=k)
        Hr, gr, cr = oracle.ndt6_assemble(p, m, S, np.eye(3), np.zeros(3), 1, [1.0, 1.0], long_double=True)
        assert_sums_close(H, g, c, Hr, gr, cr)
    prob.close()
