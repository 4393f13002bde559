"""GPU: error behaviour of the C ABI -- every misuse returns a negative code with a message and
leaves the context usable (the reference's Solve always returns true and cannot report anything;
its loss constructors throw std::out_of_range, loss_function.h:24-25,53-54)."""
import ctypes

import numpy as np
import pytest

from nonlinear_optimizer_for_slam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def test_bad_arguments_are_rejected_and_context_survives(ctx, nlo):
    with pytest.raises(nlo.NloError) as e:
        ctx.set_loss(nlo.LOSS_HUBER, [0.0])            # threshold must be > 0
    assert e.value.code == -1 and "threshold" in str(e.value)
    with pytest.raises(nlo.NloError):
        ctx.set_loss(nlo.LOSS_EXPONENTIAL, [-1.0, 1.0])
    with pytest.raises(nlo.NloError):
        ctx.set_loss(17, [1.0])
    with pytest.raises(nlo.NloError):
        nlo.NdtProblem(ctx, capacity=-5)
    point, mean, S = syn.random_ndt_records(100, seed=1)
    prob = nlo.NdtProblem(ctx, capacity=50)
    with pytest.raises(nlo.NloError):
        prob.upload(point, mean, S)                    # 100 > capacity 50
    prob.upload(point[:50], mean[:50], S[:50])
    with pytest.raises(nlo.NloError):
        prob.assemble6(nlo.identity_pose(), 10, 5)     # end < begin
    with pytest.raises(nlo.NloError):
        prob.assemble6(nlo.identity_pose(), 0, 51)     # end > n
    with pytest.raises(nlo.NloError):
        prob.solve6_batched(np.tile(nlo.identity_pose(), (1, 1)))  # not a batched problem
    rp = nlo.ReprojProblem(ctx, capacity=10)
    lib = nlo._capi.load()
    H = np.zeros(21); g = np.zeros(6); c = ctypes.c_double(0)
    pose = nlo.identity_pose()
    rc = lib.nlo_ndt6_assemble(ctx._h, rp._h, 0, pose.ctypes.data_as(nlo._capi.c_double_p), 0, 0,
                               H.ctypes.data_as(nlo._capi.c_double_p), g.ctypes.data_as(nlo._capi.c_double_p),
                               ctypes.byref(c))
    assert rc == -1 and b"family" in lib.nlo_last_error(ctx._h)   # NDT call on a reprojection problem
    # still healthy
    ctx.set_loss(nlo.LOSS_NONE)
    Hh, gg, cc = prob.assemble6(nlo.identity_pose())
    assert np.isfinite(Hh).all() and cc > 0
    prob.close(); rp.close()


def test_zero_iterations_and_empty_problem(ctx, nlo):
    prob = nlo.NdtProblem(ctx, capacity=10)
    prob.upload(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 9)))
    ctx.set_loss(nlo.LOSS_NONE)
    H, g, c = prob.assemble6(nlo.identity_pose())
    assert not H.any() and not g.any() and c == 0.0
    point, mean, S = syn.random_ndt_records(10, seed=2)
    prob.upload(point, mean, S)
    init = syn.to_pose16(syn.yaw_pose([0.1, 0.2, 0.3], 0.4))
    res = prob.solve6(init, nlo.Options(max_iterations=0))
    assert res["iterations"] == 0
    np.testing.assert_allclose(res["pose"], init, atol=1e-15)   # through quaternion and back
    assert res["final_cost"] == np.finfo(np.float64).max        # previous_cost never written
    prob.close()


def test_non_finite_input_is_reported(ctx, nlo):
    point, mean, S = syn.random_ndt_records(1000, seed=3)
    S[17, 4] = np.nan
    prob = nlo.NdtProblem(ctx, capacity=1000)
    prob.upload(point, mean, S)
    ctx.set_loss(nlo.LOSS_NONE)
    with pytest.raises(nlo.NloError) as e:
        prob.solve6(nlo.identity_pose())
    assert e.value.code == -5                                   # NLO_ENUMERIC
    S[17, 4] = 1.0
    prob.upload(point, mean, S)
    assert prob.solve6(nlo.identity_pose())["status"] == 0
    prob.close()


def test_singular_normal_equations_fall_back_to_pivoted_solve(ctx, nlo, oracle):
    """All points on one axis with rank-1 information: H is singular without damping and only
    positive semi-definite; the LDL^T path must hand over to the pivoted elimination and still
    agree with the oracle's general solve."""
    n = 512
    rng = np.random.default_rng(4)
    point = np.zeros((n, 3)); point[:, 0] = rng.uniform(-1, 1, n)
    mean = point + rng.normal(0, 0.01, (n, 3))
    S = np.zeros((n, 9)); S[:, 0] = 1.0                          # only e_x is observed
    prob = nlo.NdtProblem(ctx, capacity=n)
    prob.upload(point, mean, S)
    ctx.set_loss(nlo.LOSS_NONE)
    H, g, c = prob.assemble6(nlo.identity_pose())
    Hr, gr, cr = oracle.ndt6_assemble(point, mean, S, np.eye(3), np.zeros(3), 0, None, long_double=True)
    np.testing.assert_allclose(H, Hr, atol=1e-9); np.testing.assert_allclose(c, cr, rtol=1e-9)
    prob.close()
