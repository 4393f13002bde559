"""GPU: the chunked, pipelined host -> device ingest (csrc/nlo_ingest.cu).  Whatever the source
(pageable or pinned arrays, the reference's AoS records, several chunks or one, a batched
concatenation), the device must end up with exactly the same planes."""
import numpy as np
import pytest

from nonlinear_optimizer_for_slam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _aos(point, mean, S, stride=304):
    n = len(point)
    off_mean, off_sqrt = 24 + 8 + 24 + 72, 24 + 8 + 24 + 72 + 24 + 72
    rec = np.zeros((n, stride), dtype=np.uint8)
    rec[:, 0:24] = point.view(np.uint8).reshape(n, 24)
    rec[:, off_mean:off_mean + 24] = mean.view(np.uint8).reshape(n, 24)
    S_col = np.ascontiguousarray(S.reshape(n, 3, 3).transpose(0, 2, 1)).reshape(n, 9)
    rec[:, off_sqrt:off_sqrt + 72] = S_col.view(np.uint8).reshape(n, 72)
    return rec, off_mean, off_sqrt


@pytest.mark.parametrize("n", [1, 8191, 8193, 70001, 1_500_001])
def test_every_source_gives_the_same_planes(ctx, nlo, n):
    point, mean, S = syn.random_ndt_records(n, seed=n % 1000)
    a = nlo.NdtProblem(ctx, capacity=n)
    a.upload(point, mean, S)                                   # pageable arrays -> pinned ring
    ref = a.download(0, n)
    np.testing.assert_array_equal(ref[0], point)
    np.testing.assert_array_equal(ref[1], mean)
    np.testing.assert_allclose(ref[2], syn.information6(S), rtol=1e-13, atol=1e-13 * np.abs(S).max() ** 2)
    # pinned arrays: slices go straight to the device stage
    arr, handle = nlo.host_alloc(n * 120)
    f = arr.view(np.float64)
    hp, hm, hs = f[:3 * n], f[3 * n:6 * n], f[6 * n:15 * n]
    hp[:] = point.ravel(); hm[:] = mean.ravel(); hs[:] = S.ravel()
    b = nlo.NdtProblem(ctx, capacity=n)
    b.upload_ptr(n, hp.ctypes.data, hm.ctypes.data, hs.ctypes.data)
    got = b.download(0, n)
    for x, y in zip(ref, got):
        np.testing.assert_array_equal(x, y)
    nlo.host_free(handle)
    # the reference's 304-byte records
    rec, off_mean, off_sqrt = _aos(point, mean, S)
    c = nlo.NdtProblem(ctx, capacity=n)
    c.upload_aos(rec, n, 304, 0, off_mean, off_sqrt, True)
    got = c.download(0, n)
    for x, y in zip(ref, got):
        np.testing.assert_array_equal(x, y)
    total_ms, gather_ms = ctx.ingest_stats()
    assert total_ms > 0.0 and 0.0 <= gather_ms <= total_ms
    # a second, shorter upload into the same problem replaces the first
    if n > 10:
        c.upload_aos(rec[: n // 2], n // 2, 304, 0, off_mean, off_sqrt, True)
        assert c.size == n // 2
        got = c.download(0, n // 2)
        np.testing.assert_array_equal(got[0], point[: n // 2])
    a.close(); b.close(); c.close()


def test_float_uploads_and_reprojection_records(ctx, nlo, oracle):
    n = 100003
    point, mean, S = syn.random_ndt_records(n, seed=4)
    p32, m32, s32 = point.astype(np.float32), mean.astype(np.float32), S.astype(np.float32)
    a = nlo.NdtProblem(ctx, capacity=n, storage="f32")
    a.upload_f32(p32, m32, s32)
    got = a.download(0, n)
    np.testing.assert_array_equal(got[0], p32.astype(np.float64))
    np.testing.assert_array_equal(got[1], m32.astype(np.float64))
    info = syn.information6(s32.astype(np.float64)).astype(np.float32).astype(np.float64)
    np.testing.assert_allclose(got[2], info, rtol=2e-7, atol=2e-7 * np.abs(info).max())
    a.close()
    # reprojection: SoA arrays and the reference's 40-byte records give the same sums
    X, px, K = syn.pnp_problem(70001, 1003)
    r1 = nlo.ReprojProblem(ctx, capacity=len(X)); r1.upload(X, px, K)
    rec = np.zeros((len(X), 5))
    rec[:, :3] = X; rec[:, 3:] = px
    r2 = nlo.ReprojProblem(ctx, capacity=len(X)); r2.upload_aos(rec.view(np.uint8), len(X), 40, 0, 24, K)
    ctx.set_loss(3, [1e-2])
    pose = syn.to_pose16(syn.yaw_pose([0.01, 0.02, -0.03], 0.01))
    H1, g1, c1 = r1.assemble(pose)
    H2, g2, c2 = r2.assemble(pose)
    assert np.array_equal(H1, H2) and np.array_equal(g1, g2) and c1 == c2
    r1.close(); r2.close()
