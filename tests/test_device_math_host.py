"""CPU: the kernels' own math source (csrc/nlo_device.cuh: Ndt6Point / Ndt3Point / ReprojPoint,
EvaluateLoss, the CanonPlan rotation, Step6 / Step3) compiled for the HOST and compared with the
oracle -- the parity bar of the GPU tests (1e-6; observed ~1e-15), checked without a GPU on the very
file the kernels are built from.  The only edit made to the header on the way (by this test, into a
temporary copy): the two inline-PTX seed instructions of FastRcp / FastSqrt become a float
reciprocal / reciprocal square root; the Newton steps behind them are the product's."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from nonlinear_optimizer_for_slam_b200 import synthetic as syn
from parity import TOL, assert_sums_close, rotation_angle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "nonlinear_optimizer_for_slam_b200", "csrc")
SEEDS = {
    'asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));': "r = static_cast<double>(1.0f / static_cast<float>(x));",
    'asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));':
        "r = static_cast<double>(1.0f / sqrtf(static_cast<float>(x)));",
}
DP = ctypes.POINTER(ctypes.c_double)
LOSSES = [(0, None), (1, [1.0, 1.0]), (2, [1.0]), (3, [0.5])]


def _cuda_include():
    for cand in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if cand and os.path.exists(os.path.join(cand, "include", "cuda_runtime.h")):
            return os.path.join(cand, "include")
    pytest.skip("CUDA headers not found")


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("device_math")
    text = open(os.path.join(CSRC, "nlo_device.cuh")).read()
    for old, new in SEEDS.items():
        assert text.count(old) == 1, old
        text = text.replace(old, new)
    assert "asm(" not in text and "asm volatile" not in text  # no other inline PTX in the math header
    (tmp / "nlo_device_host.cuh").write_text(text)
    lib = str(tmp / "libdevice_math_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-Wno-attributes",
                    "-Wno-unknown-pragmas", "-I", str(tmp), "-I", CSRC, "-I", _cuda_include(),
                    os.path.join(ROOT, "tests", "device_math_host.cc"), "-o", lib], check=True)
    h = ctypes.CDLL(lib)
    h.hm_assemble.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, DP, DP, DP, DP, ctypes.c_double,
                              ctypes.c_double, DP, DP]
    h.hm_assemble.restype = None
    h.hm_solve.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, DP, DP, DP, DP, ctypes.c_double,
                           ctypes.c_double, DP, ctypes.c_int, ctypes.c_double, ctypes.c_double, DP]
    h.hm_solve.restype = ctypes.c_int
    return h


def _p(a):
    return None if a is None else a.ctypes.data_as(DP)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _assemble(hm, kind, loss, a, b, S, pose16, K=None):
    a, b = _c(a), _c(b)
    S = None if S is None else _c(S)
    K = _c(K if K is not None else np.zeros(6))
    pose = _c(pose16).reshape(16)
    params = list(loss[1] or []) + [0.0, 0.0]
    out = np.zeros(28)
    hm.hm_assemble(kind, loss[0], a.size // 3, _p(a), _p(b), _p(S), _p(pose), params[0], params[1], _p(K), _p(out))
    return (out[:6], out[6:9], out[9]) if kind == 1 else (out[:21], out[21:27], out[27])


def _solve(hm, kind, loss, a, b, S, pose16, K=None, max_iterations=40, ptol=1e-6, gtol=1e-6):
    a, b = _c(a), _c(b)
    S = None if S is None else _c(S)
    K = _c(K if K is not None else np.zeros(6))
    pose = _c(pose16).reshape(16).copy()
    params = list(loss[1] or []) + [0.0, 0.0]
    cost = ctypes.c_double(0)
    it = hm.hm_solve(kind, loss[0], a.size // 3, _p(a), _p(b), _p(S), _p(pose), params[0], params[1], _p(K),
                     max_iterations, ptol, gtol, ctypes.byref(cost))
    return pose, it, cost.value


def _rand_pose(rng, oracle):
    R = syn.random_rotation(rng)
    t = rng.uniform(-0.5, 0.5, 3)
    Rq = oracle.quat_to_rotmat(oracle.rotmat_to_quat(R))  # the reference goes through a quaternion first
    return oracle.pose_from_Rt(R, t), Rq, t


@pytest.mark.parametrize("loss", LOSSES)
def test_ndt6_point_math_and_canonical_rotation(hm, oracle, loss):
    rng = np.random.default_rng(11)
    point, mean, S = syn.random_ndt_records(5000, seed=5)
    pose16, Rq, t = _rand_pose(rng, oracle)
    H, g, c = _assemble(hm, 0, loss, point, mean, S, pose16)
    Hr, gr, cr = oracle.ndt6_assemble(point, mean, S, Rq, t, loss[0], loss[1], long_double=True)
    assert_sums_close(H, g, c, Hr, gr, cr, tol=1e-11)


@pytest.mark.parametrize("loss", LOSSES)
def test_ndt3_point_math(hm, oracle, loss):
    rng = np.random.default_rng(12)
    point, mean, S = syn.random_ndt_records(4096, seed=6)
    T = syn.yaw_pose([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), 0.0], rng.uniform(-0.4, 0.4))
    H, g, c = _assemble(hm, 1, loss, point, mean, S, syn.to_pose16(T))
    Hr, gr, cr = oracle.ndt3_assemble(point, mean, S, T[:2, :2], T[:2, 3], loss[0], loss[1], long_double=True)
    assert_sums_close(H, g, c, Hr, gr, cr, tol=1e-11)


@pytest.mark.parametrize("loss", [(0, None), (1, [1.0, 1.0]), (2, [0.01]), (3, [0.01])])
def test_reprojection_point_math_and_depth_gate(hm, oracle, loss):
    rng = np.random.default_rng(13)
    X, px, K = syn.pnp_problem(4097, seed=7)
    X[::7, 2] = -1.0  # behind the camera: the gate contributes exactly zero
    pose16, Rq, t = _rand_pose(rng, oracle)
    H, g, c = _assemble(hm, 2, loss, X, px, None, pose16, K)
    Hr, gr, cr = oracle.reproj_assemble(X, px, K, Rq, t, loss[0], loss[1], long_double=True)
    assert_sums_close(H, g, c, Hr, gr, cr, tol=1e-10)


def _check_pose(oracle, pose, pose_r, three_dof=False):
    Ra, ta = oracle.pose_to_Rt(pose)
    Rb, tb = oracle.pose_to_Rt(pose_r)
    assert np.max(np.abs(ta - tb)) < TOL
    assert rotation_angle(Ra, Rb) < TOL


def test_whole_solves_follow_the_oracle(hm, oracle):
    """Step6 / Step3 (LDL^T, quaternion update, lambda schedule, convergence tests) through whole solves."""
    point, mean, S = syn.ndt_problem(3000, 1001, syn.CFG1_TRUE)
    pose0 = oracle.pose_from_Rt(np.eye(3), np.zeros(3))
    for loss in ((1, [1.0, 1.0]), (2, [1.0]), (0, None)):
        pose, it, cost = _solve(hm, 0, loss, point, mean, S, pose0)
        pose_r, it_r, cost_r, _ = oracle.ndt6_solve(point, mean, S, pose0, loss[0], loss[1])
        assert it == it_r, (loss, it, it_r)
        _check_pose(oracle, pose, pose_r)
        assert abs(cost - cost_r) <= TOL * abs(cost_r)
    point, mean, S = syn.ndt_problem(2000, 1002, syn.CFG2_TRUE)
    n4 = (len(point) // 4) * 4  # ..._analytic_3dof.cc:33-36
    init = syn.to_pose16(syn.yaw_pose([0.02, -0.01, 0.3], 0.03))
    pose, it, cost = _solve(hm, 1, (2, [1.0]), point[:n4], mean[:n4], S[:n4], init)
    pose_r, it_r, cost_r, _ = oracle.ndt3_solve(point, mean, S, init, 2, [1.0])
    assert it == it_r
    _check_pose(oracle, pose, pose_r)
    X, px, K = syn.pnp_fixture()
    pose, it, cost = _solve(hm, 2, (0, None), X, px, None, pose0, K)
    pose_r, it_r, cost_r, _ = oracle.reproj_solve(X, px, K, pose0, 0, None)
    assert it == it_r == 6          # results/reproj_amd64.txt:5  "COST: 2.33228e-11, iter: 6"
    _check_pose(oracle, pose, pose_r)
