"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  See oracle/nlo_oracle.h for the array conventions and the reference
file:line each function follows.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnlo_oracle.so")

LOSS_NONE, LOSS_EXPONENTIAL, LOSS_HUBER, LOSS_CAUCHY = 0, 1, 2, 3
TRACE6 = 36
TRACE3 = 17

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_float_p = ctypes.POINTER(ctypes.c_float)
_c_int_p = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    """Compile oracle/libnlo_oracle.so with the Makefile committed beside it."""
    srcs = [os.path.join(_HERE, f) for f in ("nlo_oracle.cc", "nlo_oracle_simd.cc", "nlo_oracle.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "libnlo_oracle.so"], check=True,
                   stdout=subprocess.DEVNULL)
    return _LIB_PATH


def use_native_build():
    """Rebuild the library with -march=native on THIS machine (the reference's own flags,
    CMakeLists.txt:10-13) and load that copy.  For the CPU timing baseline only: the parity tests
    keep the portable, FMA-free build.  Returns True on success."""
    global _lib
    native = os.path.join(_HERE, "_native", "libnlo_oracle.so")
    try:
        if os.path.exists(native):
            os.remove(native)   # never load a copy that was built for another machine's -march
        subprocess.run(["make", "-C", _HERE, "native"], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
        candidate = ctypes.CDLL(native)
    except Exception:
        return False
    candidate.nlo_oracle_voxel_key.restype = ctypes.c_uint64
    candidate.nlo_oracle_pnp_reference_points.restype = ctypes.c_int64
    candidate.nlo_oracle_room_points.restype = ctypes.c_int64
    _lib = candidate
    return True


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.nlo_oracle_voxel_key.restype = ctypes.c_uint64
        _lib.nlo_oracle_pnp_reference_points.restype = ctypes.c_int64
        _lib.nlo_oracle_room_points.restype = ctypes.c_int64
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_c_double_p)


def _params(p):
    arr = np.zeros(2, dtype=np.float64)
    p = list(p or [])
    arr[:len(p)] = p
    return arr


def loss(kind, params, s):
    out = np.zeros(3)
    pa = _params(params)
    lib().nlo_oracle_loss(ctypes.c_int(kind), pa.ctypes.data_as(_c_double_p), ctypes.c_double(s),
                          out.ctypes.data_as(_c_double_p))
    return out


def rotmat_to_quat(R):
    R, Rp = _d(R)
    q = np.zeros(4)
    lib().nlo_oracle_rotmat_to_quat(Rp, q.ctypes.data_as(_c_double_p))
    return q


def quat_to_rotmat(q):
    q, qp = _d(q)
    R = np.zeros(9)
    lib().nlo_oracle_quat_to_rotmat(qp, R.ctypes.data_as(_c_double_p))
    return R.reshape(3, 3)


def pose_from_Rt(R, t):
    """4x4 column-major flat pose[16] (Eigen::Isometry3d memory) from row-major R, t."""
    T = np.eye(4)
    T[:3, :3] = np.asarray(R, dtype=np.float64).reshape(3, 3)
    T[:3, 3] = t
    return np.ascontiguousarray(T.T).reshape(16).copy()


def pose_to_Rt(pose16):
    T = np.asarray(pose16, dtype=np.float64).reshape(4, 4).T
    return T[:3, :3].copy(), T[:3, 3].copy()


def ndt6_jacobian_residual(R, t, p, mean, S):
    R, Rp = _d(R); t, tp = _d(t); p, pp = _d(p); mean, mp = _d(mean); S, Sp = _d(S)
    J = np.zeros(18); r = np.zeros(3)
    lib().nlo_oracle_ndt6_jacobian_residual(Rp, tp, pp, mp, Sp, J.ctypes.data_as(_c_double_p),
                                            r.ctypes.data_as(_c_double_p))
    return J.reshape(3, 6), r


def ndt3_jacobian_residual(R2, t2, p, mean, S):
    R2, Rp = _d(R2); t2, tp = _d(t2); p, pp = _d(p); mean, mp = _d(mean); S, Sp = _d(S)
    J = np.zeros(9); r = np.zeros(3)
    lib().nlo_oracle_ndt3_jacobian_residual(Rp, tp, pp, mp, Sp, J.ctypes.data_as(_c_double_p),
                                            r.ctypes.data_as(_c_double_p))
    return J.reshape(3, 3), r


def reproj_jacobian_residual(R, t, X, px, intrinsics):
    R, Rp = _d(R); t, tp = _d(t); X, Xp = _d(X); px, pxp = _d(px); K, Kp = _d(intrinsics)
    J = np.zeros(12); r = np.zeros(2)
    lib().nlo_oracle_reproj_jacobian_residual(Rp, tp, Xp, pxp, Kp, J.ctypes.data_as(_c_double_p),
                                              r.ctypes.data_as(_c_double_p))
    return J.reshape(2, 6), r


def ndt6_assemble(point, mean, sqrt_info, R, t, loss_kind=LOSS_NONE, loss_params=None,
                  begin=0, end=None, long_double=False):
    point, pp = _d(point); mean, mp = _d(mean); sqrt_info, sp = _d(sqrt_info)
    R, Rp = _d(R); t, tp = _d(t)
    n = point.size // 3
    end = n if end is None else end
    H = np.zeros(21); g = np.zeros(6); cost = ctypes.c_double(0)
    pa = _params(loss_params)
    lib().nlo_oracle_ndt6_assemble(ctypes.c_int64(begin), ctypes.c_int64(end), pp, mp, sp, Rp, tp,
                                   ctypes.c_int(loss_kind), pa.ctypes.data_as(_c_double_p),
                                   ctypes.c_int(int(long_double)), H.ctypes.data_as(_c_double_p),
                                   g.ctypes.data_as(_c_double_p), ctypes.byref(cost))
    return H, g, cost.value


def ndt3_assemble(point, mean, sqrt_info, R2, t2, loss_kind=LOSS_NONE, loss_params=None,
                  begin=0, end=None, long_double=False):
    point, pp = _d(point); mean, mp = _d(mean); sqrt_info, sp = _d(sqrt_info)
    R2, Rp = _d(R2); t2, tp = _d(t2)
    n = point.size // 3
    end = (n // 4) * 4 if end is None else end
    H = np.zeros(6); g = np.zeros(3); cost = ctypes.c_double(0)
    pa = _params(loss_params)
    lib().nlo_oracle_ndt3_assemble(ctypes.c_int64(begin), ctypes.c_int64(end), pp, mp, sp, Rp, tp,
                                   ctypes.c_int(loss_kind), pa.ctypes.data_as(_c_double_p),
                                   ctypes.c_int(int(long_double)), H.ctypes.data_as(_c_double_p),
                                   g.ctypes.data_as(_c_double_p), ctypes.byref(cost))
    return H, g, cost.value


def reproj_assemble(local_point, pixel, intrinsics, R, t, loss_kind=LOSS_NONE, loss_params=None,
                    begin=0, end=None, long_double=False):
    X, Xp = _d(local_point); px, pxp = _d(pixel); K, Kp = _d(intrinsics)
    R, Rp = _d(R); t, tp = _d(t)
    n = X.size // 3
    end = n if end is None else end
    H = np.zeros(21); g = np.zeros(6); cost = ctypes.c_double(0)
    pa = _params(loss_params)
    lib().nlo_oracle_reproj_assemble(ctypes.c_int64(begin), ctypes.c_int64(end), Xp, pxp, Kp, Rp,
                                     tp, ctypes.c_int(loss_kind), pa.ctypes.data_as(_c_double_p),
                                     ctypes.c_int(int(long_double)),
                                     H.ctypes.data_as(_c_double_p), g.ctypes.data_as(_c_double_p),
                                     ctypes.byref(cost))
    return H, g, cost.value


def _solve_common(fn, head_args, loss_kind, loss_params, max_iterations, ptol, gtol, extra_args,
                  pose16, trace_width):
    pose = np.array(pose16, dtype=np.float64).reshape(16).copy()
    iters = ctypes.c_int(0); final_cost = ctypes.c_double(0)
    trace = np.zeros((max(max_iterations, 1), trace_width))
    pa = _params(loss_params)
    fn(*head_args, ctypes.c_int(loss_kind), pa.ctypes.data_as(_c_double_p),
       ctypes.c_int(max_iterations), ctypes.c_double(ptol), ctypes.c_double(gtol), *extra_args,
       pose.ctypes.data_as(_c_double_p), ctypes.byref(iters), ctypes.byref(final_cost),
       trace.ctypes.data_as(_c_double_p))
    rows = min(iters.value + 1, max_iterations)
    return pose, iters.value, final_cost.value, trace[:rows].copy()


def ndt6_solve(point, mean, sqrt_info, pose16, loss_kind=LOSS_NONE, loss_params=None,
               max_iterations=40, parameter_tolerance=1e-6, gradient_tolerance=1e-6,
               num_threads=0):
    point, pp = _d(point); mean, mp = _d(mean); sqrt_info, sp = _d(sqrt_info)
    n = point.size // 3
    return _solve_common(lib().nlo_oracle_ndt6_solve, (ctypes.c_int64(n), pp, mp, sp), loss_kind,
                         loss_params, max_iterations, parameter_tolerance, gradient_tolerance,
                         (ctypes.c_int(num_threads),), pose16, TRACE6)


def ndt3_solve(point, mean, sqrt_info, pose16, loss_kind=LOSS_NONE, loss_params=None,
               max_iterations=40, parameter_tolerance=1e-6, gradient_tolerance=1e-6):
    point, pp = _d(point); mean, mp = _d(mean); sqrt_info, sp = _d(sqrt_info)
    n = point.size // 3
    return _solve_common(lib().nlo_oracle_ndt3_solve, (ctypes.c_int64(n), pp, mp, sp), loss_kind,
                         loss_params, max_iterations, parameter_tolerance, gradient_tolerance, (),
                         pose16, TRACE3)


def reproj_solve(local_point, pixel, intrinsics, pose16, loss_kind=LOSS_NONE, loss_params=None,
                 max_iterations=40, parameter_tolerance=1e-6, gradient_tolerance=1e-6):
    X, Xp = _d(local_point); px, pxp = _d(pixel); K, Kp = _d(intrinsics)
    n = X.size // 3
    return _solve_common(lib().nlo_oracle_reproj_solve, (ctypes.c_int64(n), Xp, pxp, Kp),
                         loss_kind, loss_params, max_iterations, parameter_tolerance,
                         gradient_tolerance, (), pose16, TRACE6)


def gn6_step(H21, g, cost, state6, parameter_tolerance=1e-6, gradient_tolerance=1e-6):
    H21, Hp = _d(H21); g, gp = _d(g)
    st = np.array(state6, dtype=np.float64).reshape(9).copy()
    conv = lib().nlo_oracle_gn6_step(Hp, gp, ctypes.c_double(cost),
                                     ctypes.c_double(parameter_tolerance),
                                     ctypes.c_double(gradient_tolerance),
                                     st.ctypes.data_as(_c_double_p))
    return int(conv), st


def gn3_step(H6, g, cost, state3, parameter_tolerance=1e-6, gradient_tolerance=1e-6):
    H6, Hp = _d(H6); g, gp = _d(g)
    st = np.array(state3, dtype=np.float64).reshape(8).copy()
    conv = lib().nlo_oracle_gn3_step(Hp, gp, ctypes.c_double(cost),
                                     ctypes.c_double(parameter_tolerance),
                                     ctypes.c_double(gradient_tolerance),
                                     st.ctypes.data_as(_c_double_p))
    return int(conv), st


def pnp_reference_points():
    n = lib().nlo_oracle_pnp_reference_points(None, ctypes.c_int64(0))
    xyz = np.zeros((n, 3))
    lib().nlo_oracle_pnp_reference_points(xyz.ctypes.data_as(_c_double_p), ctypes.c_int64(n))
    return xyz


def room_points():
    n = lib().nlo_oracle_room_points(None, ctypes.c_int64(0))
    xyz = np.zeros((n, 3))
    lib().nlo_oracle_room_points(xyz.ctypes.data_as(_c_double_p), ctypes.c_int64(n))
    return xyz


def voxel_key(point, inverse_voxel_resolution):
    p, pp = _d(point)
    return int(lib().nlo_oracle_voxel_key(pp, ctypes.c_double(inverse_voxel_resolution)))


def simd_pack(point, mean, sqrt_info):
    point, pp = _d(point); mean, mp = _d(mean); sqrt_info, sp = _d(sqrt_info)
    n = point.size // 3
    planes = np.zeros(15 * n, dtype=np.float32)
    lib().nlo_oracle_simd_pack(ctypes.c_int64(n), pp, mp, sp, planes.ctypes.data_as(_c_float_p))
    return planes


def simd_ndt6_assemble(planes, n, R, t, loss_kind=LOSS_NONE, loss_params=None, num_threads=1):
    R, Rp = _d(R); t, tp = _d(t)
    H = np.zeros(21); g = np.zeros(6); cost = ctypes.c_double(0)
    pa = _params(loss_params)
    lib().nlo_oracle_simd_ndt6_assemble(ctypes.c_int64(n), planes.ctypes.data_as(_c_float_p), Rp,
                                        tp, ctypes.c_int(loss_kind),
                                        pa.ctypes.data_as(_c_double_p), ctypes.c_int(num_threads),
                                        H.ctypes.data_as(_c_double_p),
                                        g.ctypes.data_as(_c_double_p), ctypes.byref(cost))
    return H, g, cost.value


def ndt6_assemble_threads(point, mean, sqrt_info, R, t, loss_kind=LOSS_NONE, loss_params=None,
                          num_threads=1):
    point, pp = _d(point); mean, mp = _d(mean); sqrt_info, sp = _d(sqrt_info)
    R, Rp = _d(R); t, tp = _d(t)
    n = point.size // 3
    H = np.zeros(21); g = np.zeros(6); cost = ctypes.c_double(0)
    pa = _params(loss_params)
    lib().nlo_oracle_ndt6_assemble_threads(ctypes.c_int64(n), pp, mp, sp, Rp, tp,
                                           ctypes.c_int(loss_kind), pa.ctypes.data_as(_c_double_p),
                                           ctypes.c_int(num_threads),
                                           H.ctypes.data_as(_c_double_p),
                                           g.ctypes.data_as(_c_double_p), ctypes.byref(cost))
    return H, g, cost.value


def simd_ndt3_assemble(planes, n, R2, t2, loss_kind=LOSS_NONE, loss_params=None, num_threads=1):
    """Float 8-lane twin of the planar minimizer (..._analytic_3dof_simd.cc:85-157)."""
    R2, Rp = _d(R2); t2, tp = _d(t2)
    H = np.zeros(6); g = np.zeros(3); cost = ctypes.c_double(0)
    pa = _params(loss_params)
    lib().nlo_oracle_simd_ndt3_assemble(ctypes.c_int64(n), planes.ctypes.data_as(_c_float_p), Rp, tp,
                                        ctypes.c_int(loss_kind), pa.ctypes.data_as(_c_double_p),
                                        ctypes.c_int(num_threads), H.ctypes.data_as(_c_double_p),
                                        g.ctypes.data_as(_c_double_p), ctypes.byref(cost))
    return H, g, cost.value


def simd_reproj_pack(local_point, pixel):
    X, Xp = _d(local_point); px, pxp = _d(pixel)
    n = X.size // 3
    planes = np.zeros(5 * n, dtype=np.float32)
    lib().nlo_oracle_simd_reproj_pack(ctypes.c_int64(n), Xp, pxp, planes.ctypes.data_as(_c_float_p))
    return planes


def simd_reproj_assemble(planes, n, intrinsics, R, t, loss_kind=LOSS_NONE, loss_params=None, num_threads=1):
    """Float 8-lane twin of the reprojection minimizer (..._analytic_simd.cc:55-137), quirks kept."""
    K, Kp = _d(intrinsics); R, Rp = _d(R); t, tp = _d(t)
    H = np.zeros(21); g = np.zeros(6); cost = ctypes.c_double(0)
    pa = _params(loss_params)
    lib().nlo_oracle_simd_reproj_assemble(ctypes.c_int64(n), planes.ctypes.data_as(_c_float_p), Kp, Rp, tp,
                                          ctypes.c_int(loss_kind), pa.ctypes.data_as(_c_double_p),
                                          ctypes.c_int(num_threads), H.ctypes.data_as(_c_double_p),
                                          g.ctypes.data_as(_c_double_p), ctypes.byref(cost))
    return H, g, cost.value
