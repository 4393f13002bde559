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


# ---------------------------------------------------------------------------------------------
# Eigen's SelfAdjointEigenSolver<Matrix3d>::compute(), restated.
#
# The reference's NDT fixture stores sqrt_information = diag(eigval^-1/2) * V with V =
# eigsol.eigenvectors() -- V, not V^T (tests/simple_optimization_test.cc:275-276,
# tests/3dof_6dof_comparison_test.cc:246-247) -- so S^T S = V^T D^2 V depends on the SIGN of every
# eigenvector column, which no convention fixes: it is whatever the solver's arithmetic leaves.
# The published logs (results/maha_*.txt) can therefore only be reproduced with Eigen's own
# algorithm.  Eigen is a third-party dependency that is absent from /root/reference (found by
# find_package(Eigen3), CMakeLists.txt; 3.3.7 / 3.4.0 in the distributions of the time); this is
# its published algorithm (Eigen/src/Eigenvalues/SelfAdjointEigenSolver.h, Tridiagonalization.h,
# Eigen/src/Jacobi/Jacobi.h): scale to [-1, 1], the closed-form 3x3 Householder tridiagonalisation
# on the lower triangle, implicit symmetric QR steps with Wilkinson shift and Givens rotations
# accumulated on the right, selection sort into ascending eigenvalues.  `version` selects the
# deflation test, the only difference between 3.3 and 3.4 that matters here.
# ---------------------------------------------------------------------------------------------
def _eigen_make_givens(p, q):
    if q == 0.0:
        return (-1.0 if p < 0.0 else 1.0), 0.0
    if p == 0.0:
        return 0.0, (1.0 if q < 0.0 else -1.0)
    if abs(p) > abs(q):
        t = q / p
        u = np.sqrt(1.0 + t * t)
        if p < 0.0:
            u = -u
        c = 1.0 / u
        return c, -t * c
    t = p / q
    u = np.sqrt(1.0 + t * t)
    if q < 0.0:
        u = -u
    s = -1.0 / u
    return -t * s, s


def eigen_selfadjoint3(A, version="3.4"):
    """(eigenvalues ascending, eigenvectors as columns) of a symmetric 3x3, as Eigen computes them."""
    A = np.array(A, dtype=np.float64)
    m = np.tril(A)
    scale = np.abs(m).max()
    if scale == 0.0:
        scale = 1.0
    m = m / scale
    diag = np.zeros(3)
    sub = np.zeros(2)
    tol = np.finfo(np.float64).tiny
    diag[0] = m[0, 0]
    v1norm2 = m[2, 0] * m[2, 0]
    if v1norm2 <= tol:
        diag[1], diag[2] = m[1, 1], m[2, 2]
        sub[0], sub[1] = m[1, 0], m[2, 1]
        Q = np.eye(3)
    else:
        beta = np.sqrt(m[1, 0] * m[1, 0] + v1norm2)
        inv_beta = 1.0 / beta
        m01 = m[1, 0] * inv_beta
        m02 = m[2, 0] * inv_beta
        q = 2.0 * m01 * m[2, 1] + m02 * (m[2, 2] - m[1, 1])
        diag[1] = m[1, 1] + m02 * q
        diag[2] = m[2, 2] - m02 * q
        sub[0] = beta
        sub[1] = m[2, 1] - m01 * q
        Q = np.array([[1.0, 0.0, 0.0], [0.0, m01, m02], [0.0, m02, -m01]])
    n = 3
    end, start, it = n - 1, 0, 0
    eps = np.finfo(np.float64).eps
    while end > 0:
        for i in range(start, end):
            if version == "3.4":
                if abs(sub[i]) < tol:
                    sub[i] = 0.0
                else:
                    scaled = sub[i] / eps
                    if scaled * scaled <= abs(diag[i]) + abs(diag[i + 1]):
                        sub[i] = 0.0
            else:  # 3.3: isMuchSmallerThan(|e|, |d_i| + |d_i+1|, 2 eps) || |e| <= min
                if abs(sub[i]) <= (abs(diag[i]) + abs(diag[i + 1])) * 2.0 * eps or abs(sub[i]) <= tol:
                    sub[i] = 0.0
        while end > 0 and sub[end - 1] == 0.0:
            end -= 1
        if end <= 0:
            break
        it += 1
        if it > 30 * n:
            break
        start = end - 1
        while start > 0 and sub[start - 1] != 0.0:
            start -= 1
        # tridiagonal_qr_step
        td = (diag[end - 1] - diag[end]) * 0.5
        e = sub[end - 1]
        mu = diag[end]
        if td == 0.0:
            mu -= abs(e)
        elif e != 0.0:
            e2 = e * e
            h = np.hypot(td, e)
            if e2 == 0.0:
                mu -= e / ((td + (h if td > 0.0 else -h)) / e)
            else:
                mu -= e2 / (td + (h if td > 0.0 else -h))
        x = diag[start] - mu
        z = sub[start]
        k = start
        while k < end and z != 0.0:
            c, s = _eigen_make_givens(x, z)
            sdk = s * diag[k] + c * sub[k]
            dkp1 = s * sub[k] + c * diag[k + 1]
            diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1])
            diag[k + 1] = s * sdk + c * dkp1
            sub[k] = c * sdk - s * dkp1
            if k > start:
                sub[k - 1] = c * sub[k - 1] - s * z
            x = sub[k]
            if k < end - 1:
                z = -s * sub[k + 1]
                sub[k + 1] = c * sub[k + 1]
            # Q = Q * G: applyOnTheRight(k, k+1, rot) == rotation by rot.transpose() of the two columns
            xk = Q[:, k].copy()
            yk = Q[:, k + 1].copy()
            Q[:, k] = c * xk - s * yk
            Q[:, k + 1] = s * xk + c * yk
            k += 1
    # ascending selection sort, columns follow
    for i in range(n - 1):
        kmin = int(np.argmin(diag[i:]))
        if kmin > 0:
            diag[[i, i + kmin]] = diag[[i + kmin, i]]
            Q[:, [i, i + kmin]] = Q[:, [i + kmin, i]]
    return diag * scale, Q


def reference_ndt_grid(points, voxel, version="3.4"):
    """UpdateNdtMap of the reference's test mains (tests/simple_optimization_test.cc:236-280) on a
    dense grid, literally: raw moments accumulated in point order starting from moment = Identity
    (types.h:14), cov = moment / n - mean mean^T, Eigen's eigenvectors, S = diag(w^-1/2) * V.
    Same dictionary as synthetic.build_ndt_grid."""
    inv = 1.0 / voxel
    key = np.floor(points * inv).astype(np.int64)
    kmin = key.min(0)
    dims = (key.max(0) - kmin + 1).astype(np.int64)
    k = key - kmin
    lin = (k[:, 2] * dims[1] + k[:, 1]) * dims[0] + k[:, 0]
    cells = int(dims.prod())
    count = np.bincount(lin, minlength=cells)
    s = np.stack([np.bincount(lin, weights=points[:, a], minlength=cells) for a in range(3)], 1)
    moment = np.zeros((cells, 3, 3))
    for a in range(3):
        for b in range(3):
            moment[:, a, b] = np.bincount(lin, weights=points[:, a] * points[:, b], minlength=cells)
    mean = np.zeros((cells, 3))
    S = np.zeros((cells, 3, 3))
    valid = np.zeros(cells, dtype=np.uint8)
    for c in range(cells):
        if count[c] < 5:
            continue
        mu = s[c] / count[c]
        cov = (moment[c] + np.eye(3)) / count[c] - np.outer(mu, mu)
        w, V = eigen_selfadjoint3(cov, version)
        if w[2] < 0.01:
            continue
        w[0] = max(w[0], w[2] * 0.01)
        w[1] = max(w[1], w[2] * 0.01)
        mean[c] = mu
        S[c] = np.diag(1.0 / np.sqrt(w)) @ V
        valid[c] = 1
    return {"origin": kmin.astype(np.float64) * voxel, "dims": dims.astype(np.int32), "voxel": float(voxel),
            "mean": mean, "sqrt_info": np.ascontiguousarray(S.reshape(cells, 9)), "valid": valid}
