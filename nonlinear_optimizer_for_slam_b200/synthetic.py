"""Deterministic synthetic inputs for the BASELINE.json configs (numpy; host side only).

The room, the NDT map rules and the PnP fixture restate the reference's test mains (paths relative
to /root/reference/nonlinear_optimizer/):
  mahalanobis_distance_minimizer/tests/simple_optimization_test.cc:170-204  GenerateGlobalPoints
  .../simple_optimization_test.cc:236-280                                   UpdateNdtMap
  reprojection_error_minimizer/tests/simple_optimization_test.cc:42-71,115-158  PnP fixture
The association rule used for the large configs (cell of the dense voxel grid containing the
point under the initial pose) is SURVEY.md section 8(d)'s; the reference's own KD-tree matcher is
outside the hot path.
"""
import numpy as np


def yaw_pose(t, yaw):
    """4x4 row-major homogeneous matrix: rotation `yaw` about z, translation t."""
    c, s = np.cos(yaw), np.sin(yaw)
    T = np.eye(4)
    T[:3, :3] = [[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]]
    T[:3, 3] = t
    return T


def to_pose16(T):
    """row-major 4x4 -> column-major flat pose[16] (the C ABI convention)."""
    return np.ascontiguousarray(np.asarray(T, dtype=np.float64).T).reshape(16).copy()


def _accumulated_axis(start, stop, step):
    """Values of `for (x = start; x <= stop; x += step)` with the same accumulated rounding."""
    n = int(np.ceil((stop - start) / step)) + 3
    seq = np.cumsum(np.concatenate([[start], np.full(n, step)]))
    return seq[seq <= stop]


def room_points():
    """GenerateGlobalPoints: 7 x 5 x 2.5 m room sampled at 0.01 m (954 605 points)."""
    width, length, height, step = 5.0, 7.0, 2.5, 0.01
    xs = _accumulated_axis(-length / 2.0, length / 2.0, step)
    ys = _accumulated_axis(-width / 2.0, width / 2.0, step)
    zs = _accumulated_axis(0.0, height, step)
    out = []
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    out.append(np.stack([X.ravel(), Y.ravel(), np.zeros(X.size)], 1))
    X, Z = np.meshgrid(xs, zs, indexing="ij")
    y = -width / 2.0
    lr = np.empty((X.size, 2, 3))
    lr[:, 0] = np.stack([X.ravel(), np.full(X.size, y), Z.ravel()], 1)
    lr[:, 1] = np.stack([X.ravel(), np.full(X.size, -y), Z.ravel()], 1)
    out.append(lr.reshape(-1, 3))
    Y, Z = np.meshgrid(ys, zs, indexing="ij")
    x = -length / 2.0
    fb = np.empty((Y.size, 2, 3))
    fb[:, 0] = np.stack([np.full(Y.size, -x), Y.ravel(), Z.ravel()], 1)
    fb[:, 1] = np.stack([np.full(Y.size, x), Y.ravel(), Z.ravel()], 1)
    out.append(fb.reshape(-1, 3))
    return np.concatenate(out, 0)


def room_surface_samples(n, rng, noise_sigma=0.0):
    """n points uniform (by area) on the floor and the four walls of the room, plus N(0, sigma)."""
    areas = np.array([35.0, 17.5, 17.5, 12.5, 12.5])
    which = rng.choice(5, size=n, p=areas / areas.sum())
    u = rng.random((n, 2))
    p = np.zeros((n, 3))
    f = which == 0
    p[f] = np.stack([-3.5 + 7.0 * u[f, 0], -2.5 + 5.0 * u[f, 1], np.zeros(f.sum())], 1)
    for k, y in ((1, -2.5), (2, 2.5)):
        m = which == k
        p[m] = np.stack([-3.5 + 7.0 * u[m, 0], np.full(m.sum(), y), 2.5 * u[m, 1]], 1)
    for k, x in ((3, -3.5), (4, 3.5)):
        m = which == k
        p[m] = np.stack([np.full(m.sum(), x), -2.5 + 5.0 * u[m, 0], 2.5 * u[m, 1]], 1)
    if noise_sigma > 0.0:
        p += rng.normal(0.0, noise_sigma, size=p.shape)
    return p


def build_ndt_grid(points, voxel, v_not_transposed=True):
    """UpdateNdtMap on a dense voxel grid.

    Rules restated from simple_optimization_test.cc:236-280: moment starts at Identity
    (types.h:14), a cell needs count >= 5, mean = sum / count, cov = moment / count - mean mean^T,
    reject if the largest eigenvalue < 0.01, clamp the two small eigenvalues to 0.01 * largest,
    sqrt_information = diag(eigval^-1/2) * V -- V, not V^T, as the reference writes it
    (`v_not_transposed`); information = S^T S.  (The reference `return`s out of the whole update
    on the first rejected cell, :263-266; here a rejected cell is just invalid.)
    Returns dict(origin, dims, voxel, mean[cells,3], sqrt_info[cells,9] row-major, valid[cells]).
    """
    inv = 1.0 / voxel
    key = np.floor(points * inv).astype(np.int64)
    kmin = key.min(0)
    dims = (key.max(0) - kmin + 1).astype(np.int64)
    k = key - kmin
    lin = (k[:, 2] * dims[1] + k[:, 1]) * dims[0] + k[:, 0]
    cells = int(dims.prod())
    count = np.bincount(lin, minlength=cells).astype(np.float64)
    s = np.stack([np.bincount(lin, weights=points[:, a], minlength=cells) for a in range(3)], 1)
    moment = np.zeros((cells, 3, 3))
    for a in range(3):
        for b in range(a, 3):
            m = np.bincount(lin, weights=points[:, a] * points[:, b], minlength=cells)
            moment[:, a, b] = m
            moment[:, b, a] = m
    moment += np.eye(3)
    valid = count >= 5
    safe = np.where(count > 0, count, 1.0)
    mean = s / safe[:, None]
    cov = moment / safe[:, None, None] - mean[:, :, None] * mean[:, None, :]
    cov[~valid] = np.eye(3)
    w, V = np.linalg.eigh(cov)
    valid &= w[:, 2] >= 0.01
    w = np.where(valid[:, None], w, 1.0)
    w[:, 0] = np.maximum(w[:, 0], 0.01 * w[:, 2])
    w[:, 1] = np.maximum(w[:, 1], 0.01 * w[:, 2])
    d = 1.0 / np.sqrt(w)
    Vm = V if v_not_transposed else np.transpose(V, (0, 2, 1))
    S = d[:, :, None] * Vm
    S[~valid] = 0.0
    mean[~valid] = 0.0
    return {
        "origin": kmin.astype(np.float64) * voxel,
        "dims": dims.astype(np.int32),
        "voxel": float(voxel),
        "mean": np.ascontiguousarray(mean),
        "sqrt_info": np.ascontiguousarray(S.reshape(cells, 9)),
        "valid": valid.astype(np.uint8),
    }


_GRID_CACHE = {}


def room_ndt_grid(voxel=0.5):
    if voxel not in _GRID_CACHE:
        _GRID_CACHE[voxel] = build_ndt_grid(room_points(), voxel)
    return _GRID_CACHE[voxel]


def associate_dense(local_points, init_T, grid, keep_unmatched=False):
    """One correspondence per point: the valid cell whose voxel contains init_T * point, else the
    nearest valid cell mean within 1.0 m (neighbour scan in z, y, x order, first strict minimum --
    the same rule as the device generator).  Unmatched points are dropped, or kept with S = 0
    (an exact zero contribution) when keep_unmatched is set."""
    w = local_points @ init_T[:3, :3].T + init_T[:3, 3]
    c = np.floor((w - grid["origin"]) / grid["voxel"]).astype(np.int64)
    dims = grid["dims"].astype(np.int64)
    valid = grid["valid"] != 0

    def lin_of(cc):
        inside = np.all((cc >= 0) & (cc < dims), axis=1)
        lin = (cc[:, 2] * dims[1] + cc[:, 1]) * dims[0] + cc[:, 0]
        return np.where(inside, lin, 0), inside

    lin, inside = lin_of(c)
    cell = np.where(inside & valid[lin], lin, -1)
    need = np.nonzero(cell < 0)[0]
    if need.size:
        reach = min(4, int(np.ceil(1.0 / grid["voxel"])))
        best = np.full(need.size, 1.0)
        best_cell = np.full(need.size, -1, dtype=np.int64)
        wn = w[need]
        cn = c[need]
        for oz in range(-reach, reach + 1):
            for oy in range(-reach, reach + 1):
                for ox in range(-reach, reach + 1):
                    l2, in2 = lin_of(cn + np.array([ox, oy, oz]))
                    ok = in2 & valid[l2]
                    e = wn - grid["mean"][l2]
                    d2 = (e[:, 0] * e[:, 0] + e[:, 1] * e[:, 1]) + e[:, 2] * e[:, 2]
                    better = ok & (d2 < best)
                    best = np.where(better, d2, best)
                    best_cell = np.where(better, l2, best_cell)
        cell[need] = best_cell
    ok = cell >= 0
    if keep_unmatched:
        safe = np.where(ok, cell, 0)
        mean = np.where(ok[:, None], grid["mean"][safe], 0.0)
        S = np.where(ok[:, None], grid["sqrt_info"][safe], 0.0)
        return (np.ascontiguousarray(local_points), np.ascontiguousarray(mean),
                np.ascontiguousarray(S))
    cell = cell[ok]
    return (np.ascontiguousarray(local_points[ok]), np.ascontiguousarray(grid["mean"][cell]),
            np.ascontiguousarray(grid["sqrt_info"][cell]))


def ndt_problem(n, seed, true_T, init_T=None, voxel=0.5, noise_sigma=0.01):
    """cfg1 / cfg2 style registration problem: (point[n',3], mean[n',3], sqrt_info[n',9])."""
    rng = np.random.default_rng(seed)
    world = room_surface_samples(n, rng, noise_sigma)
    Tinv = np.linalg.inv(true_T)
    local = world @ Tinv[:3, :3].T + Tinv[:3, 3]
    init_T = np.eye(4) if init_T is None else init_T
    return associate_dense(local, init_T, room_ndt_grid(voxel))


CFG1_TRUE = yaw_pose([-0.2, 0.123, 0.3], 0.1)     # simple_optimization_test.cc:85-88
CFG2_TRUE = yaw_pose([-0.15, 0.05, 0.0], 0.2)     # 3dof_6dof_comparison_test.cc:77-80
PNP_TRUE = yaw_pose([-0.1, 0.123, -0.5], 0.1)     # reprojection tests/simple_optimization_test.cc:58-61
PNP_INTRINSICS = np.array([525.0, 525.0, 320.0, 240.0, 1.0 / 525.0, 1.0 / 525.0])


def pnp_fixture():
    """The reference's deterministic 630-point PnP fixture: (X[630,3], pixel[630,2], K[6])."""
    xs = _accumulated_axis(-1.5, 1.5, 0.1)
    ys = _accumulated_axis(-1.0, 1.0, 0.1)
    Xg, Yg = np.meshgrid(xs, ys, indexing="ij")
    X = np.stack([Xg.ravel(), Yg.ravel(), np.full(Xg.size, 3.0)], 1)
    return X, project(X, PNP_TRUE), PNP_INTRINSICS.copy()


def project(X, true_T, K=PNP_INTRINSICS):
    Tinv = np.linalg.inv(true_T)
    Q = X @ Tinv[:3, :3].T + Tinv[:3, 3]
    inv_z = 1.0 / Q[:, 2]
    return np.stack([K[0] * Q[:, 0] * inv_z + K[2], K[1] * Q[:, 1] * inv_z + K[3]], 1)


def pnp_problem(n, seed, true_T=PNP_TRUE, pixel_sigma=0.5, outlier_fraction=0.05):
    """cfg3: X ~ U([-1.5,1.5] x [-1,1] x [2,4]), noisy pixels, gross outliers uniform in the image."""
    rng = np.random.default_rng(seed)
    X = np.stack([rng.uniform(-1.5, 1.5, n), rng.uniform(-1.0, 1.0, n), rng.uniform(2.0, 4.0, n)], 1)
    px = project(X, true_T) + rng.normal(0.0, pixel_sigma, size=(n, 2))
    out = rng.random(n) < outlier_fraction
    px[out] = np.stack([rng.uniform(0, 640, out.sum()), rng.uniform(0, 480, out.sum())], 1)
    return X, px, PNP_INTRINSICS.copy()


def random_ndt_records(n, seed, scale=3.0):
    """Unstructured random correspondences (dense random S) for kernel-vs-oracle parity tests."""
    rng = np.random.default_rng(seed)
    point = rng.uniform(-scale, scale, (n, 3))
    mean = point + rng.normal(0.0, 0.3, (n, 3))
    S = rng.normal(0.0, 1.0, (n, 9)) * rng.uniform(0.5, 10.0, (n, 1))
    return point, mean, S


def random_rotation(rng, max_angle=0.5):
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = rng.uniform(-max_angle, max_angle)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def information6(sqrt_info):
    """Unique entries (00 01 02 11 12 22) of S^T S for row-major S[n,9] -- what the device stores."""
    S = np.asarray(sqrt_info, dtype=np.float64).reshape(-1, 3, 3)
    L = np.einsum("nki,nkj->nij", S, S)
    return np.stack([L[:, 0, 0], L[:, 0, 1], L[:, 0, 2], L[:, 1, 1], L[:, 1, 2], L[:, 2, 2]], 1)


def sqrt_info_from_information6(info6):
    """A sqrt_information (row-major [n,9]) with S^T S equal to the given information matrices:
    the upper Cholesky factor (rank-deficient / zero records map to zero rows)."""
    L6 = np.asarray(info6, dtype=np.float64).reshape(-1, 6)
    n = len(L6)
    L = np.zeros((n, 3, 3))
    L[:, 0, 0] = L6[:, 0]; L[:, 0, 1] = L[:, 1, 0] = L6[:, 1]; L[:, 0, 2] = L[:, 2, 0] = L6[:, 2]
    L[:, 1, 1] = L6[:, 3]; L[:, 1, 2] = L[:, 2, 1] = L6[:, 4]; L[:, 2, 2] = L6[:, 5]
    S = np.zeros((n, 3, 3))
    ok = np.linalg.det(L) > 1e-300
    if ok.any():
        S[ok] = np.transpose(np.linalg.cholesky(L[ok]), (0, 2, 1))
    return S.reshape(n, 9)
