"""Thin numpy-facing wrapper over the C ABI (used by tests and bench.py).

The product is libnlo_cuda.so + the C++ drop-in headers under cxx/; this module only marshals
numpy arrays into the C calls.  Names follow the reference: correspondences, loss function,
Solve.  All paths cited relative to /root/reference/nonlinear_optimizer/.
"""
import ctypes

import numpy as np

from . import _capi
from ._capi import RegisterResult, SolveOptions, SolveResult, c_double_p

LOSS_NONE, LOSS_EXPONENTIAL, LOSS_HUBER, LOSS_CAUCHY = 0, 1, 2, 3
TRACE6, TRACE3 = 36, 17


class NloError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("nlo error %d: %s" % (code, message))
        self.code = code


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def identity_pose():
    return np.eye(4).T.reshape(16).copy()


def pose_from_Rt(R, t):
    T = np.eye(4)
    T[:3, :3] = np.asarray(R, dtype=np.float64).reshape(3, 3)
    T[:3, 3] = t
    return np.ascontiguousarray(T.T).reshape(16).copy()


def pose_to_Rt(pose16):
    T = np.asarray(pose16, dtype=np.float64).reshape(4, 4).T
    return T[:3, :3].copy(), T[:3, 3].copy()


class Options:
    """options.h:15-28 (fields the Solve() bodies read)."""

    def __init__(self, max_iterations=40, parameter_tolerance=1e-6, gradient_tolerance=1e-6):
        self.max_iterations = max_iterations
        self.parameter_tolerance = parameter_tolerance
        self.gradient_tolerance = gradient_tolerance

    def _c(self):
        return SolveOptions(self.max_iterations, 0, self.parameter_tolerance,
                            self.gradient_tolerance)


class Context:
    """One device (`device=k`), or several devices of this process (`devices=[...]`): a single
    problem is then sharded by point range over them, a batched one by registration id
    (nlo_context_create_multi)."""

    def __init__(self, device=0, devices=None):
        self._lib = _capi.load()
        h = ctypes.c_void_p()
        if devices is not None:
            arr = np.ascontiguousarray(devices, dtype=np.int32)
            rc = self._lib.nlo_context_create_multi(arr.ctypes.data_as(_capi.c_int32_p), len(arr),
                                                    ctypes.byref(h))
            if rc != 0:
                raise NloError(rc, "nlo_context_create_multi(%s) failed" % list(arr))
            device = int(arr[0])
        else:
            rc = self._lib.nlo_context_create(device, ctypes.byref(h))
            if rc != 0:
                raise NloError(rc, "nlo_context_create(device=%d) failed (no usable sm_100 GPU?)" % device)
        self._h = h
        self.device = device

    @property
    def device_count(self):
        return int(self._lib.nlo_context_device_count(self._h))

    def ingest_stats(self):
        """(wall ms, host gather ms) of the last upload on this context."""
        t = ctypes.c_double(0); g = ctypes.c_double(0)
        self._check(self._lib.nlo_ingest_stats(self._h, ctypes.byref(t), ctypes.byref(g)))
        return t.value, g.value

    def comm_suspend(self, suspended=True):
        """Later assemble / solve calls use this rank's own, un-reduced sums (parity checks)."""
        self._check(self._lib.nlo_comm_suspend(self._h, int(bool(suspended))))

    def _check(self, rc):
        if rc != 0:
            raise NloError(rc, self._lib.nlo_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.nlo_context_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        sm = ctypes.c_int(0); grid = ctypes.c_int(0)
        self._check(self._lib.nlo_context_info(self._h, ctypes.byref(sm), ctypes.byref(grid)))
        return {"sm_count": sm.value, "assemble_grid": grid.value}

    def set_loss(self, kind, params=None):
        """SetLossFunction (mahalanobis_distance_minimizer.h:29)."""
        arr = np.zeros(2)
        params = list(params or [])
        arr[:len(params)] = params
        self._check(self._lib.nlo_set_loss(self._h, kind, _dp(arr)))

    def synchronize(self):
        self._check(self._lib.nlo_synchronize(self._h))

    # ---- communicators ----
    def comm_unique_id(self):
        buf = (ctypes.c_uint8 * 128)()
        self._check(self._lib.nlo_comm_unique_id(self._h, buf))
        return bytes(buf)

    def comm_init_nccl(self, unique_id, rank, nranks):
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self._lib.nlo_comm_init_nccl(self._h, buf, rank, nranks))

    def comm_peer_export(self):
        buf = (ctypes.c_uint8 * 64)()
        self._check(self._lib.nlo_comm_peer_export(self._h, buf))
        return bytes(buf)

    def comm_peer_init(self, handles, rank, nranks):
        blob = b"".join(handles)
        buf = (ctypes.c_uint8 * len(blob)).from_buffer_copy(blob)
        self._check(self._lib.nlo_comm_peer_init(self._h, buf, rank, nranks))

    def comm_destroy(self):
        self._check(self._lib.nlo_comm_destroy(self._h))


class _Problem:
    def __init__(self, ctx):
        self.ctx = ctx
        self._lib = ctx._lib
        self._h = ctypes.c_void_p()

    def close(self):
        if getattr(self, "_h", None) and self.ctx._h:
            self._lib.nlo_problem_destroy(self.ctx._h, self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def size(self):
        return int(self._lib.nlo_problem_size(self._h))

    def set_global_range(self, global_begin, global_total):
        """This problem is the shard [global_begin, ...) of a scan of global_total points spread over
        ranks: the planar solve then drops the tail of the WHOLE list, as the reference."""
        self.ctx._check(self._lib.nlo_problem_set_global_range(self.ctx._h, self._h, global_begin,
                                                               global_total))

    def _assemble(self, fn, nh, ng, pose16, begin, end, problem_index):
        pose = _f64(pose16).reshape(16)
        H = np.zeros(nh); g = np.zeros(ng); cost = ctypes.c_double(0)
        self.ctx._check(fn(self.ctx._h, self._h, problem_index, _dp(pose), begin, end, _dp(H),
                           _dp(g), ctypes.byref(cost)))
        return H, g, cost.value

    def _solve_batched(self, fn, poses, options):
        options = options or Options()
        if not getattr(self, "batched", False):
            raise NloError(-1, "a batched solve needs a problem created with counts=[...]")
        B = len(self.counts)
        poses = _f64(poses).reshape(B, 16).copy()
        opt = options._c()
        res = (SolveResult * B)()
        self.ctx._check(fn(self.ctx._h, self._h, ctypes.byref(opt), _dp(poses), res))
        return {"poses": poses,
                "iterations": np.array([r.iterations for r in res]),
                "final_cost": np.array([r.final_cost for r in res]),
                "device_ms": res[0].device_ms}

    def _solve(self, fn, width, pose16, options, want_trace):
        pose = _f64(pose16).reshape(16).copy()
        opt = options._c()
        res = SolveResult()
        trace = np.zeros((max(options.max_iterations, 1), width)) if want_trace else None
        self.ctx._check(fn(self.ctx._h, self._h, ctypes.byref(opt), _dp(pose), ctypes.byref(res),
                           _dp(trace) if want_trace else None))
        out = {"pose": pose, "iterations": res.iterations, "final_cost": res.final_cost,
               "device_ms": res.device_ms, "status": res.status}
        if want_trace:
            rows = min(res.iterations + 1, options.max_iterations)
            out["trace"] = trace[:rows].copy()
        return out


class NdtProblem(_Problem):
    """NDT / Mahalanobis correspondences (mahalanobis_distance_minimizer/types.h:11-26)."""

    def __init__(self, ctx, capacity=None, counts=None, storage="f64"):
        """storage="f32": correspondences stored as float on the device (fp64 math), the opt-in
        throughput mode; the default "f64" is the parity mode."""
        super().__init__(ctx)
        self.batched = counts is not None
        self.counts = None
        self.storage = storage
        if self.batched:
            if storage != "f64":
                raise NloError(-1, "batched problems are fp64 only")
            self.counts = np.ascontiguousarray(counts, dtype=np.int64)
            ctx._check(self._lib.nlo_ndt_create_batched(
                ctx._h, len(self.counts), self.counts.ctypes.data_as(_capi.c_int64_p),
                ctypes.byref(self._h)))
        elif storage == "f32":
            ctx._check(self._lib.nlo_ndt_create_f32(ctx._h, int(capacity), ctypes.byref(self._h)))
        else:
            ctx._check(self._lib.nlo_ndt_create(ctx._h, int(capacity), ctypes.byref(self._h)))

    def upload(self, point, mean, sqrt_info):
        point = _f64(point); mean = _f64(mean); sqrt_info = _f64(sqrt_info)
        n = point.size // 3
        self.ctx._check(self._lib.nlo_ndt_upload(self.ctx._h, self._h, n, point.ctypes.data,
                                                 mean.ctypes.data, sqrt_info.ctypes.data))

    def upload_f32(self, point, mean, sqrt_info):
        """Float host arrays into an fp32-storage problem (half the PCIe bytes)."""
        point = np.ascontiguousarray(point, dtype=np.float32)
        mean = np.ascontiguousarray(mean, dtype=np.float32)
        sqrt_info = np.ascontiguousarray(sqrt_info, dtype=np.float32)
        self.ctx._check(self._lib.nlo_ndt_upload_f32(self.ctx._h, self._h, point.size // 3,
                                                     point.ctypes.data, mean.ctypes.data,
                                                     sqrt_info.ctypes.data))

    def upload_f32_ptr(self, n, point_ptr, mean_ptr, sqrt_info_ptr):
        self.ctx._check(self._lib.nlo_ndt_upload_f32(self.ctx._h, self._h, n, point_ptr, mean_ptr,
                                                     sqrt_info_ptr))

    def upload_ptr(self, n, point_ptr, mean_ptr, sqrt_info_ptr):
        """Upload from raw host pointers (e.g. pinned memory from nlo_host_alloc)."""
        self.ctx._check(self._lib.nlo_ndt_upload(self.ctx._h, self._h, n, point_ptr, mean_ptr,
                                                 sqrt_info_ptr))

    def upload_aos(self, records, n, stride, off_point, off_mean, off_sqrt, col_major):
        records = np.ascontiguousarray(records)
        self.ctx._check(self._lib.nlo_ndt_upload_aos(self.ctx._h, self._h, n, records.ctypes.data,
                                                     stride, off_point, off_mean, off_sqrt,
                                                     int(col_major)))

    def generate(self, n, seed, index_offset, noise_sigma, true_pose, init_pose, grid):
        """grid: dict(origin[3], dims[3], voxel, mean[cells,3], sqrt_info[cells,9], valid[cells])"""
        tp = _f64(true_pose).reshape(16); ip = _f64(init_pose).reshape(16)
        origin = _f64(grid["origin"]); dims = np.ascontiguousarray(grid["dims"], dtype=np.int32)
        mean = _f64(grid["mean"]); sq = _f64(grid["sqrt_info"])
        valid = np.ascontiguousarray(grid["valid"], dtype=np.uint8)
        self.ctx._check(self._lib.nlo_ndt_generate(
            self.ctx._h, self._h, n, seed, index_offset, noise_sigma, _dp(tp), _dp(ip), _dp(origin),
            dims.ctypes.data_as(_capi.c_int32_p), float(grid["voxel"]), _dp(mean), _dp(sq),
            valid.ctypes.data_as(_capi.c_uint8_p)))

    def generate_batched(self, seed, noise_sigma, true_poses, init_pose, grid):
        """cfg5: registration k = stream seed + k, in the sensor frame of true_poses[k]."""
        tp = _f64(true_poses).reshape(len(self.counts), 16); ip = _f64(init_pose).reshape(16)
        origin = _f64(grid["origin"]); dims = np.ascontiguousarray(grid["dims"], dtype=np.int32)
        mean = _f64(grid["mean"]); sq = _f64(grid["sqrt_info"])
        valid = np.ascontiguousarray(grid["valid"], dtype=np.uint8)
        self.ctx._check(self._lib.nlo_ndt_generate_batched(
            self.ctx._h, self._h, seed, noise_sigma, _dp(tp), _dp(ip), _dp(origin),
            dims.ctypes.data_as(_capi.c_int32_p), float(grid["voxel"]), _dp(mean), _dp(sq),
            valid.ctypes.data_as(_capi.c_uint8_p)))

    def download(self, begin, end, problem_index=0):
        """point[n,3], mean[n,3], information[n,6] = unique entries (00 01 02 11 12 22) of S^T S
        of correspondences [begin, end) (of registration `problem_index` for a batched problem)."""
        n = end - begin
        point = np.zeros((n, 3)); mean = np.zeros((n, 3)); info = np.zeros((n, 6))
        self.ctx._check(self._lib.nlo_ndt_download_problem(self.ctx._h, self._h, problem_index, begin,
                                                           end, _dp(point), _dp(mean), _dp(info)))
        return point, mean, info

    def assemble6(self, pose16, begin=0, end=None, problem_index=0):
        """..._analytic.cc:12-52 over [begin, end)."""
        if end is None:
            end = int(self.counts[problem_index]) if self.batched else self.size
        return self._assemble(self._lib.nlo_ndt6_assemble, 21, 6, pose16, begin, end, problem_index)

    def assemble3(self, pose16, begin=0, end=None, problem_index=0):
        """..._analytic_3dof.cc:33-68; default end = floor(n/4)*4 as the reference's Solve."""
        if end is None:
            n = int(self.counts[problem_index]) if self.batched else self.size
            end = (n // 4) * 4
        return self._assemble(self._lib.nlo_ndt3_assemble, 6, 3, pose16, begin, end, problem_index)

    def solve6(self, pose16, options=None, trace=False):
        """MahalanobisDistanceMinimizerAnalytic::Solve, ..._analytic.cc:54-157."""
        return self._solve(self._lib.nlo_ndt6_solve, TRACE6, pose16, options or Options(), trace)

    def solve3(self, pose16, options=None, trace=False):
        """MahalanobisDistanceMinimizerAnalytic3DOF::Solve, ..._analytic_3dof.cc:14-108."""
        return self._solve(self._lib.nlo_ndt3_solve, TRACE3, pose16, options or Options(), trace)

    def solve6_batched(self, poses, options=None):
        return self._solve_batched(self._lib.nlo_ndt6_solve_batched, poses, options)

    def solve3_batched(self, poses, options=None):
        return self._solve_batched(self._lib.nlo_ndt3_solve_batched, poses, options)


class ReprojProblem(_Problem):
    """3D-2D correspondences (reprojection_error_minimizer/types.h:14-28)."""

    def __init__(self, ctx, capacity=None, counts=None):
        super().__init__(ctx)
        self.batched = counts is not None
        self.counts = None
        if self.batched:
            self.counts = np.ascontiguousarray(counts, dtype=np.int64)
            ctx._check(self._lib.nlo_reproj_create_batched(
                ctx._h, len(self.counts), self.counts.ctypes.data_as(_capi.c_int64_p),
                ctypes.byref(self._h)))
        else:
            ctx._check(self._lib.nlo_reproj_create(ctx._h, int(capacity), ctypes.byref(self._h)))

    def solve_batched(self, poses, options=None):
        return self._solve_batched(self._lib.nlo_reproj_solve_batched, poses, options)

    def upload(self, local_point, pixel, intrinsics):
        X = _f64(local_point); px = _f64(pixel); K = _f64(intrinsics)
        n = X.size // 3
        self.ctx._check(self._lib.nlo_reproj_upload(self.ctx._h, self._h, n, X.ctypes.data,
                                                    px.ctypes.data, _dp(K)))

    def upload_aos(self, records, n, stride, off_point, off_pixel, intrinsics):
        """The reference's 40-byte Correspondence records in place (types.h:14-17)."""
        records = np.ascontiguousarray(records); K = _f64(intrinsics)
        self.ctx._check(self._lib.nlo_reproj_upload_aos(self.ctx._h, self._h, n, records.ctypes.data,
                                                        stride, off_point, off_pixel, _dp(K)))

    def assemble(self, pose16, begin=0, end=None):
        """reprojection_error_minimizer_analytic.cc:31-63."""
        end = self.size if end is None else end
        return self._assemble(self._lib.nlo_reproj_assemble, 21, 6, pose16, begin, end, 0)

    def solve(self, pose16, options=None, trace=False):
        """ReprojectionErrorMinimizerAnalytic::Solve, ..._analytic.cc:12-105."""
        return self._solve(self._lib.nlo_reproj_solve, TRACE6, pose16, options or Options(), trace)


class NdtMap:
    """NDT map on the device (UpdateNdtMap of the reference's test mains,
    mahalanobis_distance_minimizer/tests/simple_optimization_test.cc:236-280): a dense voxel grid
    over the bounding box, or -- `hashed=True`, and by itself when the box has more than 2^28
    voxels -- a voxel hash over the occupied voxels only (the reference's unordered_map, :282-294)."""

    def __init__(self, ctx, grid=None, points=None, voxel=None, v_not_transposed=False, hashed=False):
        self.ctx = ctx
        self._lib = ctx._lib
        self._h = ctypes.c_void_p()
        if grid is not None:
            origin = _f64(grid["origin"]); dims = np.ascontiguousarray(grid["dims"], dtype=np.int32)
            mean = _f64(grid["mean"]); sq = _f64(grid["sqrt_info"])
            valid = np.ascontiguousarray(grid["valid"], dtype=np.uint8)
            ctx._check(self._lib.nlo_ndt_map_create(
                ctx._h, _dp(origin), dims.ctypes.data_as(_capi.c_int32_p), float(grid["voxel"]),
                _dp(mean), _dp(sq), valid.ctypes.data_as(_capi.c_uint8_p), ctypes.byref(self._h)))
        else:
            pts = _f64(points)
            build = self._lib.nlo_ndt_map_build_hashed if hashed else self._lib.nlo_ndt_map_build
            ctx._check(build(ctx._h, pts.size // 3, pts.ctypes.data, float(voxel),
                             int(v_not_transposed), ctypes.byref(self._h)))

    def layout(self):
        """(hashed, rows of the downloaded tables)."""
        hashed = ctypes.c_int32(0); cells = ctypes.c_int64(0)
        self.ctx._check(self._lib.nlo_ndt_map_layout(self.ctx._h, self._h, ctypes.byref(hashed),
                                                     ctypes.byref(cells)))
        return bool(hashed.value), cells.value

    def to_grid(self):
        origin = np.zeros(3); dims = np.zeros(3, dtype=np.int32)
        voxel = ctypes.c_double(0); nvalid = ctypes.c_int64(0)
        self.ctx._check(self._lib.nlo_ndt_map_info(self.ctx._h, self._h, _dp(origin),
                                                   dims.ctypes.data_as(_capi.c_int32_p),
                                                   ctypes.byref(voxel), ctypes.byref(nvalid)))
        hashed, cells = self.layout()
        mean = np.zeros((cells, 3)); sq = np.zeros((cells, 9)); valid = np.zeros(cells, dtype=np.uint8)
        self.ctx._check(self._lib.nlo_ndt_map_download(self.ctx._h, self._h, _dp(mean), _dp(sq),
                                                       valid.ctypes.data_as(_capi.c_uint8_p)))
        out = {"origin": origin, "dims": dims, "voxel": voxel.value, "mean": mean,
               "sqrt_info": sq, "valid": valid, "valid_cells": nvalid.value, "hashed": hashed}
        if hashed:
            # rows are hash slots; keys = x | y << 21 | z << 42 (voxel indices from origin), all ones = free
            keys = np.zeros(cells, dtype=np.uint64)
            self.ctx._check(self._lib.nlo_ndt_map_download_keys(self.ctx._h, self._h, keys.ctypes.data))
            out["keys"] = keys
        return out

    def close(self):
        if getattr(self, "_h", None) and self.ctx._h:
            self._lib.nlo_ndt_map_destroy(self.ctx._h, self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scan:
    """Scan points (sensor frame) resident on the device."""

    def __init__(self, ctx, points):
        self.ctx = ctx
        self._lib = ctx._lib
        self._h = ctypes.c_void_p()
        pts = _f64(points)
        self.n = pts.size // 3
        ctx._check(self._lib.nlo_scan_create(ctx._h, self.n, pts.ctypes.data, ctypes.byref(self._h)))

    def match(self, ndt_map, pose16, problem, radius=1.0, max_neighbors=2):
        """MatchPointCloud (simple_optimization_test.cc:296-342) into `problem`."""
        pose = _f64(pose16).reshape(16)
        matched = ctypes.c_int64(0)
        self.ctx._check(self._lib.nlo_ndt_match(self.ctx._h, self._h, ndt_map._h, _dp(pose), radius,
                                                max_neighbors, problem._h, ctypes.byref(matched)))
        return matched.value

    def register(self, ndt_map, pose16, options=None, radius=1.0, max_neighbors=2, max_outer=10,
                 three_dof=False):
        """The outer match + Solve loop of OptimizePoseAnalytic (:473-505), on the device."""
        options = options or Options()
        pose = _f64(pose16).reshape(16).copy()
        opt = options._c()
        res = RegisterResult()
        self.ctx._check(self._lib.nlo_ndt_register(self.ctx._h, self._h, ndt_map._h, ctypes.byref(opt),
                                                   radius, max_neighbors, max_outer, int(three_dof),
                                                   _dp(pose), ctypes.byref(res)))
        return {"pose": pose, "outer_iterations": res.outer_iterations,
                "inner_iterations": res.inner_iterations, "final_cost": res.final_cost,
                "device_ms": res.device_ms, "matched": res.matched}

    def close(self):
        if getattr(self, "_h", None) and self.ctx._h:
            self._lib.nlo_scan_destroy(self.ctx._h, self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def host_alloc(nbytes):
    """Pinned host memory as a numpy uint8 array (nlo_host_alloc)."""
    lib = _capi.load()
    p = ctypes.c_void_p()
    rc = lib.nlo_host_alloc(ctypes.byref(p), nbytes)
    if rc != 0:
        raise NloError(rc, "nlo_host_alloc(%d) failed" % nbytes)
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=np.uint8)
    return arr, p


def host_free(p):
    _capi.load().nlo_host_free(p)


def guard_report():
    """State of the guard-band check of the library's device buffers (nlo_debug_guard_report;
    active when NLO_GUARD=1 was in the environment before the library's first allocation)."""
    en = ctypes.c_int32(0)
    checked, live, bad = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
    _capi.load().nlo_debug_guard_report(ctypes.byref(en), ctypes.byref(checked), ctypes.byref(live),
                                        ctypes.byref(bad))
    return {"enabled": bool(en.value), "allocations_checked": checked.value,
            "allocations_live": live.value, "corrupted_bytes": bad.value}
