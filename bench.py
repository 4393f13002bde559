#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric).

Workload (config.workload = "cfg4"): NDT 6-DoF registration of a 64M-point synthetic scan against
the 0.5 m-voxel NDT map of the 7x5x2.5 m room, Exponential(1,1) loss, fp64 -- BASELINE.json
configs[3], the configuration the metric is quoted on; it fits one B200 (7.68 GB of SoA planes).
A "step" is one damped Gauss-Newton iteration over the whole scan: residual + analytic Jacobian +
robust weight + reduction of the 28 H|g|cost doubles + 6x6 solve + pose update, all on the device.

  python bench.py [--gpus N] [--steps K] [--warmup W]            -> this repo's CUDA path
  python bench.py --impl reference ...                           -> the reference's CPU path
                                                                    (oracle port, all host threads)
Under torchrun (N > 1) the scan is sharded by point range (strong scaling: 64M points in total)
and the 28 doubles are all-reduced each iteration over NVLink (peer-memory one-shot all-reduce
fused into the iteration kernel; --comm nccl selects ncclAllReduce instead).

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


class StdoutGuard:
    """stdout must carry exactly ONE JSON line, but native libraries (NCCL's version banner, ...)
    write to file descriptor 1 behind Python's back.  For the duration of the run fd 1 is pointed
    at stderr; emit() writes the JSON line to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.real, (line + "\n").encode())


GUARD = None

BYTES_PER_CORR = 96   # 12 fp64 scalars streamed per correspondence per iteration: point 3 + mean 3 +
                      # the 6 unique entries of the information matrix S^T S, which is formed ONCE at
                      # ingest (SURVEY.md 8d counts 15 scalars = 120 B for streaming S itself)
HOST_BYTES_PER_CORR = 120  # what a caller hands over: point 3 + mean 3 + sqrt_information 9
TOTAL_POINTS = 64 * 1024 * 1024
SEED = 1004


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch(points, iterations_in_launch):
    """dram__bytes_read + dram__bytes_write per launch of the iteration kernel, scaled from the
    committed `ncu --set full` capture (profiles/ncu_traffic.json: bytes per iteration at 64M
    points; traffic is proportional to points x iterations in the launch)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        return t["dram_bytes_per_iteration"] * (points / t["points"]) * iterations_in_launch
    except Exception:
        return None


class ClockSampler:
    """SM clock and throttle reasons sampled WHILE the timed region runs: NVML polled every ~2 ms
    from a thread (the region can be as short as 20 ms); nvidia-smi -lms as a fallback."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.samples = []   # (timestamp, sm_mhz, reasons bitmask or None)
        self.rows = []
        self.proc = None
        self.thread = None
        self.stop_flag = False
        self.max_mhz = None
        self.nvml = None

    def _nvml_loop(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((time.perf_counter(), float(mhz), int(reasons)))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), line.strip()))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self, t0=None, t1=None):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            nv = self.nvml
            inside = [x for x in self.samples if t0 is None or t0 <= x[0] <= t1] or self.samples
            bits = 0
            for x in inside:
                bits |= x[2]
            names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            reasons = sorted(k for k, v in names.items() if bits & v)
            sm = [x[1] for x in inside]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "samples": len(sm), "source": "nvml, 2 ms period, inside the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for (ts, r) in self.rows if t0 is None or (t0 - 0.05 <= ts <= t1 + 0.15)] or \
               [r for (_, r) in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


# ------------------------------------------------------------------------------ reference arm
def cpu_reference_rates(sample_points, min_seconds, threads):
    """Times the reference's CPU assembly (oracle port) on `sample_points` correspondences of the
    same workload with all host threads: scalar double (..._analytic.cc:12-52 + executor split)
    and float AVX2+FMA SoA (SolveFloatIntrinsicAligned, ..._simd_various.cc:1300-1447)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import nlo_oracle_py as oracle
    from nonlinear_optimizer_for_slam_b200 import synthetic as syn
    oracle.build()
    native = oracle.use_native_build()   # -march=native on this host, as the reference builds
    point, mean, S = syn.ndt_problem(sample_points, SEED, syn.CFG1_TRUE)
    n = len(point)
    R = np.eye(3); t = np.zeros(3)
    planes = oracle.simd_pack(point, mean, S)
    out = {"native_build": native}
    for name, fn in (
            ("scalar_f64", lambda: oracle.ndt6_assemble_threads(point, mean, S, R, t, 1, [1.0, 1.0], threads)),
            ("avx2_f32", lambda: oracle.simd_ndt6_assemble(planes, n, R, t, 1, [1.0, 1.0], threads))):
        fn()  # warm
        passes, t0 = 0, time.perf_counter()
        while True:
            fn(); passes += 1
            dt = time.perf_counter() - t0
            if dt >= min_seconds and passes >= 3:
                break
        out[name] = {"gpoints_s": n * passes / dt / 1e9, "passes": passes, "seconds": dt}
    return n, out


def workload_name(total):
    return ("cfg4: NDT 6-DoF, %d-point scan vs 0.5 m-voxel NDT map, Exponential(1,1), sharded by point range"
            % total)


def run_reference(args):
    """The reference's own CPU path on the SAME workload: one step = one float AVX2+FMA SoA assembly
    pass (SolveFloatIntrinsicAligned restated, thread split of ..._analytic_simd.cc:55-76) over all
    `--points` correspondences on every host thread.  The scan is the cfg4 generator's first 4M
    correspondences tiled to the full size (the rate does not depend on the values; 64Mi points are
    4 GB of float planes)."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import nlo_oracle_py as oracle
    from nonlinear_optimizer_for_slam_b200 import synthetic as syn
    oracle.build()
    native = oracle.use_native_build()   # -march=native on this host, as the reference builds
    total = args.points
    block = min(total, 4 * 1024 * 1024)
    point, mean, S = syn.ndt_problem(block + block // 64, SEED, syn.CFG1_TRUE)
    assert len(point) >= block
    point, mean, S = point[:block], mean[:block], S[:block]
    base = oracle.simd_pack(point, mean, S).reshape(15, block)
    reps = (total + block - 1) // block
    planes = np.ascontiguousarray(np.tile(base, (1, reps))[:, :total]).reshape(-1)
    n = total
    R = np.eye(3); t = np.zeros(3)
    step = lambda: oracle.simd_ndt6_assemble(planes, n, R, t, 1, [1.0, 1.0], threads)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = n * args.steps / dt / 1e9
    # the double-precision scalar path on one block, for the record
    ts = time.perf_counter()
    oracle.ndt6_assemble_threads(point, mean, S, R, t, 1, [1.0, 1.0], threads)
    scalar = block / (time.perf_counter() - ts) / 1e9
    sample_txt = ("all %d correspondences of cfg4 per step (the generator's first %d, seed %d, tiled), float "
                  "AVX2+FMA SoA assembly (SolveFloatIntrinsicAligned restated) on %d std::threads, %s"
                  % (n, block, SEED, threads, "-O2 -march=native" if native else "-O2 -mavx2 -mfma"))
    line = {
        "impl": "reference", "metric": "NDT 6-DoF assembly Gpoints/s", "value": value,
        "unit": "Gpoints/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(total), "points_total": total},
        "cpu_baseline": {"value": value, "unit": "Gpoints/s", "cores": threads, "kind": "port",
                         "sample": sample_txt, "scalar_f64_gpoints_s": scalar},
        "e2e": {"value": value, "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    GUARD.emit(json.dumps(line))


# ------------------------------------------------------------------------------ CUDA arm
def run_cuda(args):
    import nonlinear_optimizer_for_slam_b200 as nlo
    from nonlinear_optimizer_for_slam_b200 import synthetic as syn

    rank, local_rank, world = dist_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)

    total = args.points
    n_local = total // world
    offset = rank * n_local
    if rank == world - 1:
        n_local = total - offset

    ctx = nlo.Context(local_rank)
    ctx.set_loss(nlo.LOSS_EXPONENTIAL, [1.0, 1.0])
    grid = syn.room_ndt_grid(0.5)
    prob = nlo.NdtProblem(ctx, capacity=n_local)
    true16 = syn.to_pose16(syn.CFG1_TRUE)
    prob.generate(n_local, SEED, offset, 0.01, true16, nlo.identity_pose(), grid)

    comm = "none"
    if world > 1:
        comm = args.comm
        if comm == "peer":
            try:
                handle = ctx.comm_peer_export()
                handles = [None] * world
                dist.all_gather_object(handles, handle)
                ctx.comm_peer_init(handles, rank, world)
            except Exception as e:  # CUDA IPC not permitted in this container -> NCCL
                ok = 0
                if rank == 0:
                    print("peer comm unavailable (%s); falling back to NCCL" % e, file=sys.stderr)
                comm = "nccl"
            flags = [None] * world
            dist.all_gather_object(flags, comm)
            if any(f != "peer" for f in flags) and comm == "peer":
                ctx.comm_destroy()
                comm = "nccl"
        if comm == "nccl":
            uid = [ctx.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            ctx.comm_init_nccl(uid[0], rank, world)

    def barrier():
        ctx.synchronize()
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    never = dict(parameter_tolerance=0.0, gradient_tolerance=0.0)  # norm < 0 is never true
    pose0 = nlo.identity_pose()
    # warm-up: W untimed steps (also instantiates the CUDA graph of a K-step loop)
    prob.solve6(pose0, nlo.Options(max_iterations=max(args.warmup, 3), **never))
    prob.solve6(pose0, nlo.Options(max_iterations=args.steps, **never))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    t0 = time.perf_counter()
    res = prob.solve6(pose0, nlo.Options(max_iterations=args.steps, **never))
    barrier()
    t1 = time.perf_counter()
    assert res["iterations"] == args.steps, res
    ms_local = res["device_ms"]  # CUDA events on the context stream around the K-iteration loop
    ms = ms_local
    per_rank_ms = [ms_local]
    if dist is not None:
        import torch
        tms = torch.tensor([ms_local], dtype=torch.float64, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
        per_rank_ms = [None] * world
        dist.all_gather_object(per_rank_ms, ms_local)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    parity = parity_check(nlo, ctx, prob, res, n_local, rank, world, dist)

    value = total * args.steps / (ms * 1e-3) / 1e9
    iters_per_s = args.steps / (ms * 1e-3)
    peak, peak_src = measured_peak_gbs()
    achieved = n_local * BYTES_PER_CORR * args.steps / (ms_local * 1e-3) / 1e9

    # ---- e2e: one reference-facing Solve from HOST buffers (upload + loop + pose read-back)
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(nlo, ctx, prob, n_local, args, dist, total)

    # ---- secondary: latency-bound configs (iterations/s, N = 1 only) and the batched config
    extra = {}
    if not args.no_extra:
        if world == 1:
            extra = measure_small_configs(nlo, syn, ctx)
            if not args.no_cpu:
                try:
                    sec = cpu_secondary_rates(syn, os.cpu_count() or 1)
                    extra["secondary"]["cfg2_ndt3_1m_huber"]["cpu_baseline"] = sec["cfg2"]
                    extra["secondary"]["cfg3_pnp_50k_cauchy"]["cpu_baseline"] = sec["cfg3"]
                except Exception as e:
                    extra["secondary"]["cpu_baseline_error"] = str(e)
        if comm != "none":
            ctx.comm_destroy()   # cfg5 partitions independent registrations: no collective at all
            comm_destroyed = True
        extra.setdefault("secondary", {})["cfg5_batched_4096x20k"] = measure_batched(
            nlo, syn, ctx, rank, world, dist)

    dropin = None
    if not args.no_dropin:
        if dist is not None:
            ctx.synchronize(); dist.barrier()
        if rank == 0:
            dropin = measure_dropin(world)
        if dist is not None:
            dist.barrier()       # the other ranks' GPUs stay idle while rank 0's child process uses them

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        n_s, rates = cpu_reference_rates(2_000_000, 4.0, threads)
        cpu = {"value": rates["avx2_f32"]["gpoints_s"], "unit": "Gpoints/s", "cores": threads,
               "kind": "port",
               "sample": "%d correspondences of cfg4, >= 4 s per variant; value = float AVX2+FMA SoA "
                         "path (fastest reference variant) on all threads, %s"
                         % (n_s, "-O2 -march=native" if rates["native_build"] else "-O2 -mavx2 -mfma"),
               "scalar_f64_gpoints_s": rates["scalar_f64"]["gpoints_s"]}

    if rank == 0:
        # the whole K-iteration loop is ONE persistent cooperative launch of gn_iteration_kernel per
        # GPU (leader CTA reduces, [peer all-reduce over NVLink], steps, publishes the state);
        # nccl: assemble + step kernels per iteration (plus NCCL's own kernel, not counted)
        launches = 1 if comm != "nccl" else 2 * args.steps
        line = {
            "metric": "NDT 6-DoF assembly Gpoints/s", "value": value, "unit": "Gpoints/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "gn_iterations_per_s": iters_per_s,
            "per_rank_ms_per_step": [round(v / args.steps, 6) for v in per_rank_ms],
            "config": {"workload": workload_name(total),
                       "points_total": total, "points_per_gpu": n_local, "comm": comm,
                       "bytes_per_correspondence": BYTES_PER_CORR,
                       "storage": "fp64, 12 scalars per correspondence: point 3 + mean 3 + S^T S (6 unique), "
                                  "S^T S formed once at ingest",
                       "l2": "inputs (%.2f GB per GPU) exceed the 126 MB L2; no flush needed"
                             % (n_local * BYTES_PER_CORR / 1e9),
                       "timing": "CUDA events on each rank's launch stream around the K-iteration loop, max over ranks"
                                 + ("; a sharded Solve() opens with a device-side barrier of the ranks (one NVLink "
                                    "exchange, in front of the start event), so the ranks' timed regions start together"
                                    if comm == "peer" else "")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": ncu_traffic_per_launch(n_local, args.steps if comm != "nccl" else 1),
                         "traffic_source": "profiles/ncu_traffic.json: dram__bytes_read + dram__bytes_write of "
                                           "one `ncu --set full` capture of this kernel at 64M points, scaled by "
                                           "points x iterations of this launch (not measured in this run)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_iteration": n_local * BYTES_PER_CORR,
                         "algorithmic_bytes_per_launch": n_local * BYTES_PER_CORR * (args.steps if comm != "nccl" else 1),
                         "kernel": "gn_iteration_kernel<ndt6, exponential>"},
            "cpu_baseline": cpu,
            "parity_check": parity,
            "e2e": e2e,
            "e2e_dropin": dropin,
            "gpu_launches": launches,
            "clocks": clocks,
        }
        line.update(extra)
        GUARD.emit(json.dumps(line))
    prob.close()
    if dist is not None:
        barrier()
        dist.destroy_process_group()
    ctx.close()


def parity_check(nlo, ctx, prob, res, n_local, rank, world, dist):
    """Correctness of THIS run, visible in the JSON line at every N:
      * the all-reduced H | g | cost of the sharded assembly against the ranks' own un-reduced sums
        (communicator suspended) gathered over gloo and added on the host in rank order;
      * every rank's final pose of the timed solve, bit for bit;
      * every rank's first (up to) 1M correspondences, read back from the device, against the
        long-double CPU oracle.
    The run fails above 1e-6 (BASELINE.json north_star tolerance)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import nlo_oracle_py as oracle
    from parity import rel_errors
    from nonlinear_optimizer_for_slam_b200 import sharding, synthetic as syn
    T = syn.yaw_pose([0.05, -0.02, 0.03], 0.02)          # not the identity: R enters the sums
    pose = syn.to_pose16(T)
    H, g, c = prob.assemble6(pose)                       # all-reduced over the ranks
    if world > 1:
        ctx.comm_suspend(True)
    Hl, gl, cl = prob.assemble6(pose)                    # this rank's shard only
    m = min(n_local, 1 << 20)
    Hs, gs, cs = prob.assemble6(pose, 0, m)
    if world > 1:
        ctx.comm_suspend(False)
    p, mu, info = prob.download(0, m)
    S = syn.sqrt_info_from_information6(info)
    Rq = oracle.quat_to_rotmat(oracle.rotmat_to_quat(T[:3, :3]))
    Ho, go, co = oracle.ndt6_assemble(p, mu, S, Rq, T[:3, 3], 1, [1.0, 1.0], long_double=True)
    e_oracle = max(rel_errors(Hs, gs, cs, Ho, go, co))
    local = np.concatenate([Hl, gl, [cl]])
    parts, poses, e_or = [local], [res["pose"].tobytes()], [e_oracle]
    if dist is not None:
        parts = [None] * world; poses = [None] * world; e_or = [None] * world
        dist.all_gather_object(parts, local)
        dist.all_gather_object(poses, res["pose"].tobytes())
        dist.all_gather_object(e_or, e_oracle)
    tot = sharding.ordered_sum(parts)
    eh, eg, ec = rel_errors(H, g, c, tot[:21], tot[21:27], tot[27])
    out = {"err_H": eh, "err_g": eg, "err_cost": ec, "pose_identical": all(q == poses[0] for q in poses),
           "n_ranks": world, "oracle_sample_points_per_rank": m, "err_vs_oracle_max_over_ranks": max(e_or),
           "tolerance": 1e-6,
           "what": "all-reduced sums vs rank-ordered host sum of the ranks' un-reduced sums; final pose of "
                   "the timed solve bit-identical on all ranks; each rank's first correspondences vs the "
                   "long-double CPU oracle"}
    ok = eh < 1e-6 and eg < 1e-6 and ec < 1e-6 and out["pose_identical"] and max(e_or) < 1e-6
    out["ok"] = bool(ok)
    if not ok:
        raise SystemExit("parity check failed: %s" % json.dumps(out))
    return out


def syn_module():
    from nonlinear_optimizer_for_slam_b200 import synthetic
    return synthetic


def measure_e2e(nlo, ctx, prob, n_local, args, dist, total):
    """Solve() as a caller of the reference API sees it: correspondences in (pinned) host memory,
    one upload + repack, `e2e_iters` device-resident iterations, pose read back."""
    n_e2e = min(n_local, args.e2e_points // (1 if dist is None else dist.get_world_size()))
    iters = args.e2e_iters
    nbytes = n_e2e * HOST_BYTES_PER_CORR
    try:
        arr, handle = nlo.host_alloc(nbytes)
    except Exception as e:
        return {"value": None, "unit": "Gpoints/s", "error": "pinned host allocation failed: %s" % e}
    f64 = arr.view(np.float64)
    point = f64[:3 * n_e2e]; mean = f64[3 * n_e2e:6 * n_e2e]; sq = f64[6 * n_e2e:15 * n_e2e]
    # Host buffers (untimed): this rank's first correspondences come back from the device -- the
    # device keeps S only as S^T S, so an equivalent S is rebuilt as its Cholesky factor -- and the
    # block is tiled to the full size (timing does not depend on the values).
    block = min(n_e2e, 2 * 1024 * 1024)
    p, m, info = prob.download(0, block)
    s = syn_module().sqrt_info_from_information6(info)
    for b in range(0, n_e2e, block):
        e = min(n_e2e, b + block)
        point[3 * b:3 * e] = p[:e - b].ravel(); mean[3 * b:3 * e] = m[:e - b].ravel()
        sq[9 * b:9 * e] = s[:e - b].ravel()
    e2e_prob = nlo.NdtProblem(ctx, capacity=n_e2e)
    opts = nlo.Options(max_iterations=iters, parameter_tolerance=0.0, gradient_tolerance=0.0)
    pose0 = nlo.identity_pose()

    def one_solve():
        e2e_prob.upload_ptr(n_e2e, point.ctypes.data, mean.ctypes.data, sq.ctypes.data)
        return e2e_prob.solve6(pose0, opts)

    one_solve()  # warm-up (staging buffer, graph)
    ctx.synchronize()
    if dist is not None:
        dist.barrier()
    reps = 2
    t0 = time.perf_counter()
    for _ in range(reps):
        r = one_solve()
    ctx.synchronize()
    if dist is not None:
        dist.barrier()
    dt = (time.perf_counter() - t0) / reps
    if dist is not None:
        import torch
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    world = 1 if dist is None else dist.get_world_size()
    e2e_prob.close()
    nlo.host_free(handle)
    return {"value": n_e2e * world * iters / dt / 1e9, "unit": "Gpoints/s",
            "h2d_bytes_per_step": nbytes / iters, "d2h_bytes_per_step": (128 + 32) / iters,
            "h2d_bytes_per_solve": nbytes, "solve_iterations": iters, "seconds_per_solve": dt,
            "points_per_gpu": n_e2e,
            "note": "one step of the e2e arm = one GN iteration inside a full Solve(): pinned host "
                    "correspondences uploaded once per Solve (as the reference's Solve receives them), "
                    "%d iterations on the device, pose + iteration count read back" % iters}


def measure_dropin(world):
    """The reference-facing call itself, timed by the C++ binary cxx/bench/dropin_bench:
    MahalanobisDistanceMinimizerCuda::Solve(options, std::vector<Correspondence> (304-byte AoS records,
    pageable), &pose) in ONE process over `world` GPUs (device list; the split sits inside Solve)."""
    from nonlinear_optimizer_for_slam_b200 import build as nlo_build
    out = {}
    try:
        binary = nlo_build.build_cxx_bench()
    except Exception as e:
        return {"error": "dropin_bench did not build: %s" % e}
    devices = ",".join(str(d) for d in range(world))
    cases = (("cfg1_100k_converged", ["--n", "100000"]),
             ("cfg1_100k_40_iterations", ["--n", "100000", "--force-iterations"]),
             ("cfg4_16M_40_iterations", ["--n", str(16 * 1024 * 1024), "--force-iterations"]),
             ("cfg2_1M_planar_huber_40_iterations", ["--n", "1000000", "--planar", "--loss", "huber",
                                                     "--force-iterations"]))
    for name, extra in cases:
        try:
            r = subprocess.run([binary, "--devices", devices] + extra, capture_output=True, text=True, timeout=600)
            if r.returncode != 0:
                out[name] = {"error": (r.stderr or "")[-400:]}
                continue
            out[name] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:
            out[name] = {"error": str(e)}
    out["note"] = ("timed region = ...Cuda::Solve() on the reference's AoS records in pageable memory: host gather of "
                   "the 15 hot doubles into a pinned ring, PCIe, device repack, the device-resident loop, pose read-back; "
                   "gpoints_s = n x assemblies / wall")
    return out


def cpu_secondary_rates(syn, threads):
    """The reference's float SIMD twins of the planar and reprojection minimizers, timed beside cfg2 / cfg3
    on this host: as the reference runs them (ONE thread: neither uses the executor) and with the 8-lane
    thread split applied on all threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import nlo_oracle_py as oracle
    oracle.build()
    native = oracle.use_native_build()
    flags = "-O2 -march=native" if native else "-O2 -mavx2 -mfma"

    def rate(fn, n, min_seconds=1.5):
        fn()
        passes, t0 = 0, time.perf_counter()
        while True:
            fn(); passes += 1
            dt = time.perf_counter() - t0
            if dt >= min_seconds and passes >= 3:
                return n * passes / dt / 1e9, dt / passes * 1e6

    out = {}
    p, m, s = syn.ndt_problem(1_000_000, 1002, syn.CFG2_TRUE)
    planes = oracle.simd_pack(p, m, s)
    T = syn.CFG2_TRUE
    for label, th in (("1_thread_as_reference", 1), ("all_threads", threads)):
        g, us = rate(lambda: oracle.simd_ndt3_assemble(planes, len(p), np.eye(2), np.zeros(2), 2, [1.0], th), len(p))
        out.setdefault("cfg2", {})[label] = {"gpoints_s": g, "us_per_iteration": us, "cores": th}
    out["cfg2"].update({"kind": "port", "unit": "Gpoints/s", "value": out["cfg2"]["all_threads"]["gpoints_s"],
                        "cores": threads,
                        "sample": "%d correspondences of cfg2, float 8-lane planar twin (..._analytic_3dof_simd.cc:85-157 "
                                  "restated), Huber(1.0), %s" % (len(p), flags)})
    X, px, K = syn.pnp_problem(50_000, 1003)
    planes = oracle.simd_reproj_pack(X, px)
    for label, th in (("1_thread_as_reference", 1), ("all_threads", threads)):
        g, us = rate(lambda: oracle.simd_reproj_assemble(planes, len(X), K, np.eye(3), np.zeros(3), 3, [1e-2], th),
                     len(X))
        out.setdefault("cfg3", {})[label] = {"gpoints_s": g, "us_per_iteration": us, "cores": th}
    best = max(("1_thread_as_reference", "all_threads"), key=lambda k: out["cfg3"][k]["gpoints_s"])
    out["cfg3"].update({"kind": "port", "unit": "Gpoints/s", "value": out["cfg3"][best]["gpoints_s"],
                        "cores": out["cfg3"][best]["cores"],
                        "sample": "%d correspondences of cfg3, float 8-lane reprojection twin (reprojection_error_minimizer_"
                                  "analytic_simd.cc:55-137 restated, quirks kept), Cauchy(0.01), %s; value = the faster of "
                                  "1 thread (as the reference runs it) and all threads" % (len(X), flags)})
    return out


def measure_batched(nlo, syn, ctx, rank, world, dist, num_problems=4096, points=20000, iters=40):
    """cfg5: 4096 independent 20k-point NDT 6-DoF registrations, block-partitioned over the ranks
    (sharding.problem_partition); one CTA per registration runs its whole loop; no collective."""
    from nonlinear_optimizer_for_slam_b200 import sharding
    b, e = sharding.problem_partition(num_problems, rank, world)
    nb = e - b
    rng = np.random.default_rng(2000)
    true = np.zeros((num_problems, 16))
    for k in range(num_problems):  # per-problem true poses: t ~ U(+-0.3), yaw ~ U(+-0.15)
        true[k] = syn.to_pose16(syn.yaw_pose(rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.15, 0.15)))
    ctx.set_loss(nlo.LOSS_EXPONENTIAL, [1.0, 1.0])
    prob = nlo.NdtProblem(ctx, counts=[points] * nb)
    prob.generate_batched(2000 + b, 0.01, true[b:e], nlo.identity_pose(), syn.room_ndt_grid(0.5))
    poses0 = np.tile(nlo.identity_pose(), (nb, 1))
    opts = nlo.Options(max_iterations=iters, parameter_tolerance=0.0, gradient_tolerance=0.0)
    prob.solve6_batched(poses0, opts)
    ctx.synchronize()
    if dist is not None:
        dist.barrier()
    ms = min(prob.solve6_batched(poses0, opts)["device_ms"] for _ in range(3))
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    prob.close()
    work = num_problems * points * iters
    return {"gpoints_s": work / ms / 1e6, "registrations_per_s": num_problems / ms * 1e3,
            "device_ms": ms, "iterations": iters, "hbm_gbs_per_gpu": work * BYTES_PER_CORR / ms / 1e6 / world,
            "problems_per_gpu": nb, "collective": "none (problems are independent)"}


def measure_small_configs(nlo, syn, ctx):
    """cfg1-cfg3 and cfg5 (scaled to fit the default run time): device-resident GN iterations/s."""
    out = {}
    never = dict(parameter_tolerance=0.0, gradient_tolerance=0.0)
    pose0 = nlo.identity_pose()

    def rate(fn, iters=40, reps=5):
        fn(iters)
        best = min(fn(iters)["device_ms"] for _ in range(reps))
        return iters / (best * 1e-3), best / iters * 1e3

    p, m, s = syn.ndt_problem(100_000, 1001, syn.CFG1_TRUE)
    pr = nlo.NdtProblem(ctx, capacity=len(p)); pr.upload(p, m, s)
    ctx.set_loss(nlo.LOSS_EXPONENTIAL, [1.0, 1.0])
    r, us = rate(lambda k: pr.solve6(pose0, nlo.Options(max_iterations=k, **never)))
    out["cfg1_ndt6_100k"] = {"gn_iterations_per_s": r, "us_per_iteration": us, "points": len(p)}
    pr.close()

    # cfg1 as a whole registration: device matcher (<= 2 nearest cells within 1 m) + Solve, <= 10 rounds
    rng = np.random.default_rng(1001)
    world_pts = syn.room_surface_samples(100_000, rng, 0.01)
    Tinv = np.linalg.inv(syn.CFG1_TRUE)
    local = world_pts @ Tinv[:3, :3].T + Tinv[:3, 3]
    ndt_map = nlo.NdtMap(ctx, grid=syn.room_ndt_grid(0.5))
    best = None
    for _ in range(4):
        t0 = time.perf_counter()
        scan = nlo.Scan(ctx, local)
        reg = scan.register(ndt_map, pose0)
        wall = (time.perf_counter() - t0) * 1e3
        scan.close()
        if best is None or wall < best[0]:
            best = (wall, reg)
    Rr, tr = nlo.pose_to_Rt(best[1]["pose"])
    out["cfg1_registration_100k"] = {
        "wall_ms_incl_scan_upload": best[0], "device_ms": best[1]["device_ms"],
        "outer_iterations": best[1]["outer_iterations"], "inner_iterations": best[1]["inner_iterations"],
        "correspondences": int(best[1]["matched"]),
        "translation_error_m": float(np.linalg.norm(tr - syn.CFG1_TRUE[:3, 3]))}
    ndt_map.close()

    p, m, s = syn.ndt_problem(1_000_000, 1002, syn.CFG2_TRUE)
    pr = nlo.NdtProblem(ctx, capacity=len(p)); pr.upload(p, m, s)
    ctx.set_loss(nlo.LOSS_HUBER, [1.0])
    r, us = rate(lambda k: pr.solve3(pose0, nlo.Options(max_iterations=k, **never)))
    out["cfg2_ndt3_1m_huber"] = {"gn_iterations_per_s": r, "us_per_iteration": us, "points": len(p),
                                 "gpoints_s": len(p) * r / 1e9}
    pr.close()

    # opt-in throughput mode: the cfg4 scan stored as float (48 B per correspondence), fp64 math
    n32 = TOTAL_POINTS
    try:
        pr = nlo.NdtProblem(ctx, capacity=n32, storage="f32")
        pr.generate(n32, SEED, 0, 0.01, syn.to_pose16(syn.CFG1_TRUE), pose0, syn.room_ndt_grid(0.5))
        ctx.set_loss(nlo.LOSS_EXPONENTIAL, [1.0, 1.0])
        r, us = rate(lambda k: pr.solve6(pose0, nlo.Options(max_iterations=k, **never)), iters=20, reps=3)
        out["cfg4_f32_storage_optin"] = {
            "gpoints_s": n32 * r / 1e9, "us_per_iteration": us, "points": n32,
            "hbm_gbs": n32 * 48 * r / 1e9, "bytes_per_correspondence": 48,
            "note": "records (point, mean, S^T S) stored as float, arithmetic in fp64; equals the fp64 "
                    "path on the float-rounded records (tests), ~1e-7 relative quantisation vs the parity mode"}
        # the same mode end to end from pinned FLOAT host arrays (half the PCIe bytes), 16M points
        n_e = 16 * 1024 * 1024
        arr, handle = nlo.host_alloc(n_e * 60)
        f32 = arr.view(np.float32)
        hp, hm, hs = f32[:3 * n_e], f32[3 * n_e:6 * n_e], f32[6 * n_e:15 * n_e]
        chunk = 2 * 1024 * 1024
        pp, mm, ii = pr.download(0, chunk)
        ss = syn.sqrt_info_from_information6(ii)
        for b in range(0, n_e, chunk):
            hp[3 * b:3 * (b + chunk)] = pp.ravel(); hm[3 * b:3 * (b + chunk)] = mm.ravel()
            hs[9 * b:9 * (b + chunk)] = ss.ravel()
        pr.close()
        pe = nlo.NdtProblem(ctx, capacity=n_e, storage="f32")
        opts = nlo.Options(max_iterations=40, **never)

        def one():
            pe.upload_f32_ptr(n_e, hp.ctypes.data, hm.ctypes.data, hs.ctypes.data)
            return pe.solve6(pose0, opts)
        one()
        t0 = time.perf_counter()
        for _ in range(3):
            one()
        dt = (time.perf_counter() - t0) / 3
        out["cfg4_f32_storage_optin"]["e2e_gpoints_s"] = n_e * 40 / dt / 1e9
        out["cfg4_f32_storage_optin"]["e2e_points"] = n_e
        out["cfg4_f32_storage_optin"]["e2e_h2d_bytes_per_solve"] = n_e * 60
        pe.close()
        nlo.host_free(handle)
    except Exception as e:  # never let an optional measurement break the bench line
        out["cfg4_f32_storage_optin"] = {"error": str(e)}

    X, px, K = syn.pnp_problem(50_000, 1003)
    pr = nlo.ReprojProblem(ctx, capacity=len(X)); pr.upload(X, px, K)
    ctx.set_loss(nlo.LOSS_CAUCHY, [1e-2])
    r, us = rate(lambda k: pr.solve(pose0, nlo.Options(max_iterations=k, **never)))
    out["cfg3_pnp_50k_cauchy"] = {"gn_iterations_per_s": r, "us_per_iteration": us, "points": len(X)}
    pr.close()
    return {"secondary": out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--points", type=int, default=TOTAL_POINTS)
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--e2e-points", type=int, default=TOTAL_POINTS)
    ap.add_argument("--e2e-iters", type=int, default=40)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-dropin", action="store_true")
    args = ap.parse_args()
    global GUARD
    GUARD = StdoutGuard()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
