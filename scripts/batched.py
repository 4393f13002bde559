"""cfg5: batched independent NDT 6-DoF registrations, one CTA per registration (GPU)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nonlinear_optimizer_for_slam_b200 as nlo
from nonlinear_optimizer_for_slam_b200 import synthetic as syn


def cfg5_true_poses(num, seed=2000):
    rng = np.random.default_rng(seed)
    poses = np.zeros((num, 16))
    for k in range(num):
        T = syn.yaw_pose(rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.15, 0.15))
        poses[k] = syn.to_pose16(T)
    return poses


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    ctx = nlo.Context(0)
    ctx.set_loss(nlo.LOSS_EXPONENTIAL, [1.0, 1.0])
    grid = syn.room_ndt_grid(0.5)
    prob = nlo.NdtProblem(ctx, counts=[n] * B)
    prob.generate_batched(2000, 0.01, cfg5_true_poses(B), nlo.identity_pose(), grid)
    poses0 = np.tile(nlo.identity_pose(), (B, 1))
    for label, opts in (("converge (tol 1e-6)", nlo.Options()),
                        ("fixed 40 iterations", nlo.Options(parameter_tolerance=0.0, gradient_tolerance=0.0))):
        prob.solve6_batched(poses0, opts)
        t0 = time.perf_counter()
        out = prob.solve6_batched(poses0, opts)
        wall = time.perf_counter() - t0
        iters = out["iterations"].astype(np.int64)
        work = int(np.sum(np.minimum(iters + 1, opts.max_iterations))) * n  # passes actually executed
        ms = out["device_ms"]
        print("%s: B=%d n=%d  device %.3f ms (wall %.3f ms)  iterations min/mean/max %d/%.1f/%d  "
              "%.2f Gpoints/s  %.0f GB/s  %.0f registrations/s"
              % (label, B, n, ms, wall * 1e3, iters.min(), iters.mean(), iters.max(),
                 work / ms / 1e6, work * 96 / ms / 1e6, B / ms * 1e3))
    prob.close(); ctx.close()


if __name__ == "__main__":
    main()
