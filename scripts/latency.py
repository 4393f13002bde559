"""Per-iteration latency of the device-resident loop for small / mid problems (GPU)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nonlinear_optimizer_for_slam_b200 as nlo
from nonlinear_optimizer_for_slam_b200 import synthetic as syn


def main():
    ctx = nlo.Context(0)
    never = dict(parameter_tolerance=0.0, gradient_tolerance=0.0)
    pose0 = nlo.identity_pose()
    grid = syn.room_ndt_grid(0.5)
    true16 = syn.to_pose16(syn.CFG1_TRUE)
    print("env:", {k: v for k, v in os.environ.items() if k.startswith("NLO_")})
    for n in [2000, 20000, 100000, 400000, 1000000, 4000000, 16000000]:
        pr = nlo.NdtProblem(ctx, capacity=n)
        pr.generate(n, 1001, 0, 0.01, true16, pose0, grid)
        for kind, loss, lp in (("ndt6", nlo.LOSS_EXPONENTIAL, [1.0, 1.0]), ("ndt3", nlo.LOSS_HUBER, [1.0])):
            ctx.set_loss(loss, lp)
            fn = pr.solve6 if kind == "ndt6" else pr.solve3
            for iters in (40,):
                fn(pose0, nlo.Options(max_iterations=iters, **never))
                ms = min(fn(pose0, nlo.Options(max_iterations=iters, **never))["device_ms"] for _ in range(7))
                us = ms / iters * 1e3
                print("%s n=%9d  %8.2f us/iter  %7.2f Gpoints/s  %6.1f GB/s" %
                      (kind, n, us, n / us / 1e3, n * 96 / us / 1e3))
        pr.close()
    X, px, K = syn.pnp_problem(50000, 1003)
    pr = nlo.ReprojProblem(ctx, capacity=len(X)); pr.upload(X, px, K)
    ctx.set_loss(nlo.LOSS_CAUCHY, [1e-2])
    pr.solve(pose0, nlo.Options(max_iterations=40, **never))
    ms = min(pr.solve(pose0, nlo.Options(max_iterations=40, **never))["device_ms"] for _ in range(7))
    print("pnp  n=%9d  %8.2f us/iter" % (len(X), ms / 40 * 1e3))
    pr.close()
    ctx.close()


if __name__ == "__main__":
    main()
