"""Large-N throughput of the reprojection kernel (40 B / correspondence) on one B200."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nonlinear_optimizer_for_slam_b200 as nlo
from nonlinear_optimizer_for_slam_b200 import synthetic as syn

ctx = nlo.Context(0)
never = dict(parameter_tolerance=0.0, gradient_tolerance=0.0)
for n in (4_000_000, 32_000_000):
    X, px, K = syn.pnp_problem(n, 1003)
    pr = nlo.ReprojProblem(ctx, capacity=n)
    pr.upload(X, px, K)
    for name, kind, params in (("none", 0, None), ("Exponential(1,1)", 1, [1.0, 1.0]), ("Cauchy(1e-2)", 3, [1e-2])):
        ctx.set_loss(kind, params)
        pr.solve(nlo.identity_pose(), nlo.Options(max_iterations=20, **never))
        ms = min(pr.solve(nlo.identity_pose(), nlo.Options(max_iterations=20, **never))["device_ms"] for _ in range(3))
        us = ms / 20 * 1e3
        print("pnp n=%d loss=%s: %.1f us/iter  %.1f Gpoints/s  %.0f GB/s (40 B/corr)" % (n, name, us, n / us / 1e3, n * 40 / us / 1e3))
    pr.close()
