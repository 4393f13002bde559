"""Per-iteration latency of the planar (3-DoF) loop only, Huber(1) (GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nonlinear_optimizer_for_slam_b200 as nlo
from nonlinear_optimizer_for_slam_b200 import synthetic as syn

ctx = nlo.Context(0)
ctx.set_loss(nlo.LOSS_HUBER, [1.0])
never = dict(parameter_tolerance=0.0, gradient_tolerance=0.0)
pose0 = nlo.identity_pose()
grid = syn.room_ndt_grid(0.5)
for n in [20000, 100000, 400000, 1000000, 4000000, 16000000]:
    pr = nlo.NdtProblem(ctx, capacity=n)
    pr.generate(n, 1001, 0, 0.01, syn.to_pose16(syn.CFG1_TRUE), pose0, grid)
    o = nlo.Options(max_iterations=40, **never)
    pr.solve3(pose0, o)
    ms = min(pr.solve3(pose0, o)["device_ms"] for _ in range(9))
    print("ndt3 n=%9d  %8.2f us/iter  %7.2f Gpoints/s" % (n, ms / 40 * 1e3, n * 40 / ms / 1e6))
    pr.close()
