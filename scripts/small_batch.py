import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nonlinear_optimizer_for_slam_b200 as nlo
from nonlinear_optimizer_for_slam_b200 import synthetic as syn
ctx = nlo.Context(0); ctx.set_loss(1, [1.0, 1.0])
grid = syn.room_ndt_grid(0.5)
for B, n in ((8, 100000), (16, 20000), (64, 20000), (148, 20000), (149, 20000)):
    pr = nlo.NdtProblem(ctx, counts=[n] * B)
    pr.generate_batched(2000, 0.01, np.tile(syn.to_pose16(syn.CFG1_TRUE), (B, 1)), nlo.identity_pose(), grid)
    opts = nlo.Options(parameter_tolerance=0.0, gradient_tolerance=0.0)
    pr.solve6_batched(np.tile(nlo.identity_pose(), (B, 1)), opts)
    ms = min(pr.solve6_batched(np.tile(nlo.identity_pose(), (B, 1)), opts)["device_ms"] for _ in range(3))
    print("B=%d n=%d: %.3f ms for 40 iterations, %.2f Gpoints/s" % (B, n, ms, B * n * 40 / ms / 1e6))
    pr.close()
