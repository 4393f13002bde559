import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nonlinear_optimizer_for_slam_b200 as nlo
from nonlinear_optimizer_for_slam_b200 import synthetic as syn
ctx = nlo.Context(0); ctx.set_loss(1, [1.0, 1.0])
n = 32 * 1024 * 1024
pr = nlo.NdtProblem(ctx, capacity=n, storage="f32")
pr.generate(n, 1004, 0, 0.01, syn.to_pose16(syn.CFG1_TRUE), nlo.identity_pose(), syn.room_ndt_grid(0.5))
o = nlo.Options(max_iterations=20, parameter_tolerance=0.0, gradient_tolerance=0.0)
pr.solve6(nlo.identity_pose(), o)
ms = min(pr.solve6(nlo.identity_pose(), o)["device_ms"] for _ in range(3))
print("f32 depth=%s: %.1f us/iter %.2f Gpoints/s" % (os.environ.get("NLO_STAGE_DEPTH", "default"), ms / 20 * 1e3, n * 20 / ms / 1e6))
