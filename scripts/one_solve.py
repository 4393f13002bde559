"""One device-resident solve of a generated NDT problem (profiling target): n, kind, iterations."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nonlinear_optimizer_for_slam_b200 as nlo
from nonlinear_optimizer_for_slam_b200 import synthetic as syn

n = int(sys.argv[1]); kind = sys.argv[2] if len(sys.argv) > 2 else "ndt6"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ctx = nlo.Context(0)
pr = nlo.NdtProblem(ctx, capacity=n)
pr.generate(n, 1001, 0, 0.01, syn.to_pose16(syn.CFG1_TRUE), nlo.identity_pose(), syn.room_ndt_grid(0.5))
ctx.set_loss(nlo.LOSS_EXPONENTIAL, [1.0, 1.0])
fn = pr.solve6 if kind == "ndt6" else pr.solve3
opts = nlo.Options(max_iterations=iters, parameter_tolerance=0.0, gradient_tolerance=0.0)
for _ in range(3):
    r = fn(nlo.identity_pose(), opts)
print("n=%d %s: %.2f us/iter" % (n, kind, r["device_ms"] / iters * 1e3))
