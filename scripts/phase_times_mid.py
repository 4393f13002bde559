import os, sys
os.environ["NLO_DEBUG_TIMES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nonlinear_optimizer_for_slam_b200 as nlo
from nonlinear_optimizer_for_slam_b200 import synthetic as syn
ctx = nlo.Context(0)
grid = syn.room_ndt_grid(0.5)
pose0 = nlo.identity_pose()
never = dict(parameter_tolerance=0.0, gradient_tolerance=0.0)
for n in (300000, 400000, 600000, 800000, 2000000):
    pr = nlo.NdtProblem(ctx, capacity=n)
    pr.generate(n, 1001, 0, 0.01, syn.to_pose16(syn.CFG1_TRUE), pose0, grid)
    ctx.set_loss(nlo.LOSS_EXPONENTIAL, [1.0, 1.0])
    pr.solve6(pose0, nlo.Options(max_iterations=40, **never))
    sys.stderr.write("ndt6 n=%d tiles/CTA=%.2f: " % (n, n / 256 / 296)); sys.stderr.flush()
    r = pr.solve6(pose0, nlo.Options(max_iterations=40, **never))
    sys.stderr.write("   -> %.2f us/iter by events\n" % (r["device_ms"] / 40 * 1e3))
    pr.close()
