import os, sys
os.environ["NLO_DEBUG_TIMES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nonlinear_optimizer_for_slam_b200 as nlo
from nonlinear_optimizer_for_slam_b200 import synthetic as syn
ctx = nlo.Context(0)
grid = syn.room_ndt_grid(0.5)
pose0 = nlo.identity_pose()
never = dict(parameter_tolerance=0.0, gradient_tolerance=0.0)
for n in (600, 20000, 100000, 1000000, 8000000):
    pr = nlo.NdtProblem(ctx, capacity=n)
    pr.generate(n, 1001, 0, 0.01, syn.to_pose16(syn.CFG1_TRUE), pose0, grid)
    for kind in ("ndt6", "ndt3"):
        ctx.set_loss(nlo.LOSS_EXPONENTIAL, [1.0, 1.0])
        fn = pr.solve6 if kind == "ndt6" else pr.solve3
        fn(pose0, nlo.Options(max_iterations=40, **never))
        sys.stderr.write("%s n=%d: " % (kind, n)); sys.stderr.flush()
        r = fn(pose0, nlo.Options(max_iterations=40, **never))
        sys.stderr.write("   -> %.2f us/iter by events\n" % (r["device_ms"] / 40 * 1e3))
    pr.close()
