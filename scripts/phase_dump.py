"""Dump of every CTA's phase stamps (NLO_DEBUG_TIMES=2) for one size, for offline analysis:
python scripts/phase_dump.py N ndt6|ndt3|pnp out.bin   (see scripts/phase_analyze.py)"""
import os
import sys
os.environ["NLO_DEBUG_TIMES"] = "2"
os.environ["NLO_DEBUG_FILE"] = sys.argv[3]
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nonlinear_optimizer_for_slam_b200 as nlo
from nonlinear_optimizer_for_slam_b200 import synthetic as syn
n, kind = int(sys.argv[1]), sys.argv[2]
ctx = nlo.Context(0)
pose0 = nlo.identity_pose()
never = dict(parameter_tolerance=0.0, gradient_tolerance=0.0)
if kind == "pnp":
    X, px, K = syn.pnp_problem(n, 1003)
    pr = nlo.ReprojProblem(ctx, capacity=len(X))
    pr.upload(X, px, K)
    ctx.set_loss(nlo.LOSS_CAUCHY, [1e-2])
    fn = pr.solve
else:
    grid = syn.room_ndt_grid(0.5)
    pr = nlo.NdtProblem(ctx, capacity=n)
    pr.generate(n, 1001, 0, 0.01, syn.to_pose16(syn.CFG1_TRUE), pose0, grid)
    ctx.set_loss(nlo.LOSS_EXPONENTIAL, [1.0, 1.0])
    fn = pr.solve6 if kind == "ndt6" else pr.solve3
fn(pose0, nlo.Options(max_iterations=40, **never))
r = fn(pose0, nlo.Options(max_iterations=40, **never))
print("%s n=%d %.2f us/iter" % (kind, n, r["device_ms"] / 40 * 1e3))
