"""Small end-to-end case for compute-sanitizer (memcheck / racecheck) or, where that tool is closed,
for the library's own guard bands (NLO_GUARD=1, include/nlo_cuda.h nlo_debug_guard_report): every
launch shape once, the map builder, the matcher and the outer registration loop; prints the guard
report as the last line (GUARD {json})."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nonlinear_optimizer_for_slam_b200 as nlo
from nonlinear_optimizer_for_slam_b200 import synthetic as syn

ctx = nlo.Context(0)
grid = syn.room_ndt_grid(0.5)
pose0 = nlo.identity_pose()
ctx.set_loss(nlo.LOSS_EXPONENTIAL, [1.0, 1.0])
# persistent cooperative path (grid > 1), resident and streaming tiles
for n in (5000, 400000):
    pr = nlo.NdtProblem(ctx, capacity=n)
    pr.generate(n, 7, 0, 0.01, syn.to_pose16(syn.CFG1_TRUE), pose0, grid)
    print("ndt6", n, pr.solve6(pose0, nlo.Options(max_iterations=6))["iterations"])
    ctx.set_loss(nlo.LOSS_HUBER, [1.0])
    print("ndt3", n, pr.solve3(pose0, nlo.Options(max_iterations=6))["iterations"])
    print("assemble", pr.assemble6(pose0, 3, n - 5)[2])
    ctx.set_loss(nlo.LOSS_EXPONENTIAL, [1.0, 1.0])
    pr.close()
# in-CTA loop (<= 3 tiles) and batched
pr = nlo.NdtProblem(ctx, capacity=600)
pr.generate(600, 8, 0, 0.01, syn.to_pose16(syn.CFG1_TRUE), pose0, grid)
print("tiny", pr.solve6(pose0, nlo.Options(max_iterations=6))["iterations"])
pr.close()
pr = nlo.NdtProblem(ctx, counts=[3000, 100, 2049])
pr.generate_batched(9, 0.01, np.tile(syn.to_pose16(syn.CFG1_TRUE), (3, 1)), pose0, grid)
print("batched", pr.solve6_batched(np.tile(pose0, (3, 1)), nlo.Options(max_iterations=6))["iterations"])
pr.close()
X, px, K = syn.pnp_fixture()
pr = nlo.ReprojProblem(ctx, capacity=len(X)); pr.upload(X, px, K)
print("pnp", pr.solve(pose0)["iterations"])
pr.close()
# device map (dense and hashed), matcher, outer registration loop (6-DoF and planar)
rng = np.random.default_rng(11)
cloud = syn.room_surface_samples(60000, rng, 0.005)
local = syn.room_surface_samples(5000, rng, 0.005)
for hashed in (False, True):
    m = nlo.NdtMap(ctx, points=cloud, voxel=1.0, hashed=hashed)
    sc = nlo.Scan(ctx, local)
    for three in (False, True):
        print("register hashed=%s planar=%s" % (hashed, three),
              sc.register(m, pose0, nlo.Options(max_iterations=5), max_outer=3, three_dof=three)["outer_iterations"])
    sc.close(); m.close()
# sharded over a device list (the same device twice: peer exchange, host rendezvous)
mctx = nlo.Context(devices=[0, 0])
pr = nlo.NdtProblem(mctx, capacity=120001)
pr.generate(120001, 7, 0, 0.01, syn.to_pose16(syn.CFG1_TRUE), pose0, grid)
mctx.set_loss(nlo.LOSS_EXPONENTIAL, [1.0, 1.0])
print("sharded ndt6", pr.solve6(pose0, nlo.Options(max_iterations=6))["iterations"])
mctx.set_loss(nlo.LOSS_HUBER, [1.0])
print("sharded ndt3", pr.solve3(pose0, nlo.Options(max_iterations=6))["iterations"])
pr.close()
mctx.close()
ctx.close()
print("SANITIZE_CASE_DONE")
print("GUARD " + json.dumps(nlo.guard_report()))
