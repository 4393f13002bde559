"""Soak test of the persistent cooperative loop: many solves of random sizes, every result checked
for status 0, bitwise repeatability and the expected iteration count (GPU)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nonlinear_optimizer_for_slam_b200 as nlo
from nonlinear_optimizer_for_slam_b200 import synthetic as syn

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
ctx = nlo.Context(0)
grid = syn.room_ndt_grid(0.5)
pose0 = nlo.identity_pose()
rng = np.random.default_rng(0)
sizes = [300, 2000, 19000, 75000, 100000, 333333, 1000000, 3000000]
probs = []
for n in sizes:
    pr = nlo.NdtProblem(ctx, capacity=n)
    pr.generate(n, 5 + n, 0, 0.01, syn.to_pose16(syn.CFG1_TRUE), pose0, grid)
    probs.append(pr)
ref = {}
t0 = time.time(); count = 0; iters = 0
while time.time() - t0 < seconds:
    k = int(rng.integers(len(sizes)))
    kind = int(rng.integers(2))
    loss = int(rng.integers(4))
    ctx.set_loss(loss, [[0, 0], [1.0, 1.0], [1.0], [0.5]][loss])
    mi = int(rng.integers(1, 41))
    fn = probs[k].solve6 if kind == 0 else probs[k].solve3
    r = fn(pose0, nlo.Options(max_iterations=mi))
    assert r["status"] == 0, (sizes[k], kind, loss, mi, r)
    key = (k, kind, loss, mi)
    sig = (r["iterations"], r["pose"].tobytes(), r["final_cost"])
    if key in ref:
        assert ref[key] == sig, ("not repeatable", key)
    ref[key] = sig
    count += 1; iters += r["iterations"]
print("SOAK_OK solves=%d iterations=%d distinct=%d in %.1f s" % (count, iters, len(ref), time.time() - t0))
