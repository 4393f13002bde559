"""Per-iteration latency of the latency-bound configurations (cfg1 / cfg2 / cfg3 sizes) under the
NLO_* knobs of the environment: one line per run, for sweeps (scripts/cluster_sweep.sh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nonlinear_optimizer_for_slam_b200 as nlo
from nonlinear_optimizer_for_slam_b200 import synthetic as syn


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [20000, 100000, 400000, 1000000]
    ctx = nlo.Context(0)
    never = dict(parameter_tolerance=0.0, gradient_tolerance=0.0)
    pose0 = nlo.identity_pose()
    grid = syn.room_ndt_grid(0.5)
    true16 = syn.to_pose16(syn.CFG1_TRUE)
    out = []
    for n in sizes:
        pr = nlo.NdtProblem(ctx, capacity=n)
        pr.generate(n, 1001, 0, 0.01, true16, pose0, grid)
        for kind, loss, lp in (("ndt6", nlo.LOSS_EXPONENTIAL, [1.0, 1.0]), ("ndt3", nlo.LOSS_HUBER, [1.0])):
            ctx.set_loss(loss, lp)
            fn = pr.solve6 if kind == "ndt6" else pr.solve3
            fn(pose0, nlo.Options(max_iterations=40, **never))
            ms = min(fn(pose0, nlo.Options(max_iterations=40, **never))["device_ms"] for _ in range(9))
            out.append("%s/%d %.2f" % (kind, n, ms / 40 * 1e3))
        pr.close()
    X, px, K = syn.pnp_problem(50000, 1003)
    pr = nlo.ReprojProblem(ctx, capacity=len(X))
    pr.upload(X, px, K)
    ctx.set_loss(nlo.LOSS_CAUCHY, [1e-2])
    pr.solve(pose0, nlo.Options(max_iterations=40, **never))
    ms = min(pr.solve(pose0, nlo.Options(max_iterations=40, **never))["device_ms"] for _ in range(9))
    out.append("pnp/50000 %.2f" % (ms / 40 * 1e3))
    pr.close()
    ctx.close()
    env = " ".join("%s=%s" % (k[4:], v) for k, v in sorted(os.environ.items()) if k.startswith("NLO_"))
    print("[%s] us/iter: %s" % (env, " | ".join(out)), flush=True)


if __name__ == "__main__":
    main()
