"""Measured parity of the CUDA path against the CPU oracle at the BASELINE config sizes (GPU).
Writes a markdown table to stdout (committed as profiles/parity_r1.md)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nlo_oracle_py as oracle
import nonlinear_optimizer_for_slam_b200 as nlo
from nonlinear_optimizer_for_slam_b200 import synthetic as syn
from parity import rel_errors, rotation_angle


def traj(res, ref, nh, ng):
    pose_r, it_r, cost_r, trace_r = ref
    worst = [0.0, 0.0, 0.0]
    rows = min(len(res["trace"]), len(trace_r))
    for k in range(rows):
        a, b = res["trace"][k], trace_r[k]
        e = rel_errors(a[:nh], a[nh:nh + ng], a[nh + ng], b[:nh], b[nh:nh + ng], b[nh + ng])
        worst = [max(w, v) for w, v in zip(worst, e)]
    Ra, ta = nlo.pose_to_Rt(res["pose"]); Rb, tb = nlo.pose_to_Rt(pose_r)
    return (res["iterations"], it_r, worst, float(np.max(np.abs(ta - tb))), rotation_angle(Ra, Rb),
            abs(res["final_cost"] - cost_r) / abs(cost_r))


def main():
    ctx = nlo.Context(0)
    pose0 = nlo.identity_pose()
    print("# Measured parity, CUDA path vs CPU oracle (double), one B200, round 1\n")
    print("Bar (BASELINE.json): per-iteration H, g within 1e-6 relative; pose within 1e-6 m / 1e-6 rad at the "
          "same iteration count.  Error metric: tests/parity.py.  `scripts/parity_report.py`.\n")
    print("## Whole trajectories at the config sizes (every iteration's H, g, cost compared)\n")
    print("| config | correspondences | iterations GPU / CPU | max err H | max err g | max err cost | pose |dt| (m) | pose angle (rad) | final cost rel |")
    print("|---|---|---|---|---|---|---|---|---|")
    p, m, s = syn.ndt_problem(100_000, 1001, syn.CFG1_TRUE)
    pr = nlo.NdtProblem(ctx, capacity=len(p)); pr.upload(p, m, s)
    ctx.set_loss(1, [1.0, 1.0])
    r = traj(pr.solve6(pose0, trace=True), oracle.ndt6_solve(p, m, s, pose0, 1, [1.0, 1.0]), 21, 6)
    print("| cfg1 NDT 6-DoF, Exponential(1,1) | %d | %d / %d | %.1e | %.1e | %.1e | %.1e | %.1e | %.1e |" % (len(p), r[0], r[1], *r[2], r[3], r[4], r[5]))
    pr.close()
    p, m, s = syn.ndt_problem(1_000_000, 1002, syn.CFG2_TRUE)
    pr = nlo.NdtProblem(ctx, capacity=len(p)); pr.upload(p, m, s)
    ctx.set_loss(2, [1.0])
    r = traj(pr.solve3(pose0, trace=True), oracle.ndt3_solve(p, m, s, pose0, 2, [1.0]), 6, 3)
    print("| cfg2 NDT 3-DoF, Huber(1) | %d | %d / %d | %.1e | %.1e | %.1e | %.1e | %.1e | %.1e |" % (len(p), r[0], r[1], *r[2], r[3], r[4], r[5]))
    pr.close()
    X, px, K = syn.pnp_problem(50_000, 1003)
    rp = nlo.ReprojProblem(ctx, capacity=len(X)); rp.upload(X, px, K)
    ctx.set_loss(3, [1e-2])
    r = traj(rp.solve(pose0, trace=True), oracle.reproj_solve(X, px, K, pose0, 3, [1e-2]), 21, 6)
    print("| cfg3 PnP, Cauchy(1e-2), 5 %% outliers | %d | %d / %d | %.1e | %.1e | %.1e | %.1e | %.1e | %.1e |" % (len(X), r[0], r[1], *r[2], r[3], r[4], r[5]))
    rp.close()
    X, px, K = syn.pnp_fixture()
    rp = nlo.ReprojProblem(ctx, capacity=len(X)); rp.upload(X, px, K)
    ctx.set_loss(1, [1.0, 1.0])
    res = rp.solve(pose0, trace=True)
    r = traj(res, oracle.reproj_solve(X, px, K, pose0, 1, [1.0, 1.0]), 21, 6)
    print("| reference PnP fixture (630 points), COST %.5e (published 2.33228e-11) | %d | %d / %d | %.1e | %.1e | %.1e | %.1e | %.1e | %.1e |" % (res["final_cost"], len(X), r[0], r[1], *r[2], r[3], r[4], r[5]))
    rp.close()

    print("\n## One assembly pass vs the long-double oracle sum, worst over 5 random poses\n")
    print("| kind | loss | correspondences | max err H | max err g | max err cost |")
    print("|---|---|---|---|---|---|")
    rng = np.random.default_rng(0)
    names = {0: "none", 1: "Exponential(1,1)", 2: "Huber(1)", 3: "Cauchy(0.5)"}
    params = {0: None, 1: [1.0, 1.0], 2: [1.0], 3: [0.5]}
    n = 300_000
    point, mean, S = syn.ndt_problem(n, 77, syn.CFG1_TRUE)
    pr = nlo.NdtProblem(ctx, capacity=len(point)); pr.upload(point, mean, S)
    Xp, pxp, Kp = syn.pnp_problem(n, 78)
    rp = nlo.ReprojProblem(ctx, capacity=n); rp.upload(Xp, pxp, Kp)
    for kind in (0, 1, 2, 3):
        ctx.set_loss(kind, params[kind])
        worst6 = [0, 0, 0]; worst3 = [0, 0, 0]; worstp = [0, 0, 0]
        for _ in range(5):
            R = syn.random_rotation(rng, 0.2); t = rng.uniform(-0.3, 0.3, 3)
            Rq = oracle.quat_to_rotmat(oracle.rotmat_to_quat(R))
            pose = nlo.pose_from_Rt(R, t)
            H, g, c = pr.assemble6(pose)
            e = rel_errors(H, g, c, *oracle.ndt6_assemble(point, mean, S, Rq, t, kind, params[kind], long_double=True))
            worst6 = [max(a, b) for a, b in zip(worst6, e)]
            yaw = rng.uniform(-0.3, 0.3); T = syn.yaw_pose([t[0], t[1], 0.0], yaw)
            H, g, c = pr.assemble3(syn.to_pose16(T))
            e = rel_errors(H, g, c, *oracle.ndt3_assemble(point, mean, S, T[:2, :2], T[:2, 3], kind, params[kind], long_double=True))
            worst3 = [max(a, b) for a, b in zip(worst3, e)]
            H, g, c = rp.assemble(pose)
            e = rel_errors(H, g, c, *oracle.reproj_assemble(Xp, pxp, Kp, Rq, t, kind, params[kind], long_double=True))
            worstp = [max(a, b) for a, b in zip(worstp, e)]
        print("| NDT 6-DoF | %s | %d | %.1e | %.1e | %.1e |" % (names[kind], len(point), *worst6))
        print("| NDT 3-DoF | %s | %d | %.1e | %.1e | %.1e |" % (names[kind], (len(point) // 4) * 4, *worst3))
        print("| reprojection | %s | %d | %.1e | %.1e | %.1e |" % (names[kind], n, *worstp))
    pr.close(); rp.close(); ctx.close()


if __name__ == "__main__":
    main()
