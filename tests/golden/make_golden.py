"""Writes the golden vectors in this directory from the pinned oracle.

Run from the repo root:  python tests/golden/make_golden.py
The reference itself cannot be built in this image (needs Eigen3 / Ceres / flann / simd_helper),
so the vectors come from oracle/nlo_oracle.cc AFTER it reproduced the reference's PnP known answer
(results/reproj_amd64.txt:5,10) -- see tests/test_oracle_golden.py::test_pnp_known_answer.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import nlo_oracle_py as oracle  # noqa: E402
from nonlinear_optimizer_for_slam_b200 import synthetic as syn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    X, px, K = syn.pnp_fixture()
    pose0 = oracle.pose_from_Rt(np.eye(3), np.zeros(3))
    pose, it, cost, trace = oracle.reproj_solve(X, px, K, pose0, 1, [1.0, 1.0])
    assert it == 6 and "%.5e" % cost == "2.33228e-11"
    np.savez(os.path.join(HERE, "pnp_fixture_trace.npz"), pose=pose, iterations=it, cost=cost, trace=trace)

    point, mean, S = syn.random_ndt_records(500, seed=42)
    R = syn.random_rotation(np.random.default_rng(42))
    t = np.array([0.05, -0.1, 0.2])
    H21, g6, c = oracle.ndt6_assemble(point, mean, S, R, t, 1, [1.0, 1.0])
    H6, g3, c3 = oracle.ndt3_assemble(point, mean, S, R[:2, :2], t[:2], 2, [1.0])
    np.savez(os.path.join(HERE, "ndt_small_sums.npz"), point=point, mean=mean, sqrt_info=S, R=R, t=t,
             H21=H21, g6=g6, cost=c, H6=H6, g3=g3, cost3=c3)


if __name__ == "__main__":
    main()
