"""Dense grid vs voxel hash: map build and matcher cost on the same inputs.
Room of the reference's fixture (954 605 points) at 1.0 / 0.5 / 0.25 m voxels; a 1M-point scan matched
(<= 2 nearest means within 1 m) against each form.  Wall times around the blocking C-ABI calls."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nonlinear_optimizer_for_slam_b200 as nlo
from nonlinear_optimizer_for_slam_b200 import synthetic as syn


def best_of(fn, reps=5):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        dt = (time.perf_counter() - t0) * 1e3
        best = dt if best is None or dt < best else best
    return best, out


def main():
    ctx = nlo.Context(0)
    room = syn.room_points()
    rng = np.random.default_rng(3)
    world = syn.room_surface_samples(1 << 20, rng, 0.02)
    Tinv = np.linalg.inv(syn.CFG1_TRUE)
    local = world @ Tinv[:3, :3].T + Tinv[:3, 3]
    scan = nlo.Scan(ctx, local)
    prob = nlo.NdtProblem(ctx, capacity=2 * len(local))
    pose = syn.to_pose16(syn.yaw_pose([0.03, -0.02, 0.05], 0.02))
    rows = []
    for voxel in (1.0, 0.5, 0.25):
        row = {"voxel_m": voxel, "map_points": len(room), "scan_points": len(local)}
        for hashed in (False, True):
            maps = []

            def build():
                m = nlo.NdtMap(ctx, points=room, voxel=voxel, hashed=hashed)
                maps.append(m)
                return m

            build_ms, m = best_of(build, 3)
            _, rows_in_table = m.layout()
            radius = min(1.0, 2 * voxel)
            match_ms, matched = best_of(lambda: scan.match(m, pose, prob, radius=radius))
            tag = "hashed" if hashed else "dense"
            row[tag] = {"build_ms": build_ms, "table_rows": rows_in_table, "match_ms": match_ms,
                        "matched": matched, "radius_m": radius}
            for x in maps:
                x.close()
        assert row["dense"]["matched"] == row["hashed"]["matched"]
        rows.append(row)
    print(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
