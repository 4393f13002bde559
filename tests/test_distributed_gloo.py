"""CPU, world_size 2 (gloo): the sharded Gauss-Newton loop.  The per-rank partial sums come from
the oracle standing in for the kernel; what is under test is the host-side partitioning and the
"all-reduce 28 doubles, step redundantly on every rank" protocol of SURVEY.md section 8(e)."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import nlo_oracle_py as oracle
    from nonlinear_optimizer_for_slam_b200 import sharding, synthetic as syn
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    point, mean, S = syn.ndt_problem(20001, 1004, syn.CFG1_TRUE)
    n = len(point)
    b, e = sharding.point_range(n, rank, world)
    state = np.zeros(9)
    state[6] = 1.0              # identity quaternion (x, y, z, w)
    state[7] = 0.001            # lambda
    state[8] = np.finfo(np.float64).max
    iteration = 0
    for iteration in range(40):
        R = oracle.quat_to_rotmat(state[3:7])
        H, g, c = oracle.ndt6_assemble(point, mean, S, R, state[:3], 1, [1.0, 1.0], begin=b, end=e)
        local = np.concatenate([H, g, [c]])
        gathered = [torch.zeros(28, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(local))
        total = sharding.ordered_sum([t.numpy() for t in gathered])     # rank order on every rank
        conv, state = oracle.gn6_step(total[:21], total[21:27], total[27], state)
        if conv:
            break
    else:
        iteration = 40
    np.save(os.path.join(tmpdir, "state_%d.npy" % rank), np.concatenate([state, [iteration]]))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_loop_is_identical_on_all_ranks_and_matches_single(tmp_path, oracle):
    from nonlinear_optimizer_for_slam_b200 import synthetic as syn
    world, port = 2, 29631
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    s0 = np.load(tmp_path / "state_0.npy")
    s1 = np.load(tmp_path / "state_1.npy")
    assert np.array_equal(s0, s1)          # bit-identical redundant state, no pose broadcast needed
    point, mean, S = syn.ndt_problem(20001, 1004, syn.CFG1_TRUE)
    pose, it, cost, _ = oracle.ndt6_solve(point, mean, S, oracle.pose_from_Rt(np.eye(3), np.zeros(3)),
                                          1, [1.0, 1.0])
    assert int(s0[9]) == it
    R, t = oracle.pose_to_Rt(pose)
    np.testing.assert_allclose(s0[:3], t, atol=1e-9)
    np.testing.assert_allclose(oracle.quat_to_rotmat(s0[3:7]), R, atol=1e-9)


def test_partitions():
    from nonlinear_optimizer_for_slam_b200 import sharding
    for total in (0, 1, 7, 64 * 1024 * 1024, 1000003):
        for world in (1, 2, 3, 4, 8):
            cuts = [sharding.point_range(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(cuts[k][1] == cuts[k + 1][0] for k in range(world - 1))
            parts = [sharding.problem_partition(total % 5000, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total % 5000
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.point_range(10, 2, 2)
