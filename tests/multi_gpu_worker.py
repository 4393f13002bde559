"""torchrun worker for tests/test_multi_gpu.py: a scan sharded by point range over the ranks,
28 doubles all-reduced per iteration (peer-memory or NCCL), identical redundant step on each rank.
Rank 0 compares the whole trajectory with the CPU oracle run on the unsharded problem."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import nlo_oracle_py as oracle
    import nonlinear_optimizer_for_slam_b200 as nlo
    from nonlinear_optimizer_for_slam_b200 import sharding, synthetic as syn
    from parity import assert_sums_close, rotation_angle

    comm = sys.argv[1]
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    ctx = nlo.Context(local_rank)
    if comm == "peer":
        handles = [None] * world
        dist.all_gather_object(handles, ctx.comm_peer_export())
        ctx.comm_peer_init(handles, rank, world)
    else:
        uid = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init_nccl(uid[0], rank, world)
    dist.barrier()

    for kind, n, seed, true_T, loss in (("ndt6", 300001, 1004, syn.CFG1_TRUE, (1, [1.0, 1.0])),
                                        ("ndt3", 120000, 1002, syn.CFG2_TRUE, (2, [1.0]))):
        point, mean, S = syn.ndt_problem(n, seed, true_T)
        total = len(point)
        if kind == "ndt3":
            total -= (total % 4 + 1) % 4                # n mod 4 == 3: the dropped tail spans ranks' ends
            point, mean, S = point[:total], mean[:total], S[:total]
        b, e = sharding.point_range(total, rank, world)
        prob = nlo.NdtProblem(ctx, capacity=e - b)
        prob.upload(point[b:e], mean[b:e], S[b:e])
        prob.set_global_range(b, total)                 # the planar tail rule is global
        ctx.set_loss(*loss)
        pose0 = nlo.identity_pose()
        if kind == "ndt6":
            H, g, c = prob.assemble6(pose0)     # all-reduced over ranks
            res = prob.solve6(pose0, trace=True)
            ref = oracle.ndt6_solve(point, mean, S, pose0, *loss)
            Hr, gr, cr = oracle.ndt6_assemble(point, mean, S, np.eye(3), np.zeros(3), *loss, long_double=True)
            nh, ng = 21, 6
        else:
            end_local = max(0, min(e - b, (total // 4) * 4 - b))
            H, g, c = prob.assemble3(pose0, 0, end_local)
            res = prob.solve3(pose0, trace=True)
            ref = oracle.ndt3_solve(point, mean, S, pose0, *loss)
            Hr, gr, cr = oracle.ndt3_assemble(point, mean, S, np.eye(2), np.zeros(2), *loss, long_double=True)
            nh, ng = 6, 3
        assert_sums_close(H, g, c, Hr, gr, cr)
        # the same sums from the ranks' own, un-reduced parts (communicator suspended), added in rank order
        ctx.comm_suspend(True)
        Hl, gl, cl = prob.assemble6(pose0) if kind == "ndt6" else prob.assemble3(pose0, 0, end_local)
        ctx.comm_suspend(False)
        parts = [None] * world
        dist.all_gather_object(parts, np.concatenate([Hl, gl, [cl]]))
        tot = sharding.ordered_sum(parts)
        assert_sums_close(H, g, c, tot[:nh], tot[nh:nh + ng], tot[nh + ng], tol=1e-12)
        pose_r, it_r, cost_r, trace_r = ref
        assert res["iterations"] == it_r, (res["iterations"], it_r)
        for k in range(trace_r.shape[0]):
            a, r = res["trace"][k], trace_r[k]
            assert_sums_close(a[:nh], a[nh:nh + ng], a[nh + ng], r[:nh], r[nh:nh + ng], r[nh + ng])
        Ra, ta = nlo.pose_to_Rt(res["pose"]); Rb, tb = nlo.pose_to_Rt(pose_r)
        assert np.max(np.abs(ta - tb)) < 1e-6 and rotation_angle(Ra, Rb) < 1e-6
        # every rank must hold the bit-identical pose (no broadcast in the protocol)
        poses = [None] * world
        dist.all_gather_object(poses, res["pose"].tobytes())
        assert all(p == poses[0] for p in poses)
        prob.close()
    dist.barrier()
    ctx.comm_destroy()
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_OK", comm, world)


if __name__ == "__main__":
    main()
