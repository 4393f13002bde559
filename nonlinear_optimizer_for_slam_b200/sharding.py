"""Host-side partitioning for the multi-GPU paths (one process per GPU).

Large scan (BASELINE cfg4): contiguous point ranges, one per rank; every iteration each rank
reduces its 28 (10 for 3-DoF) doubles locally, the ranks all-reduce them and every rank applies
the identical damped step (no broadcast of the pose is needed).  Batched registrations (cfg5):
block partition of the problem ids, no collective at all.
"""


def point_range(total, rank, world):
    """[begin, end) of the scan owned by `rank`: floor(total/world) points each, the remainder on
    the last rank.  Ranges are contiguous, disjoint and cover [0, total)."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad (total, rank, world) = (%r, %r, %r)" % (total, rank, world))
    per = total // world
    begin = rank * per
    end = total if rank == world - 1 else begin + per
    return begin, end


def problem_partition(num_problems, rank, world):
    """[begin, end) of the registration ids owned by `rank` (block partition, sizes differ by <= 1)."""
    if world < 1 or not (0 <= rank < world) or num_problems < 0:
        raise ValueError("bad (num_problems, rank, world)")
    base, extra = divmod(num_problems, world)
    begin = rank * base + min(rank, extra)
    end = begin + base + (1 if rank < extra else 0)
    return begin, end


def ordered_sum(per_rank_vectors):
    """Sum in rank order 0..n-1 -- the order the peer-memory all-reduce uses on every rank, which
    is what makes the reduced H|g|cost bit-identical everywhere."""
    total = None
    for v in per_rank_vectors:
        total = v.copy() if total is None else total + v
    return total
