"""Multi-GPU host logic: one process per GPU, objects sharded by id-range, node table replicated, and ONE
collective per assignment pass -- the all-gather (+sum) of the per-node load counters.

The production exchange runs inside librio_cuda (ncclAllGather on the engine's stream, see
rio_cuda_comm_init); torch.distributed is only the bootstrap plumbing that ships the NCCL unique id and
reduces the timings.  `bounded_assign_protocol` is the same round protocol written against an abstract
per-rank engine so that it can be exercised with the gloo backend on CPU (tests/test_parallel_gloo.py) and
cross-checked against the native loop on GPUs.
"""
import numpy as np

NONE = 0xFFFFFFFF


def shard_range(n_total, rank, world):
    """Contiguous id-range [lo, hi) of rank `rank`: floor split, identical on every rank."""
    lo = n_total * rank // world
    hi = n_total * (rank + 1) // world
    return lo, hi


def capacity(n_total, w, w_sum, cap_num, cap_den):
    """ceil(cap_num * n_total * w / (cap_den * w_sum)) clamped to u32."""
    if not w or not w_sum or not cap_den:
        return 0
    return min(0xFFFFFFFF, -(-(cap_num * n_total * int(w)) // (cap_den * int(w_sum))))


def init_comm(provider, dist):
    """Attach an NCCL communicator to `provider` using torch.distributed `dist` as the bootstrap."""
    from .provider import comm_unique_id

    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    provider.comm_init(rank, world, box[0])          # NCCL communicator: the portable path
    import os

    if os.environ.get("RIO_COMM", "p2p") == "p2p" and world <= 16:
        # peer-memory windows over NVLink (CUDA IPC): the exchange becomes one kernel, no NCCL launch per pass
        mine = provider.comm_ipc_export(world)
        handles = [None] * world
        dist.all_gather_object(handles, mine)
        provider.comm_ipc_attach(rank, world, handles)
        dist.barrier()


def bounded_assign_protocol(engine, weights, n_total, allreduce_counts, cap_num=5, cap_den=4, max_rounds=4):
    """Bounded-load rounds over one rank's shard.

    engine must offer:
        assign(closed: set[int]) -> None                    (re)assign every object of the shard over live - closed
        counts() -> np.ndarray[u32, M]                      local per-node counts
        spill(over: np.ndarray[bool], thr: np.ndarray[u32], round: int, closed: set[int]) -> int
                                                            re-place the spilling objects, return how many
    allreduce_counts(local) -> global (the single collective of a pass).
    Returns the number of assignment passes; identical on every rank because every decision is taken on the
    GLOBAL counters.
    """
    weights = np.asarray(weights, dtype=np.uint64)
    M = len(weights)
    w_sum = int(weights.sum())
    cap = np.array([capacity(n_total, int(w), w_sum, cap_num, cap_den) for w in weights], dtype=np.uint64)
    closed = set()
    engine.assign(closed)
    passes = 1
    for r in range(1, max_rounds):
        glob = np.asarray(allreduce_counts(engine.counts()), dtype=np.uint64)
        over = (weights > 0) & (glob > cap)
        for j in np.nonzero(over)[0]:
            closed.add(int(j))
        open_ = sum(1 for j in range(M) if weights[j] and j not in closed)
        if not over.any() or not open_:
            break
        thr = np.zeros(M, dtype=np.uint32)
        o = np.nonzero(over)[0]
        thr[o] = [int(((int(glob[j]) - int(cap[j])) << 32) // int(glob[j])) for j in o]
        engine.spill(over, thr, r, closed)
        passes += 1
    return passes
