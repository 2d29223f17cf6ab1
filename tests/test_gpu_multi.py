"""Multi-GPU parity (needs >= 2 GPUs: run under `gpurun --gpus 2`): id-range shards on two ranks, the native
bounded-load loop with its NCCL counter all-gather, and sharded rebalance events must reproduce the single-process
oracle bit for bit.  Skipped when fewer than two GPUs are visible."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, M, q, comm, solver):
    sys.path.insert(0, ROOT)
    os.environ["RIO_COMM"] = comm   # "p2p": CUDA-IPC windows over NVLink (default), "nccl": ncclAllGather
    import torch.distributed as dist

    import rio_rs_b200 as R
    from oracle import pyoracle as O
    from rio_rs_b200 import parallel

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)  # bootstrap only
    p = R.GpuObjectPlacement(device=rank)
    parallel.init_comm(p, dist)
    addrs, seeds, w = O.synth_nodes(M)
    p.set_nodes(addrs, w)
    p.set_solver(solver)
    lo, hi = parallel.shard_range(n, rank, world)
    s = p.new_set(hi - lo)
    s.synth_keys(lo, hi - lo, 1)
    out = {}
    for cap in [(5, 4), (101, 100), (5, 4), (1, 1)]:   # repeated factors: the two counter buffers of a set take turns across calls
        passes = s.assign_bounded(n, cap[0], cap[1], 4)
        out[cap] = (passes, s.read().tolist(), s.counters().tolist())
    # plain assignment + a leave and a join, counters are global after the exchange
    s.assign()
    p.node_set_active(5, False)
    moved_leave = s.rebalance("leave", 5)
    after_leave = (s.read().tolist(), s.counters().tolist())
    p.node_set_active(5, True)
    moved_join = s.rebalance("join", 5)
    after_join = (s.read().tolist(), s.counters().tolist())
    summed = p.comm_sum_counters(np.full(M, rank + 1, dtype=np.uint32)).tolist()
    q.put((rank, lo, hi, out, moved_leave, after_leave, moved_join, after_join, summed))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("solver", ["hrw", "hrw2"])
@pytest.mark.parametrize("comm", ["p2p", "nccl"])
def test_two_gpu_sharded_results_equal_single_process_oracle(oracle, comm, solver):
    """Both exchange paths x both solver policies.  Under hrw2 + p2p the whole pass (walk, histogram, exchange over NVLink peer
    memory, capacity check) is ONE kernel per rank, the last CTA of each rank's walk spinning on its peers' flags."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from rio_rs_b200 import build

    build.build()
    n, M, world = 400_000, 48, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000 + (7 if comm == "nccl" else 0) + (13 if solver == "hrw2" else 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, M, q, comm, solver)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    _, seeds, w = oracle.synth_nodes(M)
    keys = oracle.synth_keys(n, 1)
    def assign(weights):
        return oracle.assign_hrw2(keys, seeds, weights, threads=8) if solver == "hrw2" else oracle.assign_hrw(keys, seeds, weights, threads=8)

    for cap in [(5, 4), (101, 100), (1, 1)]:
        widx, wcnt, wpass = (oracle.assign_bounded_hrw2 if solver == "hrw2" else oracle.assign_bounded)(keys, seeds, w, cap[0], cap[1], 4, threads=8)
        got = np.empty(n, dtype=np.uint32)
        for rank, lo, hi, out, *_ in res:
            passes, idx, cnt = out[cap]
            got[lo:hi] = idx
            assert passes == wpass and cnt == wcnt.tolist()   # counters are the GLOBAL ones on every rank
        assert (got == widx).all(), cap
    base = assign(w)
    w2 = w.copy()
    w2[5] = 0
    left = assign(w2)
    got_l, got_j = np.empty(n, dtype=np.uint32), np.empty(n, dtype=np.uint32)
    ml = mj = 0
    for rank, lo, hi, _, moved_leave, after_leave, moved_join, after_join, summed in res:
        got_l[lo:hi] = after_leave[0]
        got_j[lo:hi] = after_join[0]
        ml += moved_leave
        mj += moved_join
        assert after_leave[1] == oracle.counts(left, M).tolist() and after_join[1] == oracle.counts(base, M).tolist()
        assert summed == [3] * M   # (rank0: 1) + (rank1: 2)
    assert (got_l == left).all() and (got_j == base).all()
    assert ml == mj == int((base != left).sum())
