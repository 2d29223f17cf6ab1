"""CPU, world_size 2, gloo: the N>1 host logic -- id-range sharding and the bounded-load round protocol with ONE
counter exchange per pass -- must give exactly the single-process result.  The per-rank engine here is the CPU oracle
(tests may use it); on GPUs the same protocol runs natively inside librio_cuda over NCCL (tests/test_gpu_multi.py)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShardEngine:
    def __init__(self, oracle, keys, seeds, weights):
        self.o, self.keys, self.seeds, self.w = oracle, keys, seeds, weights
        self.idx = np.full(len(keys), 0xFFFFFFFF, dtype=np.uint32)

    def _mask(self, closed):
        m = np.zeros((len(self.seeds) + 31) // 32, dtype=np.uint32)
        for j in closed:
            m[j >> 5] |= np.uint32(1 << (j & 31))
        return m

    def assign(self, closed):
        self.idx = self.o.assign_hrw(self.keys, self.seeds, self.w, mask=self._mask(closed))

    def counts(self):
        return self.o.counts(self.idx, len(self.seeds))

    def spill(self, over, thr, rnd, closed):
        L = self.o.lib()
        sel = [i for i, j in enumerate(self.idx) if j != 0xFFFFFFFF and over[j] and L.orc_spill_hash(int(self.keys[i]), rnd) < int(thr[j])]
        if sel:
            sel = np.array(sel)
            self.idx[sel] = self.o.assign_hrw(self.keys[sel], self.seeds, self.w, mask=self._mask(closed))
        return len(sel)


def _worker(rank, world, port, n, M, cap, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import pyoracle as oracle
    from rio_rs_b200 import parallel

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    _, seeds, w = oracle.synth_nodes(M)
    lo, hi = parallel.shard_range(n, rank, world)
    keys = oracle.synth_keys(hi - lo, 1, first=lo)
    eng = OracleShardEngine(oracle, keys, seeds, w)

    def allreduce(local):
        # the production exchange is an all-gather of M counters per rank followed by a sum
        t = torch.from_numpy(local.astype(np.int64))
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return torch.stack(parts).sum(0).numpy()

    passes = parallel.bounded_assign_protocol(eng, w, n, allreduce, cap[0], cap[1], 4)
    q.put((rank, lo, hi, eng.idx.tolist(), passes))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cap", [(5, 4), (101, 100)])
def test_two_rank_bounded_protocol_equals_single_process(oracle, cap):
    n, M, world = 6000, 24, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + cap[0]
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, M, cap, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _, seeds, w = oracle.synth_nodes(M)
    keys = oracle.synth_keys(n, 1)
    want, _, wpass = oracle.assign_bounded(keys, seeds, w, cap[0], cap[1], 4)
    got = np.empty(n, dtype=np.uint32)
    for rank, lo, hi, idx, passes in res:
        got[lo:hi] = idx
        assert passes == wpass
    assert (got == want).all()
    if cap == (101, 100):
        assert wpass > 1  # the tight cap really exercised the exchange-and-spill rounds


def test_shard_ranges_partition_the_id_space():
    from rio_rs_b200 import parallel

    for n in (0, 1, 7, 10_000_000, 100_000_001):
        for world in (1, 2, 3, 8):
            r = [parallel.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
