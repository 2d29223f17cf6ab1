"""N ranks of the engine in ONE process (run only against the host-sim library, by tests/test_engine_host_sim.py):

    RIO_HOSTSIM_LIBRARY=tests/_build/librio_cuda_hostsim.so python tests/hostsim_multirank.py <world> <solver>

Every rank is a thread with its own provider handle; the peer-memory windows are attached through the same calls a multi-process run
uses (rio_cuda_comm_ipc_export / _attach; the stand-in IPC handle is the pointer itself), and the body is the one of
tests/test_gpu_multi.py: id-range shards, bounded-load calls whose spill rounds fire, calls in flight on several sets, leave / join
events, global counters -- every rank's shard against the oracle run on the GLOBAL key set.  What this covers without a GPU is the
HOST side of the multi-rank path at world sizes the round never had a GPU box for (8): window sizing, exchange epochs and their
ordering across the two kinds of calls, the round decisions every rank must take identically.  The device side of the exchange
(NVLink stores, flags, the fused tail) is proven on the GPU boxes (N = 2 and 4 this round), not here."""
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rio_rs_b200 import _native  # noqa: E402

_native.library_path = lambda: os.environ["RIO_HOSTSIM_LIBRARY"]
_native._lib = None

import rio_rs_b200 as R  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from rio_rs_b200 import parallel  # noqa: E402


def main():
    world, solver = int(sys.argv[1]), sys.argv[2]
    n, M = 240_000, 48
    addrs, seeds, w = O.synth_nodes(M)
    handles, results, errors = [None] * world, [None] * world, []
    bar = threading.Barrier(world)

    def rank_main(rank):
        try:
            p = R.GpuObjectPlacement()
            handles[rank] = p.comm_ipc_export(world)
            bar.wait()
            p.comm_ipc_attach(rank, world, handles)
            bar.wait()
            p.set_nodes(addrs, w)
            p.set_solver(solver)
            lo, hi = parallel.shard_range(n, rank, world)
            s = p.new_set(hi - lo)
            s.synth_keys(lo, hi - lo, 1)
            out = {}
            for cap in [(5, 4), (101, 100), (5, 4), (1, 1)]:
                passes = s.assign_bounded(n, cap[0], cap[1], 4)
                out[cap] = (passes, s.read(), s.counters())
            # two more resident sets, three bounded calls in flight (begin x3, then end x3): exchanges keep their epoch order on every rank
            extra = []
            for k in (2, 3):
                t = p.new_set(hi - lo)
                t.synth_keys(lo, hi - lo, k)
                extra.append(t)
            flight = {}
            for t, cap in zip([s] + extra, [(101, 100), (5, 4), (1, 1)]):
                t.assign_bounded_begin(n, cap[0], cap[1], 4)
            for k, t in enumerate([s] + extra):
                flight[k] = (t.assign_bounded_end(), t.read(), t.counters())
            s.assign()
            p.node_set_active(5, False)
            moved_leave = s.rebalance("leave", 5)
            after_leave = (s.read(), s.counters())
            p.node_set_active(5, True)
            moved_join = s.rebalance("join", 5)
            after_join = (s.read(), s.counters())
            summed = p.comm_sum_counters(np.full(M, rank + 1, dtype=np.uint32))
            results[rank] = (lo, hi, out, flight, moved_leave, after_leave, moved_join, after_join, summed)
            bar.wait()
            del s, extra, t
        except Exception as e:  # noqa: BLE001
            errors.append("rank %d: %r" % (rank, e))
            bar.abort()

    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errors, errors

    def assign(keys, weights):
        return O.assign_hrw2(keys, seeds, weights, threads=8) if solver == "hrw2" else O.assign_hrw(keys, seeds, weights, threads=8)

    bounded = O.assign_bounded_hrw2 if solver == "hrw2" else O.assign_bounded
    keys = O.synth_keys(n, 1)
    for cap in [(5, 4), (101, 100), (1, 1)]:
        widx, wcnt, wpass = bounded(keys, seeds, w, cap[0], cap[1], 4, threads=8)
        for lo, hi, out, *_ in results:
            passes, idx, cnt = out[cap]
            assert passes == wpass and (cnt == wcnt).all() and (idx == widx[lo:hi]).all(), cap
    for k, (seed, cap) in enumerate([(1, (101, 100)), (2, (5, 4)), (3, (1, 1))]):
        widx, wcnt, wpass = bounded(O.synth_keys(n, seed), seeds, w, cap[0], cap[1], 4, threads=8)
        for lo, hi, _, flight, *_ in results:
            passes, idx, cnt = flight[k]
            assert passes == wpass and (cnt == wcnt).all() and (idx == widx[lo:hi]).all(), ("in flight", k)
    base = assign(keys, w)
    w2 = w.copy()
    w2[5] = 0
    left = assign(keys, w2)
    ml = mj = 0
    for lo, hi, _, _, moved_leave, after_leave, moved_join, after_join, summed in results:
        assert (after_leave[0] == left[lo:hi]).all() and (after_join[0] == base[lo:hi]).all()
        assert (after_leave[1] == O.counts(left, M)).all() and (after_join[1] == O.counts(base, M)).all()
        assert (summed == world * (world + 1) // 2).all()
        ml += moved_leave
        mj += moved_join
    assert ml == mj == int((base != left).sum())
    print("multirank ok: world %d solver %s" % (world, solver))


if __name__ == "__main__":
    main()
