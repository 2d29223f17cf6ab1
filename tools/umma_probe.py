"""First-light check of the tcgen05 affinity kernel against the oracle (development tool; run under gpurun + timeout)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rio_rs_b200 as R
from oracle import pyoracle as O

swap = "0"
for (M, n) in [(1024, 100_000), (37, 5001), (300, 20_000)]:
    rng = np.random.default_rng(11)
    fo = rng.uniform(-1, 1, (n, 16)).astype(np.float32)
    fn = np.random.default_rng(13).uniform(-1, 1, (M, 16)).astype(np.float32)
    addrs, _, _ = O.synth_nodes(M)
    w = np.ones(M, dtype=np.uint32)
    w[3] = 0
    p = R.GpuObjectPlacement(device=0)
    p.set_nodes(addrs, w, fn)
    os.environ["RIO_AFFINITY_VARIANT"] = "umma"
    got = p.assign_batch(obj_feats=fo)
    os.environ["RIO_AFFINITY_VARIANT"] = "ffma"
    ref = p.assign_batch(obj_feats=fo)
    idx, cost, gap = O.assign_affinity(fo, fn, w, threads=8)
    tol = 1e-5 * np.abs(cost) + 1e-12
    mism = got != idx
    bad = mism & (gap > tol)
    print("swap=%s M=%d n=%d: umma mismatches vs fp64 oracle %d (outside tolerance %d); ffma mismatches %d; umma==ffma %d/%d" % (
        swap, M, n, int(mism.sum()), int(bad.sum()), int((ref != idx).sum()), int((got == ref).sum()), n), flush=True)
    if bad.sum() == 0 and M == 1024:
        # timing at full size
        N = 10_000_000
        s = p.new_set(N)
        s.synth_keys(0, N, 1)
        big = np.random.default_rng(5).uniform(-1, 1, (N, 16)).astype(np.float32)
        s.load_feats(big)
        for var, ldw in (("umma", "1"), ("umma", "2"), ("ffma", "2")):
            os.environ["RIO_AFFINITY_VARIANT"] = var
            os.environ["RIO_UMMA_LDW"] = ldw
            s.assign(True); p.sync()
            p.event_record(0)
            for _ in range(3):
                s.assign(True)
            p.event_record(1); p.sync()
            ms = p.event_elapsed_ms(0, 1) / 3
            print("  10M x 1024 x K16 %s ldw=%s: %.3f ms  %.2f Gplacements/s  %.1f TFLOP/s(algorithmic 2KM)" % (var, ldw, ms, N / ms / 1e6, 2 * 16 * 1024 * N / ms / 1e9), flush=True)
        del s
