"""(needs a library built with RIO_BUILD_TUNING=1: the shipped one carries the default tuning point only)
A/B timing of the compiled tuning points of the rendezvous kernel (development tool; run under gpurun)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rio_rs_b200 as R
from oracle import pyoracle as O

n, M = 10_000_000, 1024
p = R.GpuObjectPlacement(device=0)
for uniform in (False, True):
    addrs, seeds, w = O.synth_nodes(M, uniform=uniform)
    p.set_nodes(addrs, w)
    sets = []
    for k in range(3):
        s = p.new_set(n)
        s.synth_keys(0, n, 1 + k)
        sets.append(s)
    want = O.assign_hrw(O.synth_keys(20000, 1), seeds, w, threads=8)
    for variant, tune in [("1", "43c"), ("2", "43a"), ("2", "43b"), ("2", "43c"), ("2", "42a"), ("2", "42c"), ("2", "82a"), ("2", "34c"), ("2", "33c"), ("2", "62c"), ("2", "52c"), ("2", "53c")]:
        os.environ["RIO_ASSIGN_VARIANT"] = variant
        os.environ["RIO_ASSIGN_TUNE"] = tune
        for i in range(3):
            sets[i % 3].assign()
        p.sync()
        p.event_record(0)
        for i in range(12):
            sets[i % 3].assign()
        p.event_record(1)
        p.sync()
        ms = p.event_elapsed_ms(0, 1) / 12
        ok = bool((sets[0].read(0, 20000) == want).all())
        print("weights=%s variant=%s tune=%s: %.3f ms  %.2f Gplacements/s  %.2f Tpairs/s  parity=%s" % ("ones" if uniform else "1..16", variant, tune, ms, n / ms / 1e6, n * M / ms / 1e9, ok), flush=True)
    del sets
print("mix probe: %.2f Tpairs/s" % (max(p.bench_mix_rate(4000) for _ in range(3)) / 1e12))
