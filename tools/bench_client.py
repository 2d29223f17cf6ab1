"""Client-side first hop (include/rio_client.h): cost per resolve on one host core, and the redirect rate it removes.
CPU only -- run anywhere:  python tools/bench_client.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as O
from rio_rs_b200 import client as CL

try:
    cpu = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
except Exception:
    cpu = "unknown"
print("host: %s, 1 thread" % cpu)
for M in (4, 64, 1024):
    addrs, seeds, w = O.synth_nodes(M)
    for policy in ("hrw", "hrw2"):
        fh = CL.FirstHop(addrs, w, policy=policy)
        n = (2_000_000 // M + 1000) if policy == "hrw" else 2_000_000
        keys = O.synth_keys(n, 1)
        fh.first_hop_batch(keys[:100])
        t = time.perf_counter()
        got = fh.first_hop_batch(keys)
        dt = time.perf_counter() - t
        owner = (O.assign_hrw if policy == "hrw" else O.assign_hrw2)(keys, seeds, w, threads=4)
        wrong = int((got != owner).sum())
        rnd = np.random.default_rng(0).integers(0, M, n)
        t = time.perf_counter()
        CL.FirstHop(addrs, w, policy=policy)
        build = time.perf_counter() - t
        print("M=%4d %-4s: %8.1f ns per first hop, view rebuild %7.1f us, redirects: %d / %d, uniform-random pick (client/mod.rs:254-263) %.1f %%"
              % (M, policy, 1e9 * dt / n, 1e6 * build, wrong, n, 100.0 * float((rnd != owner).mean())))
