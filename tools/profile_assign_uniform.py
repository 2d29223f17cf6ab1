"""One rendezvous pass over 10M x 1024 with UNIFORM weights (one class, 32 full groups) for ncu captures (development tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rio_rs_b200 as R
from oracle import pyoracle as O

n, M = 10_000_000, 1024
p = R.GpuObjectPlacement(device=0)
addrs, seeds, w = O.synth_nodes(M, uniform=True)
p.set_nodes(addrs, w)
s = p.new_set(n)
s.synth_keys(0, n, 1)
for _ in range(4):
    s.assign()
p.sync()
