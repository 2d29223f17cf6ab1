"""A few HRW2 passes over 10M x 1024 (weights 1..16) on resident keys, for ncu captures (development tool).
usage: python tools/profile_trie.py [bits]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rio_rs_b200 as R
from oracle import pyoracle as O

n, M = 10_000_000, 1024
p = R.GpuObjectPlacement(device=0)
addrs, seeds, w = O.synth_nodes(M)
p.set_nodes(addrs, w)
p.set_solver("hrw2", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
sets = []
for k in range(3):
    s = p.new_set(n)
    s.synth_keys(0, n, 1 + k)
    sets.append(s)
for i in range(6):
    sets[i % 3].assign()
p.sync()
