"""Kernel-only timing of the two hash-path solvers on resident keys (CUDA events on the engine stream).
usage: python tools/bench_hrw2.py [n] [M]"""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G

G.build()
import rio_rs_b200 as R
from oracle import pyoracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
p = R.GpuObjectPlacement(device=0)
addrs, seeds, w = O.synth_nodes(M)
p.set_nodes(addrs, w)
sets = []
for k in range(4):
    s = p.new_set(n)
    s.synth_keys(0, n, 1 + k)
    sets.append(s)
out = {}
for solver, bits in (("hrw2", 12), ("hrw2", 10), ("hrw2", 14), ("hrw", 0)):
    p.set_solver(solver, bits)
    for i in range(3):
        sets[i % 4].assign()
    p.sync()
    reps = 40 if solver == "hrw2" else 8
    p.event_record(0)
    for i in range(reps):
        sets[i % 4].assign()
    p.event_record(1)
    p.sync()
    ms = p.event_elapsed_ms(0, 1) / reps
    out["%s_bits%d" % (solver, bits)] = {"ms": ms, "placements_per_s": n / ms * 1e3, "GBps_algorithmic": 12 * n / ms / 1e6}
    # rebalance (leave + join of node 17) under this solver
    sets[0].assign()
    p.node_set_active(17, False)
    p.sync(); p.event_record(2)
    moved = sets[0].rebalance("leave", 17)
    p.event_record(3); p.sync()
    t_leave = p.event_elapsed_ms(2, 3)
    p.node_set_active(17, True)
    p.event_record(2)
    moved2 = sets[0].rebalance("join", 17)
    p.event_record(3); p.sync()
    out["%s_bits%d" % (solver, bits)].update({"leave_ms": t_leave, "join_ms": p.event_elapsed_ms(2, 3), "moved": moved, "moved_back": moved2})
print(json.dumps({"n": n, "M": M, "results": out}, indent=1))
