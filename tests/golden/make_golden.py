"""Generates tests/golden/solver_v1.json from the pure-Python spec (tests/spec_py.py).

The reference has no solver, so these vectors pin *this repo's* spec (pair hash at revision v3 -- the file name is the golden FORMAT version; parity unpinned vs rio-rs);
they exist so that neither the C oracle nor the CUDA kernels can drift silently between rounds.
Run: python tests/golden/make_golden.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spec_py as sp  # noqa: E402

g = {}
g["mix64"] = [[str(x), str(sp.mix64(x))] for x in [0, 1, 0xDEADBEEF, 2**63, 2**64 - 1]]
ids = [("obj", "1"), ("Test", "1"), ("test", "1"), ("MockService", "1"), ("Obj", "999999"), ("", "")]
g["object_key"] = [[list(p), str(sp.object_key(*p))] for p in ids]
addrs = ["0.0.0.0:8888", "0.0.0.0:5000", "0.0.0.0:5001", "0.0.0.0:80", "10.0.3.255:5000"]
g["node_seed"] = [[a, str(sp.node_seed(a))] for a in addrs]
us = [0, 1, 2, 3, 255, 256, 65535, 65536, 2**31 - 1, 2**31, 2**31 + 1, 0xDEADBEEF, 2**32 - 2, 2**32 - 1]
g["elog"] = [[u, sp.elog(u)] for u in us]
M = 24
node_addrs = ["10.0.%d.%d:5000" % (j >> 8, j & 255) for j in range(M)]
seeds = [sp.node_seed(a) for a in node_addrs]
weights = [1 + sp.mix64(((j + 1) * 0x9E3779B97F4A7C15 & sp.M64) ^ 7) % 16 for j in range(M)]
weights[5] = 0
keys = [sp.synth_key(i, 1) for i in range(400)]
g["hrw"] = {
    "addresses": node_addrs,
    "keys": [str(k) for k in keys],
    "seeds": [str(s) for s in seeds],
    "weights": weights,
    "idx": [sp.hrw(k, seeds, weights) for k in keys],
}
idx, cnt, passes = sp.assign_bounded(keys, seeds, weights, 21, 20, 4)
g["bounded"] = {"cap": [21, 20], "max_rounds": 4, "idx": idx, "counts": cnt, "passes": passes}
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "solver_v1.json")
json.dump(g, open(out, "w"), indent=0)
print("wrote", out, "passes", passes)
