"""Generates the golden vectors under tests/golden/ from the pure-Python spec (tests/spec_py.py).

The reference has no solver, so these vectors pin *this repo's* spec (parity unpinned vs rio-rs); they exist so that neither
the C oracle nor the CUDA kernels can drift silently between rounds:

  solver_hrw_v3.json   flat weighted rendezvous, pair hash at revision v3.  FROZEN: round 1 shipped
                       it (under the name solver_v1.json); tests/test_oracle_spec.py fails if this script would write
                       anything else than the committed file.
  solver_hrw2_v1.json  HRW2, the hierarchical policy with fan-out 2, revision 1.  Frozen the same way.

Run: python tests/golden/make_golden.py     (rewrites both files; a diff in git means the spec changed)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spec_py as sp  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = {"flat": "solver_hrw_v3.json", "hrw2": "solver_hrw2_v1.json"}


def _table(M=24):
    node_addrs = ["10.0.%d.%d:5000" % (j >> 8, j & 255) for j in range(M)]
    seeds = [sp.node_seed(a) for a in node_addrs]
    weights = [1 + sp.mix64(((j + 1) * 0x9E3779B97F4A7C15 & sp.M64) ^ 7) % 16 for j in range(M)]
    weights[5] = 0
    return node_addrs, seeds, weights


def build_flat():
    g = {}
    g["mix64"] = [[str(x), str(sp.mix64(x))] for x in [0, 1, 0xDEADBEEF, 2**63, 2**64 - 1]]
    ids = [("obj", "1"), ("Test", "1"), ("test", "1"), ("MockService", "1"), ("Obj", "999999"), ("", "")]
    g["object_key"] = [[list(p), str(sp.object_key(*p))] for p in ids]
    addrs = ["0.0.0.0:8888", "0.0.0.0:5000", "0.0.0.0:5001", "0.0.0.0:80", "10.0.3.255:5000"]
    g["node_seed"] = [[a, str(sp.node_seed(a))] for a in addrs]
    us = [0, 1, 2, 3, 255, 256, 65535, 65536, 2**31 - 1, 2**31, 2**31 + 1, 0xDEADBEEF, 2**32 - 2, 2**32 - 1]
    g["elog"] = [[u, sp.elog(u)] for u in us]
    node_addrs, seeds, weights = _table()
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
    return g


def build_hrw2():
    g = {"spec": "HRW2 revision 1 (DESIGN.md 3.8)"}
    g["level_seed"] = [[l, str(sp.hrw2_level_seed(l))] for l in (0, 1, 11, 12, 15, 40)]
    g["contest_v"] = [[str(k), str(s), sp.hrw2_v(k, s)] for k, s in [(0, 0), (1, 2), (sp.synth_key(0, 1), sp.node_seed("0.0.0.0:8888")),
                                                                      (2**64 - 1, sp.hrw2_level_seed(0)), (sp.synth_key(5, 3), sp.hrw2_level_seed(7))]]
    g["threshold"] = [[wl, wr, ((wl << 31) // (wl + wr)) if wl + wr else 0] for wl, wr in [(0, 0), (0, 5), (5, 0), (1, 1), (1, 2), (7, 3), (2**32 - 1, 1), (1, 2**40)]]
    node_addrs, seeds, weights = _table()
    keys = [sp.synth_key(i, 1) for i in range(400)]
    g["addresses"], g["seeds"], g["weights"] = node_addrs, [str(s) for s in seeds], weights
    g["keys"] = [str(k) for k in keys]
    g["idx"] = {str(bits): [sp.hrw2(k, seeds, weights, (), bits) for k in keys] for bits in (12, 3, 1)}   # bits 3 and 1: chains inside the buckets
    idx, cnt, passes = sp.assign_bounded_hrw2(keys, seeds, weights, 21, 20, 4, 12)
    g["bounded"] = {"bits": 12, "cap": [21, 20], "max_rounds": 4, "idx": idx, "counts": cnt, "passes": passes}
    return g


BUILDERS = {"flat": build_flat, "hrw2": build_hrw2}


def render(g):
    return json.dumps(g, indent=0)


if __name__ == "__main__":
    for name, fn in BUILDERS.items():
        out = os.path.join(HERE, FILES[name])
        with open(out, "w") as f:
            f.write(render(fn()))
        print("wrote", out)
