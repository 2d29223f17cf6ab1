"""Statistical evidence for the HRW2 spec, CPU only (the oracle): chi-square of the node loads against w/W,
the spread of a leaving node's objects over the survivors, and the movement on leave / join relative to the minimal set, for
uniform and 1..16 weights and several trie depths.  usage: python tools/hrw2_quality.py [objects] > profiles/r02_hrw2_quality.txt"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as O

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
M = 1024
T = os.cpu_count() or 8
print("HRW2 quality, %d objects x %d nodes, oracle/rio_oracle.c orc_assign_hrw2, %d threads" % (N, M, T))
keys = O.synth_keys(N, 1)
for uniform in (False, True):
    addrs, seeds, w = O.synth_nodes(M, uniform=uniform)
    for bits in (12, 10, 14, 4):
        t = time.time()
        a = O.assign_hrw2(keys, seeds, w, bits=bits, threads=T)
        dt = time.time() - t
        cnt = np.bincount(a, minlength=M).astype(np.float64)
        e = N * w / w.sum()
        chi = ((cnt - e) ** 2 / e).sum()
        w2 = w.copy()
        w2[17] = 0
        b = O.assign_hrw2(keys, seeds, w2, bits=bits, threads=T)
        moved, minimal = int((a != b).sum()), int((a == 17).sum())
        # the loads AFTER the leave against w/(W - w17): the leaver's objects land in its sibling subtree and the ancestors' slivers
        # compensate, so the end state is proportional again (that compensation is the extra movement)
        cb = np.bincount(b, minlength=M).astype(np.float64)
        eb = N * w2 / w2.sum()
        live = w2 > 0
        chi_d = ((cb[live] - eb[live]) ** 2 / eb[live]).sum()
        # a join of a brand-new node (index M) with weight 8
        seeds3 = np.concatenate([seeds, np.array([O.node_seed("10.9.9.9:5000")], dtype=np.uint64)])
        w3 = np.concatenate([w, np.array([8], dtype=np.uint32)])
        c = O.assign_hrw2(keys, seeds3, w3, bits=bits, threads=T)
        jm, jmin = int((a != c).sum()), int((c == M).sum())
        print("weights %-7s bits %2d: chi2(loads) %.0f (df %d, sigma %.0f) max/mean %.4f min/mean %.4f | leave(17): moved %d minimal %d ratio %.2f, chi2(loads after the leave) %.0f | "
              "join(w=8): moved %d minimal %d ratio %.2f, share of the joiner %.6f (expected %.6f) | %.1f s"
              % ("uniform" if uniform else "1..16", bits, chi, M - 1, (2 * (M - 1)) ** 0.5, (cnt / e).max(), (cnt / e).min(), moved, minimal, moved / max(minimal, 1), chi_d,
                 jm, jmin, jm / max(jmin, 1), jmin / N, 8 / (w.sum() + 8), dt))
        sys.stdout.flush()
