"""One affinity pass at full size for ncu captures of k_affinity_umma (development tool)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rio_rs_b200 as R
from oracle import pyoracle as O

N, M = 10_000_000, 1024
p = R.GpuObjectPlacement(device=0)
addrs, _, _ = O.synth_nodes(M)
fn = np.random.default_rng(13).uniform(-1, 1, (M, 16)).astype(np.float32)
p.set_nodes(addrs, None, fn)
s = p.new_set(N)
s.synth_keys(0, N, 1)
s.load_feats(np.random.default_rng(5).uniform(-1, 1, (N, 16)).astype(np.float32))
for _ in range(3):
    s.assign(True)
p.sync()
print("done")
