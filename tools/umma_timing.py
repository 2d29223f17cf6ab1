"""Per-role cycle breakdown of k_affinity_umma (development tool): where do producers / the MMA issuer / the epilogue wait?"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rio_rs_b200 as R
from oracle import pyoracle as O

N, M = 10_000_000, 1024

p = R.GpuObjectPlacement(device=0)
addrs, _, _ = O.synth_nodes(M)
p.set_nodes(addrs, None, np.random.default_rng(13).uniform(-1, 1, (M, 16)).astype(np.float32))
s = p.new_set(N)
s.synth_keys(0, N, 1)
s.load_feats(np.random.default_rng(5).uniform(-1, 1, (N, 16)).astype(np.float32))
s.assign(True)
p.sync()
L = C.CDLL(R.library_path())
buf = C.c_void_p()
p._ck(p.L.rio_cuda_dev_alloc(p.h, 148 * 16 * 8, C.byref(buf)))
L.rio_dev_umma_timing.argtypes = [C.c_void_p, C.c_void_p]
L.rio_dev_umma_timing(p.h, buf)
p.event_record(0)
s.assign(True)
p.event_record(1)
p.sync()
ms = p.event_elapsed_ms(0, 1)
out = np.zeros(148 * 16, dtype=np.uint64)
p._ck(p.L.rio_cuda_memcpy_d2h(p.h, out.ctypes.data_as(C.c_void_p), buf, out.nbytes))
p.sync()
L.rio_dev_umma_timing(p.h, None)
t = out.reshape(148, 16).astype(np.float64)
names = ["producer wait a_empty", "producer work", "mma wait a_full", "mma wait t_empty", "mma issue+commit", "epi wait t_full", "epi ld+reduce", "epi resolve", "setup (B operands, TMEM)", "CTA total"]
rb_per_cta = (N / 128) / 148
print("kernel %.3f ms (with timing hooks); per CTA: %.0f row blocks, %.0f tiles; kernel cycles ~%.0f" % (ms, rb_per_cta, rb_per_cta * 4, ms * 1e-3 * 1.965e9))
for k, nm in enumerate(names):
    per = t[:, k].mean()
    print("  %-24s %12.0f cycles/CTA (min %10.0f max %10.0f)  %8.1f per row block  %7.1f per tile" % (nm, per, t[:, k].min(), t[:, k].max(), per / rb_per_cta, per / (rb_per_cta * 4)))
starts = t[:, 10] - t[:, 10].min()
print("  CTA start skew: max %.0f cycles; clock64 rate check: CTA total mean %.0f cycles vs %.3f ms -> %.3f GHz" % (starts.max(), t[:, 9].mean(), ms, t[:, 9].mean() / (ms * 1e-3) / 1e9))
