"""(needs a library built with RIO_BUILD_TUNING=1) A/B timing of the compiled tuning points of the HRW2 walk kernel
(objects per thread x CTAs per SM), parity checked per run (development tool; run under gpurun)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rio_rs_b200 as R
from oracle import pyoracle as O

n, M = 10_000_000, 1024
p = R.GpuObjectPlacement(device=0)
addrs, seeds, w = O.synth_nodes(M)
p.set_nodes(addrs, w)
p.set_solver("hrw2")
sets = []
for k in range(4):
    s = p.new_set(n)
    s.synth_keys(0, n, 1 + k)
    sets.append(s)
want = O.assign_hrw2(O.synth_keys(100_000, 1), seeds, w, threads=8)
for code in ("", "45", "44", "25", "28", "63", "64", "83"):
    if code:
        os.environ["RIO_TRIE_TUNE"] = code
    else:
        os.environ.pop("RIO_TRIE_TUNE", None)
    for i in range(4):
        sets[i % 4].assign()
    p.sync()
    p.event_record(0)
    for i in range(200):
        sets[i % 4].assign()
    p.event_record(1)
    p.sync()
    ok = bool((sets[0].read(0, 100_000) == want).all())
    print("tune %-8s %.2f us  parity %s" % (code or "default", p.event_elapsed_ms(0, 1) / 200 * 1e3, ok), flush=True)
