"""Sample SM clock / power while the tcgen05 affinity kernel (and, for comparison, the rendezvous kernel) runs back to back."""
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rio_rs_b200 as R
from oracle import pyoracle as O

N, M = 10_000_000, 1024
p = R.GpuObjectPlacement(device=0)
addrs, _, w = O.synth_nodes(M)
p.set_nodes(addrs, w, np.random.default_rng(13).uniform(-1, 1, (M, 16)).astype(np.float32))
s = p.new_set(N)
s.synth_keys(0, N, 1)
s.load_feats(np.random.default_rng(5).uniform(-1, 1, (N, 16)).astype(np.float32))
Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown"
for name, aff in (("rendezvous", False), ("affinity-umma", True)):
    for _ in range(3):
        s.assign(aff)
    p.sync()
    proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=" + Q, "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
    t0 = time.perf_counter()
    p.event_record(0)
    reps = 0
    while time.perf_counter() - t0 < 2.5:
        for _ in range(20):
            s.assign(aff)
        p.sync()
        reps += 20
    p.event_record(1)
    p.sync()
    ms = p.event_elapsed_ms(0, 1) / reps
    time.sleep(0.2)
    proc.terminate()
    lines = [l.strip() for l in proc.stdout.read().splitlines() if l.strip()]
    print("%s: %.3f ms/launch over %d launches" % (name, ms, reps))
    for l in lines[2:-1][:8]:
        print("   ", l)
