"""Per-kernel measurements for the configs of BASELINE.json beyond the bench.py headline (development/evidence tool;
run under gpurun, output committed under profiles/).  Every number is CUDA-event time on the engine stream after warm-up;
HBM fractions use MEASURED_PEAKS.json when present, else the 6650 GB/s fallback of B200_PROFILING.md."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rio_rs_b200 as R
from oracle import pyoracle as O

PEAK = 6650.0
src = "fallback"
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    try:
        PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]); src = "measured"
    except Exception:
        pass
big = "--small" not in sys.argv
p = R.GpuObjectPlacement(device=0)
out = []


def timed(fn, reps=5, warm=2, prov=None):
    prov = prov or p
    for _ in range(warm):
        fn()
    prov.sync()
    prov.event_record(0)
    for _ in range(reps):
        fn()
    prov.event_record(1)
    prov.sync()
    return prov.event_elapsed_ms(0, 1) / reps


def rec(name, ms, units, unit_name, bytes_=None, note=""):
    r = {"kernel": name, "ms": round(ms, 4), unit_name + "_per_s": units / (ms * 1e-3)}
    if bytes_ is not None:
        r.update({"algorithmic_bytes": bytes_, "GBps": bytes_ / (ms * 1e-3) / 1e9, "hbm_frac": bytes_ / (ms * 1e-3) / 1e9 / PEAK})
    if note:
        r["note"] = note
    out.append(r)
    print(json.dumps(r), flush=True)


# ---- C2: 1M x 64 weighted rendezvous --------------------------------------------------------------------
addrs, seeds, w = O.synth_nodes(64)
p.set_nodes(addrs, w)
s = p.new_set(1 << 20)
s.synth_keys(0, 1 << 20, 1)
rec("C2 assign 1M x 64 (k_assign_hrw_v2)", timed(s.assign, 20, 3), 1 << 20, "placements", 12 << 20)
del s

# ---- C4 shard / C5 storm kernels on a big resident set ----------------------------------------------------
M = 1024
addrs, seeds, w = O.synth_nodes(M + 2)
p2 = R.GpuObjectPlacement(device=0)
p2.set_nodes(addrs[:M], w[:M])
N = 100_000_000 if big else 10_000_000
s = p2.new_set(N)
s.synth_keys(0, N, 1)
t0 = time.perf_counter()
s.assign()
p2.sync()
rec("assign %dM x 1024 (k_assign_hrw_v2)" % (N // 1_000_000), (time.perf_counter() - t0) * 1e3, N, "placements", 12 * N, "single launch, wall clock")
# join: streaming compare, 12 B/object
j = p2.node_upsert(addrs[M], int(w[M]))
p2.sync()
p2.event_record(0)
moved = s.rebalance("join", j)
p2.event_record(1)
p2.sync()
rec("C5 join rebalance %dM objects (k_rebalance_join)" % (N // 1_000_000), p2.event_elapsed_ms(0, 1), N, "objects", 12 * N, "moved=%d, includes node-table rebuild + result readback" % moved)
# leave: 4 B/object scan + re-place N/M objects
p2.node_set_active(17, False)
p2.sync()
p2.event_record(0)
moved = s.rebalance("leave", 17)
p2.event_record(1)
p2.sync()
rec("C5 leave rebalance %dM objects (k_select_on_node + k_assign_hrw_v2 on movers)" % (N // 1_000_000), p2.event_elapsed_ms(0, 1), N, "objects", 4 * N, "moved=%d" % moved)
# the same set under HRW2: walk kernel (12 B/object), join / leave = re-walk + compare (16 B/object)
p2.set_solver("hrw2")
s.assign()
rec("assign %dM x 1024 HRW2 (k_assign_trie)" % (N // 1_000_000), timed(s.assign, 5, 1, p2), N, "placements", 12 * N)
rec("bounded pass %dM x 1024 HRW2 (k_assign_trie + fused exchange/check tail)" % (N // 1_000_000), timed(lambda: s.assign_bounded(0, 5, 4, 4), 5, 1, p2), N, "placements", 12 * N)
p2.node_set_active(33, False)
p2.sync()
p2.event_record(0)
moved = s.rebalance("leave", 33)
p2.event_record(1)
p2.sync()
rec("C5 leave rebalance %dM objects HRW2 (k_assign_trie, compare mode)" % (N // 1_000_000), p2.event_elapsed_ms(0, 1), N, "objects", 16 * N, "moved=%d, includes node-table rebuild" % moved)
p2.node_set_active(33, True)
p2.sync()
p2.event_record(0)
moved = s.rebalance("join", 33)
p2.event_record(1)
p2.sync()
rec("C5 join rebalance %dM objects HRW2 (k_assign_trie, compare mode)" % (N // 1_000_000), p2.event_elapsed_ms(0, 1), N, "objects", 16 * N, "moved=%d, includes node-table rebuild" % moved)
p2.set_solver("hrw")
del s

# ---- directory ---------------------------------------------------------------------------------------------
nd = 50_000_000 if big else 5_000_000
s = p2.new_set(nd)
s.synth_keys(0, nd, 2)
s.assign()
p2.sync()
t0 = time.perf_counter()
s.commit()
p2.sync()
placed, slots = p2.directory_len()
rec("directory upsert %dM keys (k_dir_upsert+finish, incl. growth)" % (nd // 1_000_000), (time.perf_counter() - t0) * 1e3, nd, "upserts", 36 * nd, "first commit: table grown/rehashed to %d slots" % slots)
rec("directory upsert %dM keys steady (k_dir_upsert+finish)" % (nd // 1_000_000), timed(s.commit, 3, 1, p2), nd, "upserts", 36 * nd)
keys_h, _ = s.read(0, min(nd, 20_000_000), want_keys=True)
L = p2.L
import ctypes as C
dk, do = C.c_void_p(), C.c_void_p()
nq = len(keys_h)
p2._ck(L.rio_cuda_dev_alloc(p2.h, nq * 8, C.byref(dk)))
p2._ck(L.rio_cuda_dev_alloc(p2.h, nq * 4, C.byref(do)))
p2._ck(L.rio_cuda_memcpy_h2d(p2.h, dk, keys_h.ctypes.data_as(C.c_void_p), nq * 8))
p2.sync()
rec("directory lookup %dM keys resident (k_dir_lookup)" % (nq // 1_000_000), timed(lambda: p2._ck(L.rio_cuda_lookup_batch_dev(p2.h, dk, nq, do)), 5, 2, p2), nq, "lookups", 28 * nq,
    "28 B algorithmic; ~44 B at 32 B-sector granularity")
ms = timed(lambda: p2.directory_len(), 3, 1, p2)
rec("directory scan %d slots (k_dir_count)" % slots, ms, slots, "slots", 16 * slots)
t0 = time.perf_counter()
removed = p2.clean_node(5)
rec("clean_server scan %d slots (k_dir_clean_node)" % slots, (time.perf_counter() - t0) * 1e3, slots, "slots", 16 * slots, "removed=%d, wall clock incl. readback" % removed)
j2 = p2.node_upsert(addrs[M + 1], int(w[M + 1]))
t0 = time.perf_counter()
mv = p2.rebalance("join", j2)
rec("directory-wide join rebalance %d slots (k_dir_rebalance_join)" % slots, (time.perf_counter() - t0) * 1e3, slots, "slots", 16 * slots, "moved=%d, wall clock" % mv)
p2.set_solver("hrw2")
p2.node_set_active(44, False)
t0 = time.perf_counter()
mv = p2.rebalance("leave", 44)
rec("directory-wide re-placement %d slots HRW2 (k_dir_reassign_trie)" % slots, (time.perf_counter() - t0) * 1e3, slots, "slots", 16 * slots, "moved=%d, wall clock" % mv)
p2.set_solver("hrw")
del s

# ---- device-side id hashing -----------------------------------------------------------------------------------
nh = 5_000_000
ids = [("Obj", str(i)) for i in range(nh)]
joined = [(t + "." + i).encode() for t, i in ids]
offs = np.zeros(nh + 1, dtype=np.uint64)
offs[1:] = np.cumsum([len(b) for b in joined])
packed = np.frombuffer(b"".join(joined) + b"\0" * 16, dtype=np.uint8)
outk = np.empty(nh, dtype=np.uint64)
t0 = time.perf_counter()
p2._ck(L.rio_cuda_hash_ids(p2.h, packed.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p), nh, outk.ctypes.data_as(C.c_void_p)))
ms = (time.perf_counter() - t0) * 1e3
assert outk[12345] == O.object_key("Obj", "12345")
rec("hash_ids 5M ids host->device->host (k_hash_ids)", ms, nh, "ids", int(offs[-1]) + 16 * nh, "wall clock incl. H2D of %d B and D2H of keys" % (int(offs[-1]) + 8 * nh))

# ---- C1: the reference's own config, CPU --------------------------------------------------------------------
sec, hits = O.bench_lookup(1000, 4, 2000)
out.append({"kernel": "C1 LocalObjectPlacement::lookup restatement, 1k ids x 4 nodes, CPU 1 thread", "ns_per_lookup": sec / (1000 * 2000) * 1e9, "lookups_per_s": 1000 * 2000 / sec})
print(json.dumps(out[-1]), flush=True)
print(json.dumps({"hbm_peak_gbs": PEAK, "peak_source": src, "device": p.device_info()["name"]}))
