"""Small pass over every kernel of librio_cuda for compute-sanitizer (memcheck / racecheck); sizes kept tiny.
Usage (GPU box): compute-sanitizer --tool memcheck python tools/sanitize_smoke.py [--no-umma]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rio_rs_b200 as R
from oracle import pyoracle as O

skip_umma = "--no-umma" in sys.argv
p = R.GpuObjectPlacement(device=0, directory_capacity=1024)
addrs, seeds, w = O.synth_nodes(70)
p.set_nodes(addrs[:64], w[:64])
keys = O.synth_keys(5000, 1)
want = O.assign_hrw(keys, seeds[:64], w[:64])
assert (p.assign_batch(keys) == want).all()                         # k_assign_hrw_v2 (+ chunked host pipeline)
os.environ["RIO_ASSIGN_VARIANT"] = "1"
assert (p.assign_batch(keys) == want).all()                         # k_assign_hrw
os.environ.pop("RIO_ASSIGN_VARIANT")
big = R.GpuObjectPlacement(device=0)                                # full 32-node groups + 16/8/4/2/1 tails + rotated index scan
for M, uniform in ((203, True), (333, False)):
    a3, s3, w3 = O.synth_nodes(M, uniform=uniform)
    big.set_nodes(a3, w3)
    assert (big.assign_batch(keys[:3000]) == O.assign_hrw(keys[:3000], s3, w3)).all()
del big
assert (p.place_batch(keys, "hrw") == want).all()                   # lookup, classify, assign(sel), gather, upsert (with growth + rehash)
assert (p.lookup_many(keys) == want).all()
ids = [("Obj", str(i)) for i in range(3000)]
hk = p.hash_ids(ids)                                                # k_hash_ids
assert hk[7] == O.object_key("Obj", "7")
p.update_many(hk, np.full(len(hk), 3, dtype=np.uint32))
assert p.clean_node(3) >= len(hk)                                   # k_dir_clean_node
p.remove_many(hk[:100])
p.node_set_active(5, False)
p.place_batch(keys[:2000], "self", addrs[1])                        # clean_flagged + scatter_const path
moved = p.rebalance("leave", 5)                                     # k_dir_rebalance_leave
j = p.node_upsert(addrs[64], int(w[64]))
p.rebalance("join", j)                                              # k_dir_rebalance_join
p.load_counters(); p.directory_len()
s = p.new_set(20000)
s.synth_keys(0, 20000, 2)                                           # k_synth_keys
s.assign()
s.assign_bounded(0, 101, 100, 4)                                    # k_select_spill + masked table
p.node_set_active(9, False)
s.rebalance("leave", 9)                                             # k_select_on_node + assign(sel)
j2 = p.node_upsert(addrs[65], int(w[65]))
s.rebalance("join", j2)                                             # k_rebalance_join
s.counters(); s.commit()
r = R.Resolver(p, policy="hrw", max_wait_us=100)
for k in keys[:50]:
    r.resolve(int(k))
r.close()
p.bench_mix_rate(2)                                                 # k_mix_rate
# ---- HRW2: the trie walk (TMA-staged table, dense / gathered / compare / directory forms), fused exchange + check tail --------
h2 = R.GpuObjectPlacement(device=0, directory_capacity=1024)
a5, s5, w5 = O.synth_nodes(130)
h2.set_nodes(a5[:128], w5[:128])
for bits in (12, 3, 14):
    h2.set_solver("hrw2", bits)
    assert (h2.assign_batch(keys) == O.assign_hrw2(keys, s5[:128], w5[:128], bits=bits)).all()          # k_assign_trie<12|0, smem>, chains at bits 3
h2.set_solver("hrw2", 12)
assert (h2.place_batch(keys, "hrw2") == O.assign_hrw2(keys, s5[:128], w5[:128])).all()                   # k_assign_trie_sel
t5 = h2.new_set(20001)                                                                                   # odd size: the 128-bit load tail
t5.synth_keys(0, 20001, 3)
k5 = O.synth_keys(20001, 3)
assert t5.assign_bounded(0, 5, 4, 4) >= 1                                                                # fused tail (last-CTA ticket, check, mapped flags)
wi, wc, wp = O.assign_bounded_hrw2(k5, s5[:128], w5[:128], 101, 100, 4)
assert t5.assign_bounded(0, 101, 100, 4) == wp and (t5.read() == wi).all()                               # spill rounds, masked trie, k_exchange_check
t5.assign_bounded_begin(0, 5, 4, 4)                                                                      # check on the auxiliary stream
t5.assign_bounded_end()
t5.assign()
t5.commit()
h2.node_set_active(9, False)
t5.rebalance("leave", 9)                                                                                 # k_assign_trie compare mode
h2.rebalance("leave", 9)                                                                                 # k_dir_reassign_trie
w5b = w5[:128].copy()
w5b[9] = 0
assert (t5.read() == O.assign_hrw2(k5, s5[:128], w5b)).all() and (h2.lookup_many(k5) == t5.read()).all()
v, cleaned = h2.check_address_batch(h2.lookup_many(k5[:3000]), a5[1])                                     # k_check_address (+ clean_flagged when a dead owner is met)
hb = R.GpuObjectPlacement(device=0)                                                                      # table too large for shared memory: global walk
a6, s6, w6 = O.synth_nodes(20000)
hb.set_nodes(a6, w6)
hb.set_solver("hrw2", 14)
assert (hb.assign_batch(keys[:2000]) == O.assign_hrw2(keys[:2000], s6, w6, bits=14)).all()
del hb
# affinity: CUDA-core kernels always, tensor-core kernel unless --no-umma
q = R.GpuObjectPlacement(device=0)
fn = np.random.default_rng(1).uniform(-1, 1, (300, 16)).astype(np.float32)
fo = np.random.default_rng(2).uniform(-1, 1, (3000, 16)).astype(np.float32)
a2, _, _ = O.synth_nodes(300)
q.set_nodes(a2, None, fn)
idx, cost, gap = O.assign_affinity(fo, fn, np.ones(300, dtype=np.uint32))
for var in (["ffma"] if skip_umma else ["ffma", "umma"]):
    os.environ["RIO_AFFINITY_VARIANT"] = var
    got = q.assign_batch(obj_feats=fo)
    assert ((got == idx) | (gap <= 1e-5 * np.abs(cost))).all(), var
os.environ.pop("RIO_AFFINITY_VARIANT")
q8 = R.GpuObjectPlacement(device=0)
q8.set_nodes(a2[:20], None, fn[:20, :8].copy())
q8.assign_batch(obj_feats=fo[:500, :8].copy())                       # generic-K kernel
print("sanitize_smoke: all paths ran, results match the oracle")
