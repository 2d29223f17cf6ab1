"""C5 -- rebalance storm: 100 M objects id-sharded over the ranks, 8 sequential join/leave events (SURVEY 8d list), one
all-gather of the load counters after every event.  Launch with torchrun (one rank per GPU) under gpurun --gpus N.
Prints one JSON line from rank 0; every rank checks a sample of its shard against the oracle at the end."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import rio_rs_b200 as R
from oracle import pyoracle as O
from rio_rs_b200 import parallel

N_TOTAL = int(os.environ.get("STORM_OBJECTS", 100_000_000))
M0 = 1024
EVENTS = [("leave", 17), ("join", 1024), ("leave", 3), ("join", 1025), ("leave", 900), ("join", 1026), ("leave", 64), ("join", 1027)]
p = R.GpuObjectPlacement(device=local)
if world > 1:
    parallel.init_comm(p, dist)
addrs, seeds, w = O.synth_nodes(M0 + 4)
p.set_nodes(addrs[:M0], w[:M0])
lo, hi = parallel.shard_range(N_TOTAL, rank, world)
n = hi - lo
s = p.new_set(n)
s.synth_keys(lo, n, 1)
s.assign()
c0 = s.counters()
assert int(c0.sum()) == N_TOTAL
w_live = w.copy()
w_live[M0:] = 0


def sync_all():
    p.sync()
    if world > 1:
        dist.barrier()


sync_all()
per_event = []
t_all = time.perf_counter()
for ev, j in EVENTS:
    t0 = time.perf_counter()
    if ev == "leave":
        p.node_set_active(j, False)
        w_live[j] = 0
    else:
        assert p.node_upsert(addrs[j], int(w[j])) == j
        w_live[j] = w[j]
    moved = s.rebalance(ev, j)
    cnt = s.counters()          # the one collective of the event: all-gather + sum of the per-node counters
    sync_all()
    dt = time.perf_counter() - t0
    assert int(cnt.sum()) == N_TOTAL and (ev == "join" or cnt[j] == 0)
    per_event.append({"event": ev, "node": j, "moved_local": int(moved), "ms": dt * 1e3})
total_s = time.perf_counter() - t_all
# parity: a sample of this rank's shard against a from-scratch oracle assignment over the final live set
pick = np.sort(np.random.default_rng(rank).choice(n, 20000, replace=False))
keys = O.synth_keys(n, 1, first=lo) if n <= 20_000_000 else None
if keys is None:
    k_all, idx_all = s.read(0, n, want_keys=True)
    keys = k_all
    idx = idx_all
else:
    idx = s.read()
want = O.assign_hrw(keys[pick], seeds, w_live, threads=8)
assert (idx[pick] == want).all(), "rank %d: storm result differs from the oracle" % rank
if world > 1:
    t = torch.tensor([total_s], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_s = float(t.item())
if rank == 0:
    print(json.dumps({"bench": "C5 rebalance storm", "objects": N_TOTAL, "n_gpus": world, "events": len(EVENTS), "total_ms": total_s * 1e3,
                      "object_events_per_s": N_TOTAL * len(EVENTS) / total_s, "per_event": per_event, "parity_sample": "20000 objects/rank vs oracle: ok",
                      "timing": "wall clock per event incl. node-table rebuild, kernels, counter all-gather, readback, barrier"}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
