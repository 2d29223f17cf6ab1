"""Dry run of bench.py's control flow on a box WITHOUT a GPU: the Python mirror is pointed at the host-sim library
(tests/_build/librio_cuda_hostsim.so, see tests/test_engine_host_sim.py), torch.cuda is stubbed, sizes are shrunk.  Every number it
prints is meaningless; the point is that every branch of bench.py (steps in flight, both policies, e2e leg, C4-strong / C5 / C2 / C3 / C1
side measurements, the JSON line) executes against the current host code before the driver runs it on a real B200.
    python -m pytest tests/test_engine_host_sim.py -q -k gpu_test_bodies      # builds the library
    python tools/bench_dryrun.py"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
so = os.path.join(ROOT, "tests", "_build", "librio_cuda_hostsim.so")
assert os.path.exists(so), "build the host-sim library first (tests/test_engine_host_sim.py)"

from rio_rs_b200 import _native  # noqa: E402

_native.library_path = lambda: so
_native._lib = None

import torch  # noqa: E402

torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.get_device_properties = lambda d: types.SimpleNamespace(pci_domain_id=0, pci_bus_id=0, pci_device_id=0)

import bench  # noqa: E402
from rio_rs_b200 import parallel  # noqa: E402

n = 205_000
bench.N_OBJECTS = n
bench.N_NODES = 1024
_shard = parallel.shard_range
parallel.shard_range = lambda total, rank, world: _shard(min(total, 200_000), rank, world)   # C5's 100 M objects -> 200 k
os.environ["RIO_BENCH_NO_SMI"] = "1"
sys.argv = ["bench.py", "--steps", "6", "--warmup", "3", "--objects", str(n), "--nodes", "1024", "--no-cpu-baseline"] + sys.argv[1:]
r, w = os.pipe()
bench._JSON_FD = w
bench.claim_stdout = lambda: None
bench.main()
os.close(w)
line = json.loads(os.read(r, 1 << 20).decode())
assert line["config"]["parity_vs_oracle_200k_per_rank"] and line["gpu_launches"] > 0 and line["e2e"]["value"] > 0, line
extra = line["extra_configs"]
assert extra and "error" not in extra, extra
print("bench dry run ok: keys", sorted(line), "extras", sorted(extra))
