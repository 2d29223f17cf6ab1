#!/usr/bin/env python
"""bench.py -- placements/sec for the hot path named by BASELINE.json.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[3] -- weighted rendezvous placement of 10 M objects x 1024 nodes
(weights u32 in [1,16], seed 7; keys = splitmix stream), id-range sharded one shard per GPU, with the bounded-load
capacity check after ONE all-gather of the per-node load counters (cap 1.25, <= 4 rounds).  Weak scaling: every rank
owns a 10 M-object shard, so N GPUs place N x 10 M objects per step.

A step  = one bounded-load assignment pass over the rank's resident shard (keys already in HBM, results stay in HBM):
          score grid + argmin kernel with fused per-node histogram -> counter all-gather -> capacity check.
value   = objects placed by all ranks per second over K steps (CUDA events on the engine's stream, max over ranks).
e2e     = the same placements through the host-buffer C-ABI call rio_cuda_assign_batch (pinned host keys in, pinned host
          node indices out, H2D + D2H inside the timed region).
roofline= the assign kernel alone: algorithmic HBM bytes (12 B/object) / its launch time against the measured HBM peak;
          the kernel is integer-ALU bound by construction (1024 pair hashes per 12 bytes), so `alu_roofline` reports
          pair hashes/s against a register-only probe of the same instruction mix measured in the same run.
cpu_baseline = this repo's CPU port of the same solver spec (oracle/rio_oracle.c), all host cores, bounded sample.
--impl reference = the reference's own per-id CPU path (LocalObjectPlacement + Service::get_or_create_placement, restated
          in oracle/directory_model.cpp because rio-rs is Rust and no cargo exists here), all host threads.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_OBJECTS = 10_000_000
N_NODES = 1024
N_SETS = 4  # distinct resident key sets rotated step to step: 4 x 120 MB of traffic > 126 MB of L2
ALGO_BYTES_PER_OBJECT = 12  # 8 B key read + 4 B node index written
HBM_FALLBACK_GBS = 6650.0


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the assign kernel from the committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "r01_ncu_assign_v3.json")
    try:
        return float(json.load(open(p))["dram_bytes_per_launch"]), os.path.relpath(p, ROOT)
    except Exception:
        return None, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            if "hbm_gbs" in d:
                return float(d["hbm_gbs"]), "measured"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback"


def measured_tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            if "bf16_tflops" in d:
                return float(d["bf16_tflops"]), "measured"
        except Exception:
            pass
    return 1590.0, "fallback"


def extra_configs(p, O, n, seeds_unused):
    """The other single-GPU configs of BASELINE.json, measured with the same event discipline (N = 1 only):
    C2 = 1 M x 64 weighted rendezvous; C3 = 10 M x 1024, 16-dim affinity cost + argmin on the tensor cores."""
    import rio_rs_b200 as R

    out = {}
    q = R.GpuObjectPlacement(device=p.device_info()["device"])
    addrs, _, w = O.synth_nodes(64)
    q.set_nodes(addrs, w)
    s = q.new_set(1 << 20)
    s.synth_keys(0, 1 << 20, 1)
    for _ in range(5):
        s.assign()
    q.sync()
    q.event_record(0)
    for _ in range(50):
        s.assign()
    q.event_record(1)
    q.sync()
    ms = q.event_elapsed_ms(0, 1) / 50
    out["C2_rendezvous_1Mx64"] = {"ms": ms, "placements_per_s": (1 << 20) / (ms * 1e-3), "note": "resident keys, weights 1..16"}
    del s
    M, K = 1024, 16
    addrs, _, _ = O.synth_nodes(M)
    fn = np.random.default_rng(13).uniform(-1, 1, (M, K)).astype(np.float32)
    q.set_nodes(addrs, None, fn)
    s = q.new_set(n)
    s.load_keys(np.arange(n, dtype=np.uint64))
    fo = np.random.default_rng(11).uniform(-1, 1, (n, K)).astype(np.float32)
    s.load_feats(fo)
    for _ in range(3):
        s.assign(True)
    q.sync()
    q.event_record(0)
    for _ in range(10):
        s.assign(True)
    q.event_record(1)
    q.sync()
    ms = q.event_elapsed_ms(0, 1) / 10
    # parity spot-check of the tensor-core path against the fp64 oracle (checker only)
    got = s.read(0, 20000)
    idx, cost, gap = O.assign_affinity(fo[:20000], fn, np.ones(M, dtype=np.uint32), threads=8)
    ok = bool(((got == idx) | (gap <= 1e-5 * np.abs(cost) + 1e-12)).all())
    peak, src = measured_tensor_peak()
    issued = 6 * 2 * K * M * n / (ms * 1e-3) / 1e12  # six bf16 cross-term MMAs per (object, node) pair
    out["C3_affinity_10Mx1024xK16"] = {
        "ms": ms, "placements_per_s": n / (ms * 1e-3), "kernels": "k_affinity_umma (tcgen05/TMEM, bf16x3 split) + k_affinity_resolve",
        "parity_vs_fp64_oracle_20k": ok,
        "roofline": {"bound": "tensor", "achieved": issued, "peak": peak, "unit": "TFLOP/s", "frac": issued / peak, "peak_source": src,
                     "note": "bf16 FLOP/s actually issued (6 cross terms); algorithmic 2KM FLOP/s = achieved / 6"},
    }
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_indices):
        self.idx = ",".join(str(i) for i in gpu_indices)
        self.proc = None
        self.lines = []

    def start(self):
        if os.environ.get("RIO_BENCH_NO_SMI"):
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", self.idx, "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def mark(self):
        """Samples collected before this call (warm-up) are dropped."""
        self.lines = []

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def pinned_array(p, nbytes, dtype):
    ptr = C.c_void_p()
    p._ck(p.L.rio_cuda_host_alloc(p.h, nbytes, C.byref(ptr)))
    n = nbytes // np.dtype(dtype).itemsize
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(nbytes,)).view(dtype)[:n]
    return arr, ptr


def run_reference(args, rank, world):
    """The reference's own CPU path for this metric: per-id Service::get_or_create_placement over LocalObjectPlacement
    (service.rs:193-254, local.rs:12-68), all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import pyoracle as O

    cores = os.cpu_count() or 1
    per_step = 400_000
    O.bench_resolve(20_000, N_NODES, cores)  # warm caches / allocator
    for _ in range(max(args.warmup, 0)):
        O.bench_resolve(per_step // 4, N_NODES, cores)
    tot_s, tot_n = 0.0, 0
    for k in range(args.steps):
        s, placed = O.bench_resolve(per_step, N_NODES, cores, first=k * per_step)
        tot_s += s
        tot_n += placed
    v = tot_n / tot_s
    sample = "%d steps x %d fresh ids ('Obj', decimal i) resolved per-id against a %d-member cluster" % (args.steps, per_step, N_NODES)
    line = {
        "impl": "reference", "metric": "placements/sec", "value": v, "unit": "placements/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": "10M objects x 1024 nodes placement (BASELINE.json configs[3]); reference policy: first server to see the id claims it",
                   "objects_per_step": per_step, "nodes": N_NODES},
        "cpu_baseline": {"value": v, "unit": "placements/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "placements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "rio-rs is Rust; no cargo/rustc in this image, so this is the C++ restatement oracle/directory_model.cpp of local.rs + service.rs:193-254",
    }
    emit(line)


_JSON_FD = None


def claim_stdout():
    """stdout carries ONE JSON line.  Libraries print there too (NCCL's version banner when the box sets NCCL_DEBUG, nvcc
    notes, torch warnings), so file descriptor 1 is pointed at stderr for the rest of the run and the line is written to
    the original stdout at the end."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--objects", type=int, default=N_OBJECTS, help="objects per rank (default: the BASELINE size)")
    ap.add_argument("--nodes", type=int, default=N_NODES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C2/C3 side measurements (N = 1 only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import __graft_entry__ as G

    if rank == 0:
        G.build()
    if dist:
        dist.barrier()
    import rio_rs_b200 as R
    from rio_rs_b200 import parallel
    from oracle import pyoracle as O  # synthetic-input helpers + the cpu_baseline leg only

    n, M = args.objects, args.nodes
    p = R.GpuObjectPlacement(device=local_rank)
    info = p.device_info()
    addrs, seeds, w = O.synth_nodes(M)
    p.set_nodes(addrs, w)
    if dist:
        parallel.init_comm(p, dist)
    n_global = n * world

    def barrier_sync():
        p.sync()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    def max_over_ranks(x):
        if not dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # resident shards: id range [rank*n, (rank+1)*n) of N_SETS independent key streams
    sets = []
    for k in range(N_SETS):
        s = p.new_set(n)
        s.synth_keys(rank * n, n, 1 + k)
        sets.append(s)
    p.sync()

    def step(i):
        return sets[i % N_SETS].assign_bounded(n_global, 5, 4, 4)

    # ONE sampler for the whole job (rank 0 watches every GPU of the run): a poller per rank contends for the driver
    # and showed up as ~0.4 ms per step at N = 8 (profiles/r01_scale_n8.json vs r01_scale_n8_one_sampler.json)
    clocks = ClockSampler(range(world)) if rank == 0 else None
    if clocks:
        clocks.start()
    for i in range(args.warmup):
        passes = step(i)
    barrier_sync()
    if clocks:
        time.sleep(0.3)   # make sure the sampler is producing before the timed region starts
        clocks.mark()
    barrier_sync()
    l0 = p.launch_count()
    p.event_record(0)
    for i in range(args.steps):
        passes = step(i)
    p.event_record(1)
    barrier_sync()
    ms_total = max_over_ranks(p.event_elapsed_ms(0, 1))
    launches = p.launch_count() - l0
    ms_per_step = ms_total / args.steps
    value = n_global / (ms_per_step * 1e-3)

    # the dominant kernel alone (score grid + argmin + fused histogram), same resident inputs, events on its stream
    for i in range(3):
        sets[i % N_SETS].assign()
    p.sync()
    p.event_record(2)
    for i in range(args.steps):
        sets[i % N_SETS].assign()
    p.event_record(3)
    p.sync()
    if clocks and len(clocks.lines) < 3 * world:   # short runs: keep the same kernel running until a few samples exist
        t_end = time.perf_counter() + 0.6
        while time.perf_counter() < t_end:
            for i in range(10):
                sets[i % N_SETS].assign()
            p.sync()
    clk = clocks.stop() if clocks else None   # covers the timed steps, the kernel-only loop (and the burst above, if any)
    kern_ms = p.event_elapsed_ms(2, 3) / args.steps
    peak, peak_src = measured_peaks()
    traffic, traffic_src = ncu_traffic() if (n == N_OBJECTS and M == N_NODES) else (None, None)
    achieved_gbs = ALGO_BYTES_PER_OBJECT * n / (kern_ms * 1e-3) / 1e9
    pair_rate = n * M / (kern_ms * 1e-3)
    mix_peak = max(p.bench_mix_rate(4000) for _ in range(3))

    # e2e: host buffers through the C ABI (H2D + grid + D2H, chunk-pipelined)
    e2e = None
    if not args.no_e2e:
        hk, hk_ptr = pinned_array(p, n * 8, np.uint64)
        ho, ho_ptr = pinned_array(p, n * 4, np.uint32)
        hk[:] = O.synth_keys(n, 1, first=rank * n)
        for _ in range(2):
            p.assign_batch(hk, out=ho)
        barrier_sync()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            p.assign_batch(hk, out=ho)
        p.sync()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e_ms = 1e3 * dt / args.steps
        e2e = {"value": n_global / (e2e_ms * 1e-3), "unit": "placements/s", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 4 * n,
               "ms_per_step": e2e_ms, "api": "rio_cuda_assign_batch (pinned host keys -> pinned host node indices)"}
        # the result that came back over PCIe is the resident result
        assert (ho[:100000] == sets[0].read(0, 100000)).all()

    # parity spot-check inside the bench (the checker, never the thing measured)
    chk = sets[0].read(0, 20000)
    assert (chk == O.assign_hrw(O.synth_keys(20000, 1, first=rank * n), seeds, w, threads=4)).all(), "GPU result differs from the oracle"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        probe = O.synth_keys(20_000, 1)
        t0 = time.perf_counter()
        O.assign_hrw(probe, seeds, w, threads=cores)
        rate = len(probe) / (time.perf_counter() - t0)
        m = int(min(n, max(50_000, rate * 12)))  # ~12 s of CPU work
        ks = O.synth_keys(m, 1)
        t0 = time.perf_counter()
        O.assign_hrw(ks, seeds, w, threads=cores)
        dt = time.perf_counter() - t0
        cpu = {"value": m / dt, "unit": "placements/s", "cores": cores, "kind": "port",
               "sample": "first %d of the 10M objects x %d nodes, oracle/rio_oracle.c orc_assign_hrw, %d threads, %.1f s" % (m, M, cores, dt)}

    extra = None
    if rank == 0 and world == 1 and n == N_OBJECTS and not args.no_extra:
        try:
            extra = extra_configs(p, O, n, seeds)
        except Exception as e:  # the headline line must still be printed
            extra = {"error": repr(e)}

    if rank == 0:
        line = {
            "metric": "placements/sec", "value": value, "unit": "placements/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "10M objects x 1024 nodes weighted-rendezvous placement, id-range shard per GPU, bounded-load check after one all-gather of load counters (BASELINE.json configs[3])",
                       "objects_per_gpu": n, "global_objects": n_global, "nodes": M, "weights": "u32 in [1,16], seed 7", "capacity": "1.25", "max_rounds": 4,
                       "passes_run": passes, "l2": "inputs larger than L2: %d resident key sets rotated step to step" % N_SETS, "parallelism": "id-range shard x%d" % world,
                       "device": info["name"], "sms": info["sm_count"]},
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": peak, "unit": "GB/s", "frac": achieved_gbs / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "k_assign_hrw_v2", "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_OBJECT * n, "peak_source": peak_src,
                         "note": "integer-ALU bound by construction (1024 pair hashes per 12 B); see alu_roofline"},
            "alu_roofline": {"bound": "int-alu", "achieved": pair_rate, "peak": mix_peak, "unit": "pair-hashes/s", "frac": pair_rate / mix_peak,
                             "peak_source": "rio_cuda_bench_mix_rate: register-only replay of the same IMAD/IMAD/VIMNMX3 mix, measured in this run"},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clk,
            "extra_configs": extra,
        }
        emit(line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
