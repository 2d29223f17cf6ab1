"""The host-only layers of librio_cuda on a box WITHOUT a GPU.

csrc/durable.cu (write-through into the reference's SQLite schema) and csrc/resolver.cu (micro-batching front
end for the per-id call sites) contain no device code: they are plain C++ written against the PUBLIC C ABI
(include/rio_cuda.h).  Here they are compiled as C++ (`g++ -x c++`, the very source files nvcc builds into the product) and linked
against an in-memory TEST DOUBLE of the dozen engine calls they use (tests/cpp/model_backend.cpp: LocalObjectPlacement's semantics,
local.rs:12-68, and the self-claim rule of service.rs:193-254), then driven by

  * tests/cpp/durable_conformance.cpp -- the harness the GPU box runs against the real engine (sqlite.rs:149-193,
    tests/object_placement_backend.rs:11-34, restart recovery, written-through place_batch), under AddressSanitizer + UBSan;
  * tests/cpp/durable_edge_cases.cpp  -- addresses longer than any fixed buffer, NULL ids, empty batches;
  * tests/cpp/resolver_stress.cpp     -- 16 concurrent per-id callers through the queue, also under ThreadSanitizer.

The double is test infrastructure (it lives under tests/, the product never sees it, and the real rio_cuda_create still fails
loudly without a GPU: tests/test_abi.py); the code UNDER test is the product's.
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "rio_rs_b200", "csrc")
TCPP = os.path.join(ROOT, "tests", "cpp")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")

pytestmark = pytest.mark.skipif(GXX is None, reason="no host C++ compiler")


def _build(tmp_path, name, product, harness, sanitize=None):
    exe = str(tmp_path / name)
    cmd = [GXX, "-std=c++17", "-O1", "-g", "-Wall", "-Wextra", "-Werror"]
    if sanitize:
        cmd += ["-fsanitize=" + sanitize, "-fno-sanitize-recover=all"]
    cmd += ["-x", "c++"] + [os.path.join(CSRC, f) for f in product] + [os.path.join(TCPP, "model_backend.cpp"), os.path.join(TCPP, harness), "-o", exe, "-ldl", "-lpthread"]
    try:
        subprocess.check_call(cmd)
    except subprocess.CalledProcessError:
        if not sanitize:
            raise
        pytest.skip("this toolchain has no runtime for -fsanitize=" + sanitize)
    return exe


def _sqlite_available():
    import ctypes

    for nm in ("libsqlite3.so.0", "libsqlite3.so"):
        try:
            ctypes.CDLL(nm)
            return True
        except OSError:
            pass
    return False


def _run(exe, *args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, timeout=300, env=e)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ThreadSanitizer" not in r.stderr and "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-4000:]
    return r.stdout


def test_product_layers_do_not_know_the_double():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rio_rs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                assert "model_backend" not in open(os.path.join(dirpath, f), errors="replace").read(), f
    for f in os.listdir(os.path.join(ROOT, "include")):
        assert "model_backend" not in open(os.path.join(ROOT, "include", f)).read()


@pytest.mark.parametrize("sanitize", [None, "address,undefined"])
def test_durable_write_through_conformance_on_the_double(tmp_path, sanitize):
    if not _sqlite_available():
        pytest.skip("libsqlite3 is not installed")
    exe = _build(tmp_path, "durable_cpu", ["durable.cu", "resolver.cu"], "durable_conformance.cpp", sanitize)
    assert "durable: all passed" in _run(exe, tmp_path)


def test_durable_edge_cases_on_the_double(tmp_path):
    if not _sqlite_available():
        pytest.skip("libsqlite3 is not installed")
    exe = _build(tmp_path, "durable_edge", ["durable.cu", "resolver.cu"], "durable_edge_cases.cpp", "address,undefined")
    assert "durable edge cases: all passed" in _run(exe, tmp_path)


@pytest.mark.parametrize("sanitize,threads,ids", [(None, 16, 300), ("thread", 8, 120)])
def test_resolver_queue_under_concurrent_callers(tmp_path, sanitize, threads, ids):
    exe = _build(tmp_path, "resolver_cpu", ["resolver.cu"], "resolver_stress.cpp", sanitize)
    out = _run(exe, threads, ids, env={"TSAN_OPTIONS": "halt_on_error=1"})
    assert "resolver: all passed" in out
