"""Client-side deterministic first hop (include/rio_client.h): bit-exact against the oracle's
weighted rendezvous hash, the reference client's behaviour around it (client/mod.rs:235-267), and the redirect rate
the reference's random pick would have had.  CPU only."""
import os
import re

import numpy as np
import pytest

from oracle import pyoracle as O
from rio_rs_b200 import client as CL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "rio_client.h")).read()
    declared = set(re.findall(r"\b(rio_client_[a-z_]+)\s*\(", hdr))
    assert declared == set(CL.SIGNATURES), declared ^ set(CL.SIGNATURES)
    L = CL.lib()
    for name in declared:
        assert hasattr(L, name)


@pytest.mark.parametrize("M,uniform", [(1, True), (4, True), (64, False), (300, False), (1024, False), (1024, True)])
def test_first_hop_equals_oracle_rendezvous(M, uniform):
    addrs, seeds, w = O.synth_nodes(M, uniform=uniform)
    fh = CL.FirstHop(addrs, w)
    keys = O.synth_keys(20000 if M <= 300 else 6000, 3)
    assert (fh.first_hop_batch(keys) == O.assign_hrw(keys, seeds, w, threads=4)).all()


def test_dead_nodes_raw_keys_and_no_servers():
    addrs, seeds, w = O.synth_nodes(16)
    w2 = w.copy()
    w2[::2] = 0                                               # weight 0 == not live
    fh = CL.FirstHop(addrs, w2)
    keys = np.concatenate([np.arange(0, 3000, dtype=np.uint64), np.array([2**64 - 1, 2**64 - 2, 0], dtype=np.uint64)])
    assert (fh.first_hop_batch(keys) == O.assign_hrw(keys, seeds, w2)).all()
    empty = CL.FirstHop([])
    assert (empty.first_hop_batch(keys[:10]) == CL.NONE).all()
    with pytest.raises(CL.NoServersAvailable):                # ClientError::NoServersAvailable (client/mod.rs:260-261)
        empty.get_service_object_address("Obj", "1")


def test_string_level_key_is_the_directory_key():
    assert CL.object_key("Obj", "7") == O.object_key("Obj", "7")
    assert CL.object_key("a.b", "c") == CL.object_key("a", "b.c")   # same aliasing as format!("{}.{}") (local.rs:26-29)
    addrs, seeds, w = O.synth_nodes(64)
    fh = CL.FirstHop(addrs, w)
    ids = [("Obj", str(i)) for i in range(500)]
    keys = np.array([O.object_key(t, i) for t, i in ids], dtype=np.uint64)
    want = O.assign_hrw(keys, seeds, w)
    assert [fh.get_service_object_address(t, i) for t, i in ids] == [addrs[j] for j in want]


def test_cache_hit_precedes_the_hash_and_is_bounded():
    addrs, _, w = O.synth_nodes(8)
    fh = CL.FirstHop(addrs, w, cache_size=3)
    owner = fh.get_service_object_address("Obj", "1")
    other = next(a for a in addrs if a != owner)
    fh.record_redirect("Obj", "1", other)                     # the server corrected us (tower_services.rs:158-168)
    assert fh.get_service_object_address("Obj", "1") == other
    for i in range(2, 6):
        fh.record_redirect("Obj", str(i), other)
    assert fh.get_service_object_address("Obj", "1") == owner  # evicted (LRU limit) -> back to the hash


def test_membership_change_moves_only_the_leavers_objects():
    addrs, seeds, w = O.synth_nodes(64)
    fh = CL.FirstHop(addrs, w)
    keys = O.synth_keys(20000, 9)
    before = fh.first_hop_batch(keys)
    w2 = w.copy()
    w2[17] = 0
    fh.set_active_servers(addrs, w2)                          # fetch_active_servers replaces the whole view
    after = fh.first_hop_batch(keys)
    assert ((before != after) == (before == 17)).all() and (after != 17).all()


def test_redirect_rate_against_the_reference_pick():
    """Objects placed by the servers with policy "hrw" (== the oracle's rendezvous assignment): the reference's uniform
    random first hop (client/mod.rs:254-263) is wrong (M-1)/M of the time, the rendezvous first hop never."""
    M = 64
    addrs, seeds, w = O.synth_nodes(M)
    keys = O.synth_keys(50000, 4)
    owner = O.assign_hrw(keys, seeds, w, threads=4)
    fh = CL.FirstHop(addrs, w)
    assert int((fh.first_hop_batch(keys) != owner).sum()) == 0
    rnd = np.random.default_rng(0).integers(0, M, len(keys))
    rate = float((rnd != owner).mean())
    assert abs(rate - (M - 1) / M) < 0.01


def test_server_product_does_not_use_the_client_library():
    for f in ("provider.py", "_native.py", "parallel.py", "durable.py", "__init__.py"):
        src = open(os.path.join(ROOT, "rio_rs_b200", f)).read()
        assert "librio_client" not in src and "rio_client" not in src and "from .client" not in src and "import client" not in src, f
    eng = "".join(open(os.path.join(ROOT, "rio_rs_b200", "csrc", f)).read() for f in ("engine.cu", "resolver.cu"))
    assert "rio_client" not in eng


def test_cpp_mirror_conformance(tmp_path):
    """rio_rs_b200/host/first_hop.hpp (the C++ mirror of the client piece) against oracle-made expectations."""
    import subprocess

    CL.lib()
    addrs, seeds, w = O.synth_nodes(40)
    ids = [("Obj", str(i)) for i in range(400)]
    keys = np.array([O.object_key(t, i) for t, i in ids], dtype=np.uint64)
    want = O.assign_hrw(keys, seeds, w)
    (tmp_path / "nodes.txt").write_text("".join("%s %d\n" % (a, x) for a, x in zip(addrs, w)))
    (tmp_path / "ids.txt").write_text("".join("%s %s %s\n" % (t, i, addrs[j]) for (t, i), j in zip(ids, want)))
    exe = str(tmp_path / "first_hop_conformance")
    libdir = os.path.join(ROOT, "rio_rs_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cpp", "first_hop_conformance.cpp"),
                           "-L" + libdir, "-lrio_client", "-Wl,-rpath," + libdir])
    r = subprocess.run([exe, str(tmp_path / "nodes.txt"), str(tmp_path / "ids.txt")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "all passed" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("bits", [0, 12, 4])
def test_first_hop_hrw2_equals_the_oracle_and_ignores_listing_order(oracle, bits):
    """The client under the hierarchical policy picks the oracle's node for every id, also after a node left,
    and -- positions and chains being ordered by a hash of the address -- independently of the order it lists the servers in."""
    from rio_rs_b200 import client as CL

    addrs, seeds, w = oracle.synth_nodes(200)
    w[5] = 0
    keys = oracle.synth_keys(20000, 4)
    fh = CL.FirstHop(addrs, w, policy="hrw2", trie_bits=bits)
    want = oracle.assign_hrw2(keys, seeds, w, bits=bits or 12)
    assert (fh.first_hop_batch(keys) == want).all()
    perm = np.random.default_rng(3).permutation(200)
    fh2 = CL.FirstHop([addrs[j] for j in perm], w[perm], policy="hrw2", trie_bits=bits)
    assert (perm[fh2.first_hop_batch(keys)] == want).all()
    w2 = w.copy()
    w2[9] = 0
    fh.set_active_servers(addrs, w2)
    assert (fh.first_hop_batch(keys) == oracle.assign_hrw2(keys, seeds, w2, bits=bits or 12)).all()
    assert CL.FirstHop([], None, policy="hrw2").first_hop_batch(keys[:5]).tolist() == [CL.NONE] * 5


@pytest.mark.parametrize("M,bits,weights", [(1, 12, "ones"), (2, 1, "ones"), (300, 2, "mixed"), (1024, 12, "mixed"), (1024, 10, "ones"), (3000, 14, "mixed"),
                                            (64, 6, "huge"), (500, 3, "huge")])
def test_shared_table_builder_walked_on_the_host_equals_the_oracle(oracle, M, bits, weights):
    """csrc/trie_table.hpp is the ONE builder of the HRW2 table: engine.cu uploads its blob for the kernels, the client library
    walks the same blob on the host.  Long chains (few bits, many nodes), single-member buckets, weights near 2^32 (the 128-bit
    branch of the threshold division), dead nodes and the empty live set -- all against the oracle, which never sees a blob."""
    from rio_rs_b200 import client as CL

    addrs, seeds, w = oracle.synth_nodes(M, uniform=(weights == "ones"))
    if weights == "huge":
        w = w.astype(np.uint64) * np.uint64(0x0FFFFFFF) + np.uint64(7)
        w = np.minimum(w, np.uint64(0xFFFFFFFF)).astype(np.uint32)
    if M > 2:
        w[::5] = 0
    keys = oracle.synth_keys(20000, 11)
    fh = CL.FirstHop(addrs, w, policy="hrw2", trie_bits=bits)
    got = fh.first_hop_batch(keys)
    assert (got == oracle.assign_hrw2(keys, seeds, w, bits=bits)).all()
    assert M <= 2 or not np.isin(got, np.nonzero(w == 0)[0]).any()
    fh.set_active_servers(addrs, np.zeros(M, dtype=np.uint32))          # nobody live: every walk ends on an empty bucket
    assert (fh.first_hop_batch(keys[:100]) == CL.NONE).all()


def test_table_builder_shortcuts_equal_the_plain_statement(tmp_path):
    """tests/cpp/trie_table_selftest.cpp: the shipped builder (no full sort, no 64-bit divisions: a rebuild is on the critical path of
    every membership event) against the layout of DESIGN.md 3.8 / 4.1 read literally -- byte for byte on 324 member sets, and the fast
    threshold division against the reference one on 4 M operand pairs (edges, random, next to exact multiples)."""
    import shutil
    import subprocess

    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not gxx:
        pytest.skip("no host C++ compiler")
    exe = str(tmp_path / "selftest")
    subprocess.check_call([gxx, "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "tests", "cpp", "trie_table_selftest.cpp"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "all passed" in r.stdout, r.stdout + r.stderr
