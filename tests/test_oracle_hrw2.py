"""CPU tests of the HRW2 spec: the C oracle (oracle/rio_oracle.c, prefix sums over members sorted by
position) against the literal set recursion in tests/spec_py.py, the committed golden vectors, and the properties the
policy promises: P(node) = w/W, bounded movement on a membership change, independence of the order nodes are listed in."""
import importlib.util
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
NONE = 0xFFFFFFFF


def _spec():
    import spec_py

    return spec_py


def _mask(closed, M):
    m = np.zeros((M + 31) // 32, dtype=np.uint32)
    for j in closed:
        m[j >> 5] |= np.uint32(1 << (j & 31))
    return m


@pytest.mark.parametrize("M,bits", [(24, 12), (24, 3), (24, 0), (200, 5), (1, 12), (3, 1), (40, 14)])
def test_c_oracle_equals_python_restatement(oracle, M, bits):
    sp = _spec()
    addrs, seeds, w = oracle.synth_nodes(M)
    if M > 5:
        w[5] = 0
    keys = oracle.synth_keys(250, 3)
    pyseeds, pyw = [int(s) for s in seeds], [int(x) for x in w]
    got = oracle.assign_hrw2(keys, seeds, w, bits=bits)
    assert got.tolist() == [sp.hrw2(int(k), pyseeds, pyw, (), bits) for k in keys]
    closed = {1, 2, 7} if M > 8 else set()
    got = oracle.assign_hrw2(keys, seeds, w, mask=_mask(closed, M), bits=bits)
    assert got.tolist() == [sp.hrw2(int(k), pyseeds, pyw, closed, bits) for k in keys]
    ia, ca, pa = oracle.assign_bounded_hrw2(keys, seeds, w, 21, 20, 4, bits=bits)
    ib, cb, pb = sp.assign_bounded_hrw2([int(k) for k in keys], pyseeds, pyw, 21, 20, 4, bits)
    assert ia.tolist() == ib and ca.tolist() == cb and pa == pb


def test_scalars_match_python(oracle):
    sp = _spec()
    L = oracle.lib()
    for l in (0, 1, 5, 12, 63):
        assert L.orc_hrw2_level_seed(l) == sp.hrw2_level_seed(l)
    rng = np.random.default_rng(2)
    for k, s in rng.integers(0, 2**63, (200, 2)):
        assert L.orc_hrw2_v(int(k), int(s)) == sp.hrw2_v(int(k), int(s)) < 2**31
    for wl, wr in [(0, 0), (0, 9), (9, 0), (1, 1), (3, 5), (2**32 - 1, 1), (1, 2**40), (2**41, 2**41)]:
        want = ((wl << 31) // (wl + wr)) if wl + wr else 0
        assert L.orc_hrw2_threshold(wl, wr) == want
    assert L.orc_hrw2_threshold(7, 0) == 2**31 and L.orc_hrw2_threshold(0, 7) == 0   # forced contests


def test_no_live_node_is_none_and_single_node_takes_all(oracle):
    addrs, seeds, w = oracle.synth_nodes(8)
    keys = oracle.synth_keys(100, 1)
    assert (oracle.assign_hrw2(keys, seeds, np.zeros(8, dtype=np.uint32)) == NONE).all()
    w1 = np.zeros(8, dtype=np.uint32)
    w1[6] = 3
    assert (oracle.assign_hrw2(keys, seeds, w1) == 6).all()


def test_result_does_not_depend_on_listing_order(oracle):
    """Positions come from a hash of the address, chains are ordered by that hash: permuting the node list (another
    process interning the same live set in another order) must give the same ADDRESS for every object."""
    addrs, seeds, w = oracle.synth_nodes(300)
    keys = oracle.synth_keys(5000, 4)
    perm = np.random.default_rng(1).permutation(300)
    for bits in (12, 4):
        a = oracle.assign_hrw2(keys, seeds, w, bits=bits)
        b = oracle.assign_hrw2(keys, seeds[perm], w[perm], bits=bits)
        assert (perm[b] == a).all()


@pytest.mark.parametrize("uniform", [False, True])
def test_proportions_and_movement_bound(oracle, uniform):
    """chi-square of the node loads against w/W, and the price of the hierarchy: a leave/join moves at most about
    (1 + depth/2) x the minimal set, and nothing moves between two untouched siblings' subtrees."""
    M, n, bits = 256, 1_000_000, 12
    addrs, seeds, w = oracle.synth_nodes(M, uniform=uniform)
    keys = oracle.synth_keys(n, 1)
    a = oracle.assign_hrw2(keys, seeds, w, bits=bits, threads=8)
    cnt = np.bincount(a, minlength=M).astype(np.float64)
    e = n * w / w.sum()
    chi = ((cnt - e) ** 2 / e).sum()
    assert chi < (M - 1) + 5 * np.sqrt(2 * (M - 1)), chi
    w2 = w.copy()
    w2[17] = 0
    b = oracle.assign_hrw2(keys, seeds, w2, bits=bits, threads=8)
    moved, minimal = int((a != b).sum()), int((a == 17).sum())
    assert (b != 17).all() and (b[a == 17] != 17).all()
    depth = np.log2(M)
    assert minimal <= moved <= (1 + depth / 2 + 1.5) * minimal, (moved, minimal)
    # join back == the exact inverse
    assert (oracle.assign_hrw2(keys, seeds, w, bits=bits, threads=8) == a).all()


def test_golden_vectors_hrw2(oracle):
    g = json.load(open(os.path.join(GOLD, "solver_hrw2_v1.json")))
    L = oracle.lib()
    for l, v in g["level_seed"]:
        assert L.orc_hrw2_level_seed(l) == int(v)
    for k, s, v in g["contest_v"]:
        assert L.orc_hrw2_v(int(k), int(s)) == v
    for wl, wr, t in g["threshold"]:
        assert L.orc_hrw2_threshold(wl, wr) == t
    keys = np.array([int(k) for k in g["keys"]], dtype=np.uint64)
    seeds = np.array([int(s) for s in g["seeds"]], dtype=np.uint64)
    w = np.array(g["weights"], dtype=np.uint32)
    for bits, idx in g["idx"].items():
        assert oracle.assign_hrw2(keys, seeds, w, bits=int(bits)).tolist() == idx
    b = g["bounded"]
    idx, cnt, passes = oracle.assign_bounded_hrw2(keys, seeds, w, *b["cap"], b["max_rounds"], bits=b["bits"])
    assert idx.tolist() == b["idx"] and cnt.tolist() == b["counts"] and passes == b["passes"]


def test_golden_files_are_frozen():
    """A golden that moves with the code pins nothing: regenerating from the Python spec must reproduce the committed
    files byte for byte (flat rendezvous at pair-hash revision v3, HRW2 at revision 1)."""
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    for name, fn in mg.BUILDERS.items():
        committed = open(os.path.join(GOLD, mg.FILES[name])).read()
        assert mg.render(fn()) == committed, "tests/golden/%s differs from what tests/spec_py.py generates" % mg.FILES[name]
