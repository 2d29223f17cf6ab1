"""GPU parity tests of the HRW2 policy (hierarchical weighted rendezvous, fan-out 2): every call goes
through the C ABI and is compared, index for index, with the CPU oracle (oracle/rio_oracle.c: orc_assign_hrw2), which
re-derives every contest from prefix sums while the engine precomputes a heap of thresholds -- two independent routes to
the same integers.  There are no ties in this policy (every contest is a strict compare), so there is no tie path to miss."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NONE = 0xFFFFFFFF
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gp():
    from rio_rs_b200 import build

    build.build()
    import rio_rs_b200 as R

    return R


def provider(gp, bits=0, **kw):
    p = gp.GpuObjectPlacement(**kw)
    p.set_solver("hrw2", bits)
    return p


@pytest.mark.parametrize("M,n,bits,uniform", [(64, 1_000_000, 12, False), (1024, 200_000, 12, False), (1024, 200_000, 12, True), (1, 1000, 12, False),
                                              (3, 5001, 12, False), (200, 50_001, 5, False), (64, 20_000, 1, False), (500, 30_000, 14, False),
                                              (5000, 40_000, 12, False)])
def test_assign_matches_oracle(gp, oracle, M, n, bits, uniform):
    """C2 (1M x 64) and C4-shaped cases; bits 5 and 1 put many nodes in one bucket (long member-keyed chains), bits 14 is
    the deepest trie, M = 5000 at bits 12 has a chain in most buckets; ragged n exercises the 128-bit load tail."""
    p = provider(gp, bits)
    addrs, seeds, w = oracle.synth_nodes(M, uniform=uniform)
    if M > 10:
        w[5] = 0
    p.set_nodes(addrs, w)
    assert p.get_solver() == ("hrw2", bits or 12)
    keys = oracle.synth_keys(n, 1 + (M % 3))
    want = oracle.assign_hrw2(keys, seeds, w, bits=bits or 12, threads=8)
    assert (p.assign_batch(keys) == want).all()                         # host-buffer API (pipelined H2D / D2H)
    s = p.new_set(n)
    s.load_keys(keys)
    s.assign()                                                          # resident set + fused histogram
    assert (s.read() == want).all()
    assert (s.counters() == oracle.counts(want, M)).all()


def test_table_too_large_for_shared_memory_falls_back_to_global(gp, oracle):
    """70 000 live nodes: ~17 members per bucket at 12 bits, 1.4 MB of chain records -- the kernel walks the table through
    the read-only path instead of the TMA-staged copy."""
    M = 70_000
    p = provider(gp)
    addrs, seeds, w = oracle.synth_nodes(M)
    p.set_nodes(addrs, w)
    keys = oracle.synth_keys(40_000, 2)
    want = oracle.assign_hrw2(keys, seeds, w, threads=8)
    assert (p.assign_batch(keys) == want).all()
    s = p.new_set(len(keys))
    s.load_keys(keys)
    s.assign()
    assert (s.read() == want).all() and (s.counters() == oracle.counts(want, M)).all()


def test_edge_cases_empty_dead_and_raw_keys(gp, oracle):
    p = provider(gp)
    addrs, seeds, w = oracle.synth_nodes(16)
    assert (p.assign_batch(oracle.synth_keys(100, 1)) == NONE).all()      # no nodes at all
    w2 = w.copy()
    w2[::2] = 0
    p.set_nodes(addrs, w2)
    p.node_set_active(1, False)
    w2[1] = 0
    keys = np.concatenate([np.arange(0, 5000, dtype=np.uint64), np.array([2**64 - 1, 2**64 - 2, 0], dtype=np.uint64)])   # raw, unmixed keys
    assert (p.assign_batch(keys) == oracle.assign_hrw2(keys, seeds, w2)).all()
    assert p.assign_batch(np.empty(0, np.uint64)).shape == (0,)
    p.node_set_active(1, True)
    w2[1] = w[1]
    assert (p.assign_batch(keys) == oracle.assign_hrw2(keys, seeds, w2)).all()
    big = np.random.default_rng(4).integers(1, 2**31, 16).astype(np.uint32)   # u32 weights: the subtree sums need 64 bits
    big[3] = 0xFFFFFFFF
    p.set_nodes(addrs, big)
    assert (p.assign_batch(keys) == oracle.assign_hrw2(keys, seeds, big)).all()


def test_golden_vectors_on_gpu(gp):
    g = json.load(open(os.path.join(GOLD, "solver_hrw2_v1.json")))
    keys = np.array([int(k) for k in g["keys"]], dtype=np.uint64)
    w = np.array(g["weights"], dtype=np.uint32)
    for bits, idx in g["idx"].items():
        p = provider(gp, int(bits))
        p.set_nodes(g["addresses"], w)
        assert p.assign_batch(keys).tolist() == idx, bits
    b = g["bounded"]
    p = provider(gp, b["bits"])
    p.set_nodes(g["addresses"], w)
    s = p.new_set(len(keys))
    s.load_keys(keys)
    passes = s.assign_bounded(0, *b["cap"], b["max_rounds"])
    assert passes == b["passes"] and s.read().tolist() == b["idx"] and s.counters().tolist() == b["counts"]


def test_bounded_rounds(gp, oracle):
    p = provider(gp)
    addrs, seeds, w = oracle.synth_nodes(64)
    p.set_nodes(addrs, w)
    n = 200_000
    s = p.new_set(n)
    s.synth_keys(0, n, 1)
    keys = oracle.synth_keys(n, 1)
    for cap in [(5, 4), (101, 100), (1, 1)]:
        passes = s.assign_bounded(0, cap[0], cap[1], 4)
        widx, wcnt, wpass = oracle.assign_bounded_hrw2(keys, seeds, w, cap[0], cap[1], 4, threads=8)
        assert passes == wpass, cap
        assert (s.read() == widx).all(), cap
        assert (s.counters() == wcnt).all(), cap


def test_rebalance_storm_matches_fresh_assignment(gp, oracle):
    """C5 at test scale under HRW2: after each of 8 join/leave events the resident set and the directory equal a
    from-scratch assignment over the new live set; `moved` is exactly the number of objects whose node changed, a leaving
    node ends empty, and the movement stays within the hierarchy's bound."""
    M0 = 128
    p = provider(gp)
    addrs, seeds, w = oracle.synth_nodes(M0 + 4)
    w_live = w.copy()
    w_live[M0:] = 0
    p.set_nodes(addrs[:M0], w[:M0])
    n = 300_000
    keys = oracle.synth_keys(n, 2)
    s = p.new_set(n)
    s.load_keys(keys)
    s.assign()
    s.commit()
    events = [("leave", 17), ("join", M0), ("leave", 3), ("join", M0 + 1), ("leave", 100), ("join", M0 + 2), ("leave", 64), ("join", M0 + 3)]
    prev = s.read().copy()
    assert (prev == oracle.assign_hrw2(keys, seeds, w_live, threads=8)).all()
    for ev, j in events:
        if ev == "leave":
            p.node_set_active(j, False)
            w_live[j] = 0
        else:
            assert p.node_upsert(addrs[j], int(w[j])) == j
            w_live[j] = w[j]
        moved = s.rebalance(ev, j)
        dmoved = p.rebalance(ev, j)
        want = oracle.assign_hrw2(keys, seeds, w_live, threads=8)
        got = s.read()
        assert (got == want).all(), (ev, j)
        assert moved == int((prev != want).sum()) == dmoved, (ev, j)
        minimal = int((prev == j).sum()) if ev == "leave" else int((want == j).sum())
        assert minimal <= moved <= (2.5 + np.log2(M0) / 2) * minimal, (ev, j, moved, minimal)
        if ev == "leave":
            assert (want != j).all()
        assert (p.lookup_many(keys) == want).all(), (ev, j)
        assert (s.counters() == oracle.counts(want, M0 + 4)[: len(s.counters())]).all()
        prev = got.copy()


def test_place_batch_hrw2_policy(gp, oracle):
    """Service::get_or_create_placement batched (service.rs:193-254) with the hierarchical solver as the placing rule; the
    handle's own solver stays flat (place_batch carries its policy)."""
    p = gp.GpuObjectPlacement()
    addrs, seeds, w = oracle.synth_nodes(32)
    p.set_nodes(addrs, w)
    keys = oracle.synth_keys(20000, 3)
    want = oracle.assign_hrw2(keys, seeds, w, threads=4)
    got = p.place_batch(keys, "hrw2")
    assert (got == want).all() and (p.lookup_many(keys) == want).all()
    assert (p.place_batch(keys, "hrw2") == want).all()   # idempotent: everything already placed on live nodes
    p.node_set_active(5, False)
    w2 = w.copy()
    w2[5] = 0
    got2 = p.place_batch(keys[:10000], "hrw2")
    # lazy re-placement (service.rs:224-238): only objects recorded on the dead node are re-placed, the others keep
    # their node even where a fresh assignment would now differ
    fresh = oracle.assign_hrw2(keys[:10000], seeds, w2, threads=4)
    on_dead = want[:10000] == 5
    assert (got2[on_dead] == fresh[on_dead]).all() and (got2[~on_dead] == want[:10000][~on_dead]).all()
    rest = p.lookup_many(keys[10000:])
    assert (rest[want[10000:] == 5] == NONE).all() and (rest[want[10000:] != 5] == want[10000:][want[10000:] != 5]).all()
    r = gp.Resolver(p, policy="hrw2")
    k3 = oracle.synth_keys(200, 9)
    assert [r.resolve(int(k)) for k in k3] == oracle.assign_hrw2(k3, seeds, w2).tolist()
    r.close()


def test_full_size_10m_x_1024_every_object(gp, oracle):
    """BASELINE size, FULL comparison (not a sample): all 10 M placements against the oracle, the node loads against w/W,
    and the leave / join round trip with its movement bound."""
    n, M = 10_000_000, 1024
    p = provider(gp)
    addrs, seeds, w = oracle.synth_nodes(M)
    p.set_nodes(addrs, w)
    s = p.new_set(n)
    s.synth_keys(0, n, 1)
    s.assign()
    idx = s.read()
    cnt = s.counters()
    keys = oracle.synth_keys(n, 1)
    threads = os.cpu_count() or 8
    want = oracle.assign_hrw2(keys, seeds, w, threads=threads)
    assert (idx == want).all()
    assert cnt.sum() == n and (cnt == np.bincount(idx, minlength=M)).all()
    e = n * w / w.sum()
    chi = ((cnt - e) ** 2 / e).sum()
    assert chi < (M - 1) + 6 * np.sqrt(2 * (M - 1)), chi
    p.node_set_active(17, False)
    moved = s.rebalance("leave", 17)
    idx2 = s.read()
    w2 = w.copy()
    w2[17] = 0
    assert (idx2 == oracle.assign_hrw2(keys, seeds, w2, threads=threads)).all()
    assert moved == int((idx != idx2).sum()) and (idx2 != 17).all() and cnt[17] <= moved <= 7.5 * cnt[17]
    p.node_set_active(17, True)
    assert s.rebalance("join", 17) == moved and (s.read() == idx).all()
    # same result through the host-buffer API (H2D/D2H pipelined path)
    assert (p.assign_batch(keys[:3_000_000]) == idx[:3_000_000]).all()


def test_client_first_hop_reaches_the_owner_without_redirect(gp, oracle):
    """SURVEY 8(f) row 2 under HRW2: ids the servers placed with policy "hrw2" are found by the client's own walk of the same
    trie (include/rio_client.h), node for node, also after a node left."""
    from rio_rs_b200 import client as CL

    try:
        CL.lib()
    except Exception as e:
        pytest.skip("librio_client.so unavailable: %r" % (e,))
    p = provider(gp)
    addrs, seeds, w = oracle.synth_nodes(64)
    p.set_nodes(addrs, w)
    keys = oracle.synth_keys(30000, 6)
    owner = p.place_batch(keys, "hrw2")
    fh = CL.FirstHop(addrs, w, policy="hrw2")
    assert (fh.first_hop_batch(keys) == owner).all()
    p.node_set_active(9, False)
    p.rebalance("leave", 9)
    w2 = w.copy()
    w2[9] = 0
    fh.set_active_servers(addrs, w2)
    assert (fh.first_hop_batch(keys) == p.lookup_many(keys)).all()


def test_bounded_calls_in_flight_on_several_sets(gp, oracle):
    """rio_cuda_set_assign_bounded_begin / _end: three resident sets have their pass 0 (walk + histogram + capacity check, one
    kernel each) queued back to back before any check is read; one of them needs spill rounds.  Each must end exactly like the
    one-call form and like the oracle, whatever order the _end calls come in."""
    p = provider(gp)
    addrs, seeds, w = oracle.synth_nodes(64)
    p.set_nodes(addrs, w)
    n = 150_000
    sets, keys = [], []
    for k in range(3):
        s = p.new_set(n)
        s.synth_keys(k * n, n, 5)
        sets.append(s)
        keys.append(oracle.synth_keys(n, 5, first=k * n))
    caps = [(5, 4), (101, 100), (1, 1)]
    for rep in range(3):   # repeated: the two counter buffers of each set take turns, the closed-set epochs advance
        for s, cap in zip(sets, caps):
            s.assign_bounded_begin(0, cap[0], cap[1], 4)
        for k in (2, 0, 1):
            passes = sets[k].assign_bounded_end()
            widx, wcnt, wpass = oracle.assign_bounded_hrw2(keys[k], seeds, w, caps[k][0], caps[k][1], 4, threads=8)
            assert passes == wpass, (rep, k)
            assert (sets[k].read() == widx).all() and (sets[k].counters() == wcnt).all(), (rep, k)
    with pytest.raises(gp.Unknown):
        sets[0].assign_bounded_end()               # nothing in flight
    sets[0].assign_bounded_begin(0, 5, 4, 4)
    with pytest.raises(gp.Unknown):
        sets[0].assign_bounded_begin(0, 5, 4, 4)   # one bounded call per set at a time
    assert sets[0].assign_bounded_end() >= 1


def test_membership_touched_between_the_two_halves_of_a_bounded_call(gp, oracle):
    """What may happen between _begin and _end of a bounded call on a busy provider.  (1) Another caller records a placement on a
    never-seen address (update() may record anything, local.rs:34-36): the node table grows by a non-live entry, the spill rounds of
    the call in flight must neither read past their closed set (found by tests/cpp/abi_fuzz.cpp: a host heap overflow in the masked
    table build) nor change their result.  (2) A server JOINS: the capacities, counters and closed set of the call describe the old
    cluster, so a spill round is refused loudly -- and the same call run again gives the oracle's answer for the new cluster."""
    p = provider(gp)
    addrs, seeds, w = oracle.synth_nodes(65)
    p.set_nodes(addrs[:64], w[:64])
    n = 120_000
    s = p.new_set(n)
    s.synth_keys(0, n, 5)
    keys = oracle.synth_keys(n, 5)
    widx, wcnt, wpass = oracle.assign_bounded_hrw2(keys, seeds[:64], w[:64], 101, 100, 4, threads=8)
    assert wpass > 1                                               # the spill rounds do fire at this cap
    s.assign_bounded_begin(0, 101, 100, 4)
    p.update(gp.ObjectPlacementItem.new(gp.ObjectId.new("Obj", "elsewhere"), "203.0.113.7:9"))   # interns a 65th, non-live address
    assert s.assign_bounded_end() == wpass
    assert (s.read() == widx).all() and (s.counters()[:64] == wcnt).all()
    s.assign_bounded_begin(0, 101, 100, 4)
    p.node_upsert(addrs[64], int(w[64]))                           # a live node joins mid-call
    with pytest.raises(gp.Unknown) as e:
        s.assign_bounded_end()
    assert "live node set changed" in str(e.value)
    # interning order: the 64 nodes, the recorded address (not live), the joiner
    seeds2 = np.concatenate([seeds[:64], np.array([oracle.node_seed("203.0.113.7:9")], dtype=np.uint64), seeds[64:65]])
    w2 = np.concatenate([w[:64], np.zeros(1, dtype=np.uint32), w[64:65]])
    assert seeds2.dtype == np.uint64 and w2.dtype == np.uint32
    widx2, wcnt2, wpass2 = oracle.assign_bounded_hrw2(keys, seeds2, w2, 101, 100, 4, threads=8)
    assert s.assign_bounded(0, 101, 100, 4) == wpass2
    assert (s.read() == widx2).all() and (s.counters() == wcnt2).all()
