"""CPU tests: the directory oracles against the reference's own known-answer tests.

Each test restates one reference test (cited) against BOTH restatements
(oracle/directory_model.cpp for LocalObjectPlacement, oracle/sqlite_model.py for SqliteObjectPlacement),
then the two are cross-checked on random op sequences.  No GPU, no product code.
"""
import random

import pytest

from oracle.sqlite_model import SqliteDirectoryModel


def _providers(oracle):
    return [oracle.DirectoryModel(), SqliteDirectoryModel()]


def test_no_placement(oracle):
    """rio-rs/tests/object_placement_backend.rs:11-16"""
    for p in _providers(oracle):
        p.prepare()
        assert p.lookup("obj", "1") is None


def test_save_and_load(oracle):
    """rio-rs/tests/object_placement_backend.rs:18-34"""
    for p in _providers(oracle):
        p.prepare()
        p.update("obj", "1", "0.0.0.0:8888")
        assert p.lookup("obj", "1") == "0.0.0.0:8888"
        p.clean_server("0.0.0.0:8888")
        assert p.lookup("obj", "1") is None


def test_local_provider_is_clonable_semantics(oracle):
    """rio-rs/src/object_placement/local.rs:75-114 -- a clone shares state; here: one handle, two views."""
    p = oracle.DirectoryModel()
    q = p  # Arc clone == same underlying map
    p.update("test", "1", "0.0.0.0:80")
    assert p.lookup("test", "1") is not None and q.lookup("test", "1") is not None
    q.clean_server("0.0.0.0:80")
    assert p.lookup("test", "1") is None and q.lookup("test", "1") is None


def test_sqlite_overwrite_then_clean(oracle):
    """rio-rs/src/object_placement/sqlite.rs:149-193 -- update :5000 then :5001 => lookup :5001; clean => None"""
    for p in _providers(oracle):
        p.prepare()
        p.update("Test", "1", "0.0.0.0:5000")
        p.update("Test", "1", "0.0.0.0:5001")
        assert p.lookup("Test", "1") == "0.0.0.0:5001"
        p.clean_server("0.0.0.0:5000")  # stale address: nothing to delete
        assert p.lookup("Test", "1") == "0.0.0.0:5001"
        p.clean_server("0.0.0.0:5001")
        assert p.lookup("Test", "1") is None


def test_update_none_and_remove(oracle):
    """local.rs:34-38 (None => remove key) and local.rs:60-68"""
    p = oracle.DirectoryModel()
    p.update("obj", "1", "0.0.0.0:1")
    p.update("obj", "1", None)
    assert p.lookup("obj", "1") is None and len(p) == 0
    p.update("obj", "2", "0.0.0.0:1")
    p.remove("obj", "2")
    p.remove("obj", "2")  # idempotent
    assert p.lookup("obj", "2") is None


def test_local_and_sqlite_models_agree_on_random_ops(oracle):
    rng = random.Random(7)
    a, b = oracle.DirectoryModel(), SqliteDirectoryModel()
    b.prepare()
    addrs = ["10.0.0.%d:5000" % j for j in range(6)]
    ids = [("T%d" % (i % 3), str(i)) for i in range(60)]
    for _ in range(3000):
        op = rng.random()
        t, i = rng.choice(ids)
        if op < 0.45:
            ad = rng.choice(addrs)
            a.update(t, i, ad)
            b.update(t, i, ad)
        elif op < 0.55:
            a.remove(t, i)
            b.remove(t, i)
        elif op < 0.60:
            ad = rng.choice(addrs)
            a.clean_server(ad)
            b.clean_server(ad)
        else:
            assert a.lookup(t, i) == b.lookup(t, i)
    for t, i in ids:
        assert a.lookup(t, i) == b.lookup(t, i)


def test_get_or_create_placement_policy(oracle):
    """rio-rs/src/service.rs:193-254 + the behaviour pinned by rio-rs/tests/object_allocation.rs:75-137:
    unallocated -> claimed by the serving node; owner dies -> next request re-places on the new serving node
    and clean_server drops every object of the dead node."""
    m = oracle.DirectoryModel()
    m.member_push("0.0.0.0", "5000", True)
    m.member_push("0.0.0.0", "5001", True)
    assert m.get_or_create_placement("0.0.0.0:5000", "MockService", "1") == "0.0.0.0:5000"
    assert m.get_or_create_placement("0.0.0.0:5000", "MockService", "2") == "0.0.0.0:5000"
    # a request landing on the other server still resolves to the recorded owner (-> Redirect upstream)
    assert m.get_or_create_placement("0.0.0.0:5001", "MockService", "1") == "0.0.0.0:5000"
    m.member_set_active("0.0.0.0", "5000", False)
    assert m.get_or_create_placement("0.0.0.0:5001", "MockService", "1") == "0.0.0.0:5001"
    assert m.lookup("MockService", "2") is None  # cleaned with its dead server (service.rs:233-237)
    # malformed record is dropped and re-placed (service.rs:213-222)
    m.update("MockService", "3", "garbage")
    assert m.get_or_create_placement("0.0.0.0:5001", "MockService", "3") == "0.0.0.0:5001"


def test_check_address_mismatch_restated(oracle):
    """service.rs:261-298: local -> Ok; active elsewhere -> Redirect; not active -> clean_server + DeallocateServiceObject;
    no ':' -> Unknown("Malformed address: Missing PORT").  `split(':')` takes the first two pieces only."""
    m = oracle.DirectoryModel()
    m.member_push("0.0.0.0", "5000", True)
    m.member_push("0.0.0.0", "5001", True)
    m.member_push("0.0.0.0", "5002", False)
    for i, a in enumerate(["0.0.0.0:5000", "0.0.0.0:5001", "0.0.0.0:5002", "0.0.0.0:5002", "9.9.9.9:1"]):
        m.update("T", str(i), a)
    me = "0.0.0.0:5000"
    assert m.check_address_mismatch(me, me) == oracle.ADDR_LOCAL
    assert m.check_address_mismatch(me, "0.0.0.0:5001") == oracle.ADDR_REDIRECT
    assert m.lookup("T", "2") == "0.0.0.0:5002"
    assert m.check_address_mismatch(me, "0.0.0.0:5002") == oracle.ADDR_DEALLOCATE      # inactive member
    assert m.lookup("T", "2") is None and m.lookup("T", "3") is None                  # clean_server dropped both
    assert m.check_address_mismatch(me, "9.9.9.9:1") == oracle.ADDR_DEALLOCATE         # unknown member == not active
    assert m.lookup("T", "4") is None and m.lookup("T", "1") == "0.0.0.0:5001"
    assert m.check_address_mismatch(me, "garbage") == oracle.ADDR_MALFORMED
    assert m.check_address_mismatch(me, "0.0.0.0:5001:extra") == oracle.ADDR_REDIRECT  # ip "0.0.0.0", port "5001": third piece ignored
    assert m.check_address_mismatch("garbage", "garbage") == oracle.ADDR_LOCAL         # equality is checked before the format
