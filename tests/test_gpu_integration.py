"""The reference's N-server integration tests against the GPU provider: N PlacementService instances on one
provider handle, one killed mid-run, the client following redirects -- rio-rs/tests/object_allocation.rs:75-137,
tests/object_service_error_handling.rs:90-171, tests/client_server_integration_test.rs (1, 2 and 10 servers)."""
import threading

import numpy as np
import pytest

from integration_utils import Client, Cluster

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gp():
    from rio_rs_b200 import build

    build.build()
    import rio_rs_b200 as R

    return R


def test_move_object_on_server_failure(gp):
    """tests/object_allocation.rs:75-137, statement for statement."""
    c = Cluster(gp, 2)
    try:
        client = Client(c, seed=1)
        assert not c.is_allocated("MockService", "1")                       # starts not allocated
        assert client.send("MockService", "1", "OkMessage") == "ok"         # first message allocates it
        assert c.is_allocated("MockService", "1")
        first_server = c.provider.lookup(gp.ObjectId("MockService", "1"))
        client.send("MockService", "1", "KillServer")                       # the owner dies
        assert c.active_members() == [a for a in c.addresses if a != first_server]
        assert client.send("MockService", "1", "OkMessage") == "ok"         # re-allocated somewhere else
        assert c.is_allocated("MockService", "1")
        second_server = c.provider.lookup(gp.ObjectId("MockService", "1"))
        assert first_server != second_server
        assert ("MockService", "1") in c.servers[second_server].registry
    finally:
        c.close()


def test_single_server_and_not_allocated_after_panic(gp):
    """tests/object_service_error_handling.rs:90-171: allocated after Ok, NOT allocated after a handler panic."""
    from rio_rs_b200 import service as S

    c = Cluster(gp, 1)
    try:
        client = Client(c)
        assert client.send("MockService", "ok", "OkMessage") == "ok" and c.is_allocated("MockService", "ok")
        with pytest.raises(S.Unknown):
            client.send("MockService", "boom", "Panic")
        assert not c.is_allocated("MockService", "boom")
        assert ("MockService", "boom") not in c.servers[c.addresses[0]].registry
    finally:
        c.close()


@pytest.mark.parametrize("policy,solver", [("self", None), ("hrw", None), ("hrw2", "hrw2")])
def test_ten_servers_concurrent_clients_and_three_failures(gp, oracle, policy, solver):
    """10 servers (tests/client_server_integration_test.rs:246-251 scale), 8 client threads, 400 objects; three servers are
    killed while requests are in flight.  Afterwards every object is allocated on a live server, it is activated on exactly
    that server's registry, and under the solver policies the owner is the oracle's pick over the surviving members."""
    c = Cluster(gp, 10, policy=policy, solver=solver)
    try:
        ids = [("MockService", str(i)) for i in range(400)]
        errors = []

        def worker(t):
            cl = Client(c, seed=10 + t)
            try:
                for rep in range(3):
                    for k in range(t, len(ids), 8):
                        assert cl.send(*ids[k], "OkMessage") == "ok"
                    if rep == 0 and t < 3:
                        c.kill(c.addresses[2 + 3 * t])          # servers 2, 5, 8 die while the others keep sending
            except Exception as e:  # noqa: BLE001
                errors.append(repr(e))

        th = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        assert not errors, errors[:3]
        live = set(c.active_members())
        assert len(live) == 7
        settle = Client(c, seed=99)
        for t, i in ids:                                        # one more request per object: everything settles on live servers
            settle.send(t, i, "OkMessage")
        owners = [c.provider.lookup(gp.ObjectId(t, i)) for t, i in ids]
        assert all(o in live for o in owners)
        for (t, i), o in zip(ids, owners):
            assert (t, i) in c.servers[o].registry
        if policy != "self":
            addrs, w = c.addresses, np.array([1 if a in live else 0 for a in c.addresses], dtype=np.uint32)
            seeds = np.array([oracle.node_seed(a) for a in addrs], dtype=np.uint64)
            keys = np.array([oracle.object_key(t, i) for t, i in ids], dtype=np.uint64)
            # Flat rendezvous: an object placed at ANY moment of the run sits on the oracle's pick over the FINAL survivors (removing
            # other nodes never changes a surviving winner; objects on a dead node were re-placed by the settle pass).  HRW2 gives
            # up exactly that property for speed: an object placed between two failures keeps the node the solver
            # chose over the members alive then, so only liveness and single activation are asserted for it.
            if policy == "hrw":
                pick = oracle.assign_hrw(keys, seeds, w)
                for k, o in enumerate(owners):
                    assert o == addrs[int(pick[k])], (k, o)
    finally:
        c.close()


def test_client_first_hop_needs_no_redirect(gp, oracle):
    """SURVEY 8(f) row 2 end to end: with the deterministic first hop the client reaches the owner directly; with the
    reference's random pick most first requests are redirected."""
    from rio_rs_b200 import client as CL

    c = Cluster(gp, 8, policy="hrw2", solver="hrw2")
    try:
        fh = CL.FirstHop(c.addresses, policy="hrw2")
        smart = Client(c, first_hop=lambda t, i: fh.get_service_object_address(t, i))
        rnd = Client(c, seed=5)
        for i in range(300):
            smart.send("MockService", "s%d" % i, "OkMessage")
            rnd.send("MockService", "r%d" % i, "OkMessage")
        assert smart.redirects == 0 and smart.attempts == 300
        assert rnd.redirects > 200
    finally:
        c.close()
