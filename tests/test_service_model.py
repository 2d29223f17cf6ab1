"""The reference's N-server scenarios (rio-rs/tests/object_allocation.rs:75-137, tests/object_service_error_handling.rs:90-171,
tests/client_server_integration_test.rs) on CPU: the SAME test bodies the GPU box runs against `GpuObjectPlacement`
(tests/test_gpu_integration.py) are run here against the restated reference -- `oracle.DirectoryModel` = LocalObjectPlacement +
LocalStorage + Service::get_or_create_placement / check_address_mismatch, statement for statement -- behind the provider interface.

Two things are checked without a GPU: the host logic of the product's service mirror (rio_rs_b200/service.py: Service::call's
placement half, the Redirect / DeallocateServiceObject / panic paths) and of the harness (tests/integration_utils.py), and that the
restated reference passes the reference's own multi-server tests -- so when the GPU provider passes the very same bodies on the GPU
box, it is being held to behaviour the reference model exhibits too.  The oracle appears here as what it is: the checker.
"""
import threading
import types

import pytest

import test_gpu_integration as G
from rio_rs_b200 import _native as N
from rio_rs_b200 import service as S
from rio_rs_b200.provider import ObjectId


def _split(address):
    ip, _, port = address.partition(":")
    return ip, port


class ModelProvider:
    """The subset of GpuObjectPlacement the service mirror and the harness use, answered by the restated reference."""

    def __init__(self, oracle):
        self._oracle = oracle
        self.model = oracle.DirectoryModel()
        self.mu = threading.Lock()          # the Python wrapper of the model shares one output buffer
        self.known = []

    def clone(self):                        # clones share state (local.rs:12-18)
        return self

    def set_nodes(self, addresses, weights=None):
        assert weights is None
        for a in addresses:
            ip, port = _split(a)
            with self.mu:
                self.model.member_push(ip, port, True)
            self.known.append(a)

    def node_index(self, address):
        return self.known.index(address) if address in self.known else None

    def node_set_active(self, idx, active):
        ip, port = _split(self.known[idx])
        with self.mu:
            self.model.member_set_active(ip, port, bool(active))

    def lookup(self, object_id):
        with self.mu:
            return self.model.lookup(*object_id)

    def remove(self, object_id):
        with self.mu:
            self.model.remove(*object_id)

    def check_address_mismatch(self, self_address, server_address):
        with self.mu:
            return self.model.check_address_mismatch(self_address, server_address)


class ModelResolver:
    """rio_rs_b200.provider.Resolver's per-request call, answered by the restated Service::get_or_create_placement."""

    def __init__(self, provider, policy="hrw", self_address=None, max_batch=4096, max_wait_us=50):
        assert policy == "self", "the reference has one placement rule: the serving node claims the object (service.rs:244-252)"
        self.p, self.me = provider, self_address

    def get_or_create_placement(self, handler_type, handler_id):
        with self.p.mu:
            return self.p.model.get_or_create_placement(self.me, handler_type, handler_id)

    def close(self):
        pass


@pytest.fixture()
def model_backend(oracle, monkeypatch):
    assert (oracle.ADDR_LOCAL, oracle.ADDR_REDIRECT, oracle.ADDR_DEALLOCATE, oracle.ADDR_MALFORMED) == (N.ADDR_LOCAL, N.ADDR_REDIRECT, N.ADDR_DEALLOCATE, N.ADDR_MALFORMED)
    monkeypatch.setattr(S, "Resolver", ModelResolver)
    return types.SimpleNamespace(GpuObjectPlacement=lambda: ModelProvider(oracle), ObjectId=ObjectId)


def test_move_object_on_server_failure_on_the_reference_model(model_backend):
    G.test_move_object_on_server_failure(model_backend)


def test_not_allocated_after_panic_on_the_reference_model(model_backend):
    G.test_single_server_and_not_allocated_after_panic(model_backend)


def test_ten_servers_three_failures_on_the_reference_model(model_backend, oracle):
    G.test_ten_servers_concurrent_clients_and_three_failures(model_backend, oracle, "self", None)


def test_service_mirror_error_paths(model_backend, oracle):
    """Service::call's verdicts one by one (service.rs:54-110, 261-298) through rio_rs_b200/service.py."""
    from integration_utils import Cluster

    c = Cluster(model_backend, 3)
    try:
        a0, a1, a2 = c.addresses
        assert c.servers[a0].call("T", "1", "OkMessage") == "ok"                     # unplaced -> a0 claims it
        with pytest.raises(S.Redirect) as r:                                        # owner is alive elsewhere
            c.servers[a1].call("T", "1", "OkMessage")
        assert r.value.to == a0
        c.provider.model.update("T", "2", "garbage")                                # a malformed record is dropped and re-placed
        assert c.servers[a2].call("T", "2", "OkMessage") == "ok" and c.provider.lookup(ObjectId("T", "2")) == a2
        c.kill(a0)
        with pytest.raises(S.ServerNotAvailable):                                   # what a client sees when it dials a dead server
            c.servers[a0].call("T", "1", "OkMessage")
        assert c.servers[a1].call("T", "1", "OkMessage") == "ok"                     # the next server to see it re-places it
        assert c.provider.lookup(ObjectId("T", "1")) == a1 and ("T", "1") in c.servers[a1].registry
        with pytest.raises(S.Unknown):                                              # handler panic -> remove (service.rs:92-106)
            c.servers[a1].call("T", "3", "Panic")
        assert c.provider.lookup(ObjectId("T", "3")) is None and ("T", "3") not in c.servers[a1].registry
    finally:
        c.close()
