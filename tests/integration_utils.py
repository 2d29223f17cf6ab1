"""The reference's multi-node test harness (rio-rs/tests/server_utils.rs:49-139) against the GPU provider: N in-process "servers"
(rio_rs_b200.service.PlacementService) share ONE provider handle (clones) and ONE membership view, a client follows Redirect /
DeallocateServiceObject like the reference's retry loop (client/tower_services.rs:134-225), and a server can be killed mid-run."""
import random
import threading

from rio_rs_b200 import service as S


class Cluster:
    """run_integration_test: `num_servers` servers on addresses 0.0.0.0:50xx sharing provider + membership."""

    def __init__(self, gp, num_servers, policy="self", weights=None, handlers=None, solver=None):
        self.gp = gp
        self.provider = gp.GpuObjectPlacement()
        if solver:
            self.provider.set_solver(solver)
        self.addresses = ["0.0.0.0:%d" % (5000 + j) for j in range(num_servers)]
        self.provider.set_nodes(self.addresses, weights)        # MembershipStorage: every server pushed itself as active
        self._mu = threading.Lock()
        self.active = set(self.addresses)
        base = {"OkMessage": lambda svc, t, i: "ok", "KillServer": self._kill_handler, "Panic": self._panic_handler}
        base.update(handlers or {})
        self.servers = {a: S.PlacementService(self.provider, a, base, policy=policy) for a in self.addresses}

    def close(self):
        for s in self.servers.values():
            s.close()

    # handlers ------------------------------------------------------------------------------------------------
    def _kill_handler(self, svc, t, i):
        """tests/object_allocation.rs:45-62: the handler kills its own server; the gossip marks it inactive."""
        self.kill(svc.address)
        return "dying"

    @staticmethod
    def _panic_handler(svc, t, i):
        raise RuntimeError("handler panic")

    def kill(self, address):
        self.servers[address].alive = False
        with self._mu:
            self.active.discard(address)
        self.provider.node_set_active(self.provider.node_index(address), False)    # set_inactive, peer_to_peer.rs:170-173

    def active_members(self):
        with self._mu:
            return sorted(self.active)

    def is_allocated(self, t, i):
        """tests/server_utils.rs:106-114"""
        return self.provider.lookup(self.gp.ObjectId(t, i)) is not None


class Client:
    """The reference client's send loop: random active server first (client/mod.rs:254-263), Redirect -> go there,
    DeallocateServiceObject / dead server -> pick again; at most `retries` attempts (tower_services.rs:142-146)."""

    def __init__(self, cluster, seed=0, first_hop=None, retries=20):
        self.c = cluster
        self.rng = random.Random(seed)
        self.first_hop = first_hop
        self.retries = retries
        self.redirects = 0
        self.attempts = 0

    def send(self, t, i, message):
        target = None
        for _ in range(self.retries):
            if target is None:
                members = self.c.active_members()
                if not members:
                    raise S.ServerNotAvailable("no servers")
                target = self.first_hop(t, i) if self.first_hop else self.rng.choice(members)
            self.attempts += 1
            try:
                return self.c.servers[target].call(t, i, message)
            except S.Redirect as r:
                self.redirects += 1
                target = r.to
            except (S.DeallocateServiceObject, S.ServerNotAvailable):
                target = None
        raise S.Unknown("retries exhausted")
