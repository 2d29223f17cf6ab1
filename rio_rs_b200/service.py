"""Host-side mirror of the placement half of the reference's `Service` (rio-rs/src/service.rs:54-110, 193-298) over the GPU
provider: what a rio-rs server does per request BEFORE it dispatches to the registry --

    get_or_create_placement(type, id)      service.rs:193-254   -> coalesced through the Resolver (one place_batch per micro-batch)
    check_address_mismatch(address)        service.rs:261-298   -> Ok | Redirect(address) | DeallocateServiceObject | Unknown
    handler panic -> provider.remove(id)   service.rs:92-106

It exists so that the N-server integration tests of the reference (tests/server_utils.rs:49-102, tests/object_allocation.rs) can
run against `GpuObjectPlacement` without the Rust runtime: same decisions, same error names.  Not a transport, not a registry.
"""
import threading

from . import _native as N
from .provider import ObjectId, Resolver


class ResponseError(Exception):
    """protocol.rs:78-105"""


class Redirect(ResponseError):
    def __init__(self, to):
        super().__init__(to)
        self.to = to


class DeallocateServiceObject(ResponseError):
    pass


class Unknown(ResponseError):
    pass


class ServerNotAvailable(ResponseError):
    """what the client sees when the TCP connection to a dead server fails (client/mod.rs:174-220)"""


class PlacementService:
    """One server's `Service`: shares the provider (a clone) and the membership view with its siblings."""

    def __init__(self, provider, address, handlers, policy="self", max_batch=256, max_wait_us=50):
        self.provider = provider.clone()
        self.address = address
        self.handlers = handlers                      # message name -> callable(service, type, id) -> response
        self.resolver = Resolver(self.provider, policy=policy, self_address=address, max_batch=max_batch, max_wait_us=max_wait_us)
        self.registry = set()                         # objects activated on this server (registry/mod.rs)
        self._mu = threading.Lock()
        self.alive = True

    def close(self):
        self.resolver.close()

    def call(self, handler_type, handler_id, message):
        """Service::call, service.rs:54-110 (placement part + handler dispatch)."""
        if not self.alive:
            raise ServerNotAvailable(self.address)
        server_address = self.resolver.get_or_create_placement(handler_type, handler_id)                    # :59-61
        verdict = self.provider.check_address_mismatch(self.address, server_address)                        # :62
        if verdict == N.ADDR_REDIRECT:
            raise Redirect(server_address)
        if verdict == N.ADDR_DEALLOCATE:
            with self._mu:
                self.registry.discard((handler_type, handler_id))
            raise DeallocateServiceObject()
        if verdict == N.ADDR_MALFORMED:
            raise Unknown("Malformed address: Missing PORT in '%s'" % server_address)
        with self._mu:
            self.registry.add((handler_type, handler_id))                                                   # start_service_object :304-359
        try:
            return self.handlers[message](self, handler_type, handler_id)
        except ResponseError:
            raise
        except Exception:                                                                                   # panic: :92-106
            with self._mu:
                self.registry.discard((handler_type, handler_id))
            self.provider.remove(ObjectId(handler_type, handler_id))
            raise Unknown("Panic")
