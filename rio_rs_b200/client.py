"""Client-side deterministic first hop: ctypes binding of include/rio_client.h plus a
mirror of the piece of the reference client it replaces.

    FirstHop.get_service_object_address  <->  Client::get_service_object_address   rio-rs/src/client/mod.rs:235-267
    FirstHop.set_active_servers          <->  Client::fetch_active_servers         rio-rs/src/client/mod.rs:153-172
    FirstHop.record_redirect             <->  the Redirect arm of the retry loop   rio-rs/src/client/tower_services.rs:158-168

The reference picks a uniformly random active server on a cache miss (client/mod.rs:254-263) and lets the server
answer Redirect; here the miss is resolved with the same weighted rendezvous hash the servers' solver uses, so an object
placed with policy "hrw" is reached on the first hop.  CPU only by nature (clients have no GPU); the server-side product
(librio_cuda.so) never loads this module or its library.
"""
import collections
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "librio_client.so")
SRC = os.path.join(HERE, "csrc", "client.cpp")
DEPS = [SRC, os.path.join(HERE, "csrc", "spec.cuh"), os.path.join(HERE, "csrc", "trie_table.hpp"), os.path.join(HERE, "..", "include", "rio_client.h")]
NONE = 0xFFFFFFFF
_lib = None

u32p, u64p, szp = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_size_t)
# name -> (restype, argtypes); every function declared in include/rio_client.h (tests/test_client_first_hop.py checks)
SIGNATURES = {
    "rio_client_ring_create": (C.c_int32, [C.POINTER(C.c_char_p), szp, u32p, C.c_uint32, C.POINTER(C.c_void_p)]),
    "rio_client_ring_destroy": (None, [C.c_void_p]),
    "rio_client_ring_set_policy": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "rio_client_ring_size": (C.c_uint32, [C.c_void_p]),
    "rio_client_ring_address": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_char_p, C.c_size_t, szp]),
    "rio_client_object_key": (C.c_uint64, [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]),
    "rio_client_first_hop": (C.c_int32, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, u32p]),
    "rio_client_first_hop_key": (C.c_int32, [C.c_void_p, C.c_uint64, u32p]),
    "rio_client_first_hop_batch": (C.c_int32, [C.c_void_p, u64p, C.c_size_t, u32p]),
}


class NoServersAvailable(Exception):
    """ClientError::NoServersAvailable (rio-rs/src/client/mod.rs:260-261)."""


def build(force=False):
    """g++ -shared of csrc/client.cpp (plain C++, no CUDA) into rio_rs_b200/librio_client.so."""
    if not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in DEPS):
        return SO
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra", "-o", SO, SRC])
    return SO


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def object_key(type_name, object_id):
    t, i = type_name.encode(), object_id.encode()
    return int(lib().rio_client_object_key(t, len(t), i, len(i)))


class FirstHop:
    """The client's placement cache + first-hop pick.  `cache_size` = the reference's LruCache limit (client/mod.rs:137)."""

    def __init__(self, addresses=(), weights=None, cache_size=1000, policy="hrw", trie_bits=0):
        self._policy, self._bits = (2 if policy == "hrw2" else 1), trie_bits
        self._ring = C.c_void_p()
        self._cache = collections.OrderedDict()
        self._cache_size = cache_size
        self.addresses = []
        self.set_active_servers(addresses, weights)

    def close(self):
        if self._ring:
            lib().rio_client_ring_destroy(self._ring)
            self._ring = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_active_servers(self, addresses, weights=None):
        """Replaces the whole view, like fetch_active_servers ("not an incremental operation", client/mod.rs:150-152)."""
        addresses = [a if isinstance(a, str) else a.decode() for a in addresses]
        enc = [a.encode() for a in addresses]
        n = len(enc)
        arr = (C.c_char_p * max(n, 1))(*enc)
        lens = (C.c_size_t * max(n, 1))(*[len(e) for e in enc])
        w = None
        if weights is not None:
            w = np.ascontiguousarray(weights, dtype=np.uint32)
            if w.shape != (n,):
                raise ValueError("one weight per address")
        ring = C.c_void_p()
        st = lib().rio_client_ring_create(arr, lens, w.ctypes.data_as(u32p) if w is not None else None, n, C.byref(ring))
        if st != 0:
            raise ValueError("rio_client_ring_create failed")
        if lib().rio_client_ring_set_policy(ring, self._policy, self._bits) != 0:
            raise ValueError("rio_client_ring_set_policy failed")
        self.close()
        self._ring, self.addresses = ring, addresses

    # -- the pick --------------------------------------------------------------------------------------------
    def first_hop_index(self, type_name, object_id):
        t, i = type_name.encode(), object_id.encode()
        out = C.c_uint32(NONE)
        if lib().rio_client_first_hop(self._ring, t, len(t), i, len(i), C.byref(out)) != 0:
            raise ValueError("rio_client_first_hop failed")
        return int(out.value)

    def first_hop_batch(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.empty(len(keys), dtype=np.uint32)
        if lib().rio_client_first_hop_batch(self._ring, keys.ctypes.data_as(u64p), len(keys), out.ctypes.data_as(u32p)) != 0:
            raise ValueError("rio_client_first_hop_batch failed")
        return out

    def get_service_object_address(self, type_name, object_id):
        """Cached address if any (client/mod.rs:251-253), else the rendezvous owner instead of a random server."""
        k = (type_name, object_id)
        if k in self._cache:
            self._cache.move_to_end(k)
            return self._cache[k]
        j = self.first_hop_index(type_name, object_id)
        if j == NONE:
            raise NoServersAvailable()
        return self.addresses[j]

    def record_redirect(self, type_name, object_id, address):
        """The server said Redirect(address) (tower_services.rs:158-168).  Same key order as the lookup above: the
        reference inserts (type, id) but looks up (id, type) (client/mod.rs:241-244), so its cache never hits."""
        k = (type_name, object_id)
        self._cache[k] = address
        self._cache.move_to_end(k)
        while len(self._cache) > self._cache_size:
            self._cache.popitem(last=False)
