"""Durable write-through of the GPU directory into the reference's own SQL schema.

The GPU directory is volatile; rio-rs deployments keep placements in SQLite/Postgres
(rio-rs/src/object_placement/sqlite.rs:68-126, schema migrations/0001-sqlite-init.sql:1-9).  This provider keeps that table
as the source of truth: every mutation is applied to the GPU directory and written through with the reference's SQL
(batched calls use one transaction per batch), and `recover()` bulk-loads the table back after a restart:
rows -> packed "{type}.{id}" bytes -> k_hash_ids on the GPU -> upsert_batch.

Guarantee: every mutation made THROUGH THIS CLASS reaches the table in the same call.  The SQL schema is keyed by the id
strings while the GPU directory is keyed by their 64-bit hash, so the key-only batched calls of the base class
(update_many / remove_many / place_batch / ObjectSet.commit on raw keys) cannot be written through and raise here; their
id-carrying forms (update_many_ids, remove_many_ids, place_batch_ids) do both sides in one transaction, and `rebalance`
re-reads the placement of every durable row from the GPU afterwards.
"""
import sqlite3

import numpy as np

from . import _native as N
from .provider import GpuObjectPlacement

_SCHEMA = """
CREATE TABLE IF NOT EXISTS object_placement
(
    struct_name     TEXT                NOT NULL,
    object_id       TEXT                NOT NULL,
    server_address  TEXT                NULL,

    PRIMARY KEY (struct_name, object_id)
);
CREATE INDEX IF NOT EXISTS idx_object_placement_server_address on object_placement(server_address);
"""
_UPSERT = ("INSERT INTO object_placement(struct_name, object_id, server_address) VALUES (:p1, :p2, :p3) "
           "ON CONFLICT(struct_name, object_id) DO UPDATE SET server_address=:p3")   # sqlite.rs:72-78


class DurableGpuObjectPlacement(GpuObjectPlacement):
    def __init__(self, db_path=":memory:", **kw):
        super().__init__(**kw)
        self.db = sqlite3.connect(db_path)
        self._active = {}

    def prepare(self):  # sqlite.rs:58-66, then warm the GPU directory from the table
        with self.db:
            self.db.executescript(_SCHEMA)
        return self.recover()

    # ---- the trait, written through -------------------------------------------------------------------------
    def update(self, item):
        super().update(item)
        t, i = item.object_id
        with self.db:
            if item.server_address is None:   # LocalObjectPlacement semantics: update(None) removes the key (local.rs:34-38)
                self.db.execute("DELETE FROM object_placement WHERE struct_name = :p1 and object_id = :p2", {"p1": t, "p2": i})
            else:
                self.db.execute(_UPSERT, {"p1": t, "p2": i, "p3": item.server_address})

    def remove(self, object_id):
        super().remove(object_id)
        with self.db:
            self.db.execute("DELETE FROM object_placement WHERE struct_name = :p1 and object_id = :p2", {"p1": object_id[0], "p2": object_id[1]})   # sqlite.rs:115-125

    def clean_server(self, address):
        super().clean_server(address)
        with self.db:
            self.db.execute("DELETE FROM object_placement WHERE server_address = :p1", {"p1": address})   # sqlite.rs:102-110 (uses the address index)

    # ---- batched -----------------------------------------------------------------------------------------------
    _DELETE_ID = "DELETE FROM object_placement WHERE struct_name = :p1 and object_id = :p2"

    def _write_rows(self, ids, addresses):
        """Apply (id -> address | None) to the table with the GPU's rule for duplicates: the LAST occurrence of an id wins."""
        last = {}
        for k, oid in enumerate(ids):
            last[tuple(oid)] = k
        ups = [{"p1": t, "p2": i, "p3": addresses[k]} for (t, i), k in last.items() if addresses[k] is not None]
        dels = [{"p1": t, "p2": i} for (t, i), k in last.items() if addresses[k] is None]
        self.db.executemany(_UPSERT, ups)
        self.db.executemany(self._DELETE_ID, dels)

    def update_many_ids(self, ids, addresses):
        """ids: [(type, id)], addresses: [str | None]; one GPU upsert + one SQL transaction."""
        keys = self.hash_ids(ids)
        idx = np.array([N.NONE if a is None else self.node_intern(a) for a in addresses], dtype=np.uint32)
        GpuObjectPlacement.update_many(self, keys, idx)
        with self.db:
            self._write_rows(ids, addresses)
        return keys

    def remove_many_ids(self, ids):
        return self.update_many_ids(ids, [None] * len(ids))

    def place_batch_ids(self, ids, policy="hrw", self_address=None):
        """Service::get_or_create_placement for a batch of ids (service.rs:193-254), written through: the rows that were
        (re)placed are upserted and the servers the call cleaned (recorded but inactive) are deleted by address, one transaction."""
        keys = self.hash_ids(ids)
        before = self.lookup_many(keys)
        out = GpuObjectPlacement.place_batch(self, keys, policy, self_address)
        changed = np.nonzero(before != out)[0]
        cleaned = [j for j in np.unique(before[before != N.NONE]) if not self._node_is_active(int(j))]
        with self.db:
            for j in cleaned:   # clean_server(address) of every inactive node the batch met (service.rs:233-237)
                self.db.execute("DELETE FROM object_placement WHERE server_address = :p1", {"p1": self.node_address(int(j))})
            self._write_rows([ids[k] for k in changed], [self.node_address(int(out[k])) for k in changed])
        return out

    def _node_is_active(self, idx):
        a = self.node_address(idx)
        return ":" in a and self._active.get(a, False)

    def set_nodes(self, addresses, weights=None, feats=None):
        self._active = {a: True for a in addresses}
        return GpuObjectPlacement.set_nodes(self, addresses, weights, feats)

    def node_upsert(self, address, weight=1, feat=None):
        self._active[address] = True
        return GpuObjectPlacement.node_upsert(self, address, weight, feat)

    def node_set_active(self, idx, active):
        self._active[self.node_address(idx)] = bool(active)
        return GpuObjectPlacement.node_set_active(self, idx, active)

    def rebalance(self, event, idx):
        """Eager re-placement on the GPU, then the table follows: every durable row is looked up again (bulk: device-side id
        hashing + batched lookup) and rewritten where its node changed."""
        moved = GpuObjectPlacement.rebalance(self, event, idx)
        self.sync_table_from_gpu()
        return moved

    def sync_table_from_gpu(self, batch=1_000_000):
        cur = self.db.execute("SELECT struct_name, object_id, server_address FROM object_placement")
        todo = []
        while True:
            rows = cur.fetchmany(batch)
            if not rows:
                break
            now = self.lookup_many(self.hash_ids([(t, i) for t, i, _ in rows]))
            for (t, i, a), j in zip(rows, now):
                b = None if j == N.NONE else self.node_address(int(j))
                if a != b:
                    todo.append(((t, i), b))
        with self.db:
            self._write_rows([x[0] for x in todo], [x[1] for x in todo])
        return len(todo)

    # key-only mutators of the base class cannot reach the id-keyed table
    def _no_ids(self, *a, **k):
        from .provider import Unknown

        raise Unknown("DurableGpuObjectPlacement writes through an id-keyed SQL table: use the *_ids form of this call")

    update_many = remove_many = place_batch = _no_ids

    def new_set(self, capacity):
        self._no_ids()

    def recover(self, batch=1_000_000):
        """Bulk-load the durable table into the (empty) GPU directory; returns the number of placements restored."""
        cur = self.db.execute("SELECT struct_name, object_id, server_address FROM object_placement WHERE server_address IS NOT NULL")
        total = 0
        interned = {}
        while True:
            rows = cur.fetchmany(batch)
            if not rows:
                break
            keys = self.hash_ids([(t, i) for t, i, _ in rows])
            idx = np.empty(len(rows), dtype=np.uint32)
            for k, (_, _, a) in enumerate(rows):
                j = interned.get(a)
                if j is None:
                    j = interned[a] = self.node_intern(a)
                idx[k] = j
            GpuObjectPlacement.update_many(self, keys, idx)
            total += len(rows)
        return total
