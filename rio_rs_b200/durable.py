"""Durable write-through of the GPU directory into the reference's own SQL schema.

The GPU directory is volatile; rio-rs deployments keep placements in SQLite/Postgres
(rio-rs/src/object_placement/sqlite.rs:68-126, schema migrations/0001-sqlite-init.sql:1-9).  This provider keeps that table
as the source of truth: every mutation is applied to the GPU directory and written through with the reference's SQL
(batched calls use one transaction per batch), and `recover()` bulk-loads the table back after a restart:
rows -> packed "{type}.{id}" bytes -> k_hash_ids on the GPU -> upsert_batch.
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
    def update_many_ids(self, ids, addresses):
        """ids: [(type, id)], addresses: [str | None]; one GPU upsert + one SQL transaction."""
        keys = self.hash_ids(ids)
        idx = np.array([N.NONE if a is None else self.node_intern(a) for a in addresses], dtype=np.uint32)
        self.update_many(keys, idx)
        with self.db:
            self.db.executemany(_UPSERT, [{"p1": t, "p2": i, "p3": a} for (t, i), a in zip(ids, addresses) if a is not None])
            self.db.executemany("DELETE FROM object_placement WHERE struct_name = :p1 and object_id = :p2",
                                [{"p1": t, "p2": i} for (t, i), a in zip(ids, addresses) if a is None])
        return keys

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
            self.update_many(keys, idx)
            total += len(rows)
        return total
