"""CPU ORACLE (test infrastructure): SqliteObjectPlacement restated with the reference's own SQL.

Follows rio-rs/src/object_placement/sqlite.rs:68-126 and the schema in
rio-rs/src/object_placement/migrations/0001-sqlite-init.sql:1-9, through Python's sqlite3
(SQLite 3.45) instead of sqlx.  Pinned by the reference's tests (sqlite.rs:149-193,
tests/object_placement_backend.rs:11-34), restated in tests/test_oracle_directory.py.
Never imported by the product.
"""
import sqlite3

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


class SqliteDirectoryModel:
    def __init__(self, path=":memory:"):
        self.db = sqlite3.connect(path)

    def prepare(self):  # sqlite.rs:58-66
        with self.db:
            self.db.executescript(_SCHEMA)

    def update(self, type_, id_, address):  # sqlite.rs:68-85
        with self.db:
            self.db.execute(
                "INSERT INTO object_placement(struct_name, object_id, server_address) VALUES (:p1, :p2, :p3) "
                "ON CONFLICT(struct_name, object_id) DO UPDATE SET server_address=:p3",
                {"p1": type_, "p2": id_, "p3": address},
            )

    def lookup(self, type_, id_):  # sqlite.rs:86-100 (errors are swallowed to None there too)
        row = self.db.execute(
            "SELECT server_address FROM object_placement WHERE struct_name = :p1 and object_id = :p2",
            {"p1": type_, "p2": id_},
        ).fetchone()
        return None if row is None else row[0]

    def clean_server(self, address):  # sqlite.rs:101-112
        with self.db:
            self.db.execute("DELETE FROM object_placement WHERE server_address = :p1", {"p1": address})

    def remove(self, type_, id_):  # sqlite.rs:114-126
        with self.db:
            self.db.execute(
                "DELETE FROM object_placement WHERE struct_name = :p1 and object_id = :p2", {"p1": type_, "p2": id_}
            )
