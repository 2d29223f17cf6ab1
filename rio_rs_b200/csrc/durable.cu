// durable.cu -- durable write-through of the GPU directory into the reference's own SQLite schema, below the language
// bindings.  Host code only; built on the public C ABI of this library plus libsqlite3.so.0, which is
// dlopen'ed (the image ships the shared object without headers, and the product keeps no link-time dependency on it).
//
// Reference restated: SqliteObjectPlacement, rio-rs/src/object_placement/sqlite.rs:58-126 (the SQL text of every statement
// below is the reference's, bound in the same order) over the table of
// rio-rs/src/object_placement/migrations/0001-sqlite-init.sql:1-9.  The table is the source of truth across restarts; the GPU
// directory in front of it answers every lookup.  update / remove / clean_server are applied to the table first (inside a
// transaction for the batched call) and to the GPU directory second, so a crash between the two leaves the durable side ahead,
// never behind; place_batch is decided by the GPU and then written to the table in one transaction; recover() rebuilds the GPU
// side from the table (device-side id hashing + batched upsert).
#include "../../include/rio_cuda.h"

#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

// ---- the few libsqlite3 entry points used, bound at run time ---------------------------------------------------------
struct sqlite3;
struct sqlite3_stmt;
constexpr int kSqliteOk = 0, kSqliteRow = 100, kSqliteDone = 101;
typedef void (*sqlite_destructor)(void *);
#define RIO_SQLITE_TRANSIENT ((sqlite_destructor)-1)

struct SqliteApi {
    void *lib = nullptr;
    int (*open)(const char *, sqlite3 **) = nullptr;
    int (*close)(sqlite3 *) = nullptr;
    int (*exec)(sqlite3 *, const char *, int (*)(void *, int, char **, char **), void *, char **) = nullptr;
    int (*prepare_v2)(sqlite3 *, const char *, int, sqlite3_stmt **, const char **) = nullptr;
    int (*bind_text)(sqlite3_stmt *, int, const char *, int, sqlite_destructor) = nullptr;
    int (*bind_null)(sqlite3_stmt *, int) = nullptr;
    int (*step)(sqlite3_stmt *) = nullptr;
    int (*reset)(sqlite3_stmt *) = nullptr;
    int (*finalize)(sqlite3_stmt *) = nullptr;
    const unsigned char *(*column_text)(sqlite3_stmt *, int) = nullptr;
    int (*column_bytes)(sqlite3_stmt *, int) = nullptr;
    const char *(*errmsg)(sqlite3 *) = nullptr;
    std::string load_error;
    bool load() {
        if (lib) return true;
        const char *names[] = {getenv("RIO_SQLITE_LIB"), "libsqlite3.so.0", "libsqlite3.so"};
        for (const char *nm : names) {
            if (!nm || !*nm) continue;
            lib = dlopen(nm, RTLD_NOW);
            if (lib) break;
            load_error = dlerror();
        }
        if (!lib) return false;
#define RIO_SYM(field, name) field = (decltype(field))dlsym(lib, name); if (!field) { load_error = std::string("libsqlite3 lacks ") + name; lib = nullptr; return false; }
        RIO_SYM(open, "sqlite3_open") RIO_SYM(close, "sqlite3_close") RIO_SYM(exec, "sqlite3_exec") RIO_SYM(prepare_v2, "sqlite3_prepare_v2")
        RIO_SYM(bind_text, "sqlite3_bind_text") RIO_SYM(bind_null, "sqlite3_bind_null") RIO_SYM(step, "sqlite3_step") RIO_SYM(reset, "sqlite3_reset")
        RIO_SYM(finalize, "sqlite3_finalize") RIO_SYM(column_text, "sqlite3_column_text") RIO_SYM(column_bytes, "sqlite3_column_bytes") RIO_SYM(errmsg, "sqlite3_errmsg")
#undef RIO_SYM
        return true;
    }
};
SqliteApi g_sql;
std::mutex g_sql_mu;
thread_local std::string t_durable_error;

// migrations/0001-sqlite-init.sql:1-9, verbatim
const char *kMigration =
    "CREATE TABLE IF NOT EXISTS object_placement\n"
    "(\n"
    "    struct_name     TEXT                NOT NULL,\n"
    "    object_id       TEXT                NOT NULL,\n"
    "    server_address  TEXT                NULL,\n"
    "\n"
    "    PRIMARY KEY (struct_name, object_id)\n"
    ");\n"
    "CREATE INDEX IF NOT EXISTS idx_object_placement_server_address on object_placement(server_address);";
// sqlite.rs:72-78, :102-110, :115-125
const char *kUpsert = "INSERT INTO object_placement(struct_name, object_id, server_address) VALUES ($1, $2, $3) "
                      "ON CONFLICT(struct_name, object_id) DO UPDATE SET server_address=$3";
const char *kCleanServer = "DELETE FROM object_placement WHERE server_address = $1";
const char *kRemove = "DELETE FROM object_placement WHERE struct_name = $1 and object_id = $2";
const char *kSelectAll = "SELECT struct_name, object_id, server_address FROM object_placement WHERE server_address IS NOT NULL";

}  // namespace

struct rio_durable {
    rio_placement *h = nullptr;
    sqlite3 *db = nullptr;
    sqlite3_stmt *st_upsert = nullptr, *st_clean = nullptr, *st_remove = nullptr;
    std::mutex mu;   // one writer at a time: SQLite serialises writers anyway, and the table / GPU pair must change together
};

namespace {

rio_status fail(rio_durable *d, rio_status code, const std::string &what) {
    t_durable_error = what + (d && d->db && g_sql.errmsg ? std::string(": ") + g_sql.errmsg(d->db) : std::string());
    return code;
}
rio_status gpu_fail(rio_durable *d, rio_status st) {
    const char *m = rio_cuda_last_error(d->h);
    t_durable_error = m ? m : "";
    return st;
}
bool exec(rio_durable *d, const char *sql) { return g_sql.exec(d->db, sql, nullptr, nullptr, nullptr) == kSqliteOk; }

bool run(rio_durable *d, sqlite3_stmt *st, const char *a, size_t al, const char *b, size_t bl, const char *c, size_t cl, int nargs) {
    g_sql.reset(st);
    const char *v[3] = {a, b, c};
    const size_t l[3] = {al, bl, cl};
    for (int i = 0; i < nargs; i++) {
        const int rc = v[i] ? g_sql.bind_text(st, i + 1, v[i], (int)l[i], RIO_SQLITE_TRANSIENT) : g_sql.bind_null(st, i + 1);
        if (rc != kSqliteOk) return false;
    }
    const int rc = g_sql.step(st);
    g_sql.reset(st);
    (void)d;
    return rc == kSqliteDone || rc == kSqliteRow;
}

}  // namespace

extern "C" {

const char *rio_cuda_durable_last_error(void) { return t_durable_error.c_str(); }

/* SqliteObjectPlacement::prepare (sqlite.rs:58-66): open the database and run the migration inside one transaction. */
rio_status rio_cuda_durable_open(rio_placement *h, const char *path, rio_durable **out) {
    if (!h || !path || !out) { t_durable_error = "null argument"; return RIO_ERR_UNKNOWN; }
    *out = nullptr;
    {
        std::lock_guard<std::mutex> g(g_sql_mu);
        if (!g_sql.load()) { t_durable_error = "cannot load libsqlite3: " + g_sql.load_error; return RIO_ERR_UPSTREAM; }
    }
    rio_durable *d = new rio_durable();
    d->h = h;
    if (g_sql.open(path, &d->db) != kSqliteOk) { rio_status st = fail(d, RIO_ERR_UPSTREAM, "sqlite3_open"); if (d->db) g_sql.close(d->db); delete d; return st; }
    if (!exec(d, "BEGIN") || !exec(d, kMigration) || !exec(d, "COMMIT")) { rio_status st = fail(d, RIO_ERR_UPSTREAM, "migration"); g_sql.close(d->db); delete d; return st; }
    if (g_sql.prepare_v2(d->db, kUpsert, -1, &d->st_upsert, nullptr) != kSqliteOk || g_sql.prepare_v2(d->db, kCleanServer, -1, &d->st_clean, nullptr) != kSqliteOk ||
        g_sql.prepare_v2(d->db, kRemove, -1, &d->st_remove, nullptr) != kSqliteOk) {
        rio_status st = fail(d, RIO_ERR_UPSTREAM, "prepare");
        g_sql.close(d->db);
        delete d;
        return st;
    }
    *out = d;
    return RIO_OK;
}

void rio_cuda_durable_close(rio_durable *d) {
    if (!d) return;
    for (sqlite3_stmt *s : {d->st_upsert, d->st_clean, d->st_remove}) if (s) g_sql.finalize(s);
    if (d->db) g_sql.close(d->db);
    delete d;
}

/* Bulk-load the table into the GPU directory (after a restart): rows -> packed "{type}.{id}" bytes -> device-side hashing ->
 * batched upsert, `batch` rows at a time. */
rio_status rio_cuda_durable_recover(rio_durable *d, uint64_t *out_rows) {
    if (!d) { t_durable_error = "null handle"; return RIO_ERR_UNKNOWN; }
    std::lock_guard<std::mutex> g(d->mu);
    sqlite3_stmt *sel = nullptr;
    if (g_sql.prepare_v2(d->db, kSelectAll, -1, &sel, nullptr) != kSqliteOk) return fail(d, RIO_ERR_UPSTREAM, "prepare select");
    const size_t batch = 1u << 20;
    uint64_t total = 0;
    std::string packed;
    std::vector<uint64_t> offs{0}, keys;
    std::vector<uint32_t> idx;
    std::unordered_map<std::string, uint32_t> interned;
    rio_status st = RIO_OK;
    auto flush = [&]() -> rio_status {
        if (idx.empty()) return RIO_OK;
        packed.append(16, '\0');   // k_hash_ids reads whole 16-byte vectors
        keys.resize(idx.size());
        rio_status s = rio_cuda_hash_ids(d->h, packed.data(), offs.data(), idx.size(), keys.data());
        if (s == RIO_OK) s = rio_cuda_upsert_batch(d->h, keys.data(), idx.data(), idx.size());
        total += idx.size();
        packed.clear(); offs.assign(1, 0); idx.clear();
        return s;
    };
    for (;;) {
        const int rc = g_sql.step(sel);
        if (rc == kSqliteDone) break;
        if (rc != kSqliteRow) { st = fail(d, RIO_ERR_UPSTREAM, "select"); break; }
        const char *t = (const char *)g_sql.column_text(sel, 0), *i = (const char *)g_sql.column_text(sel, 1), *a = (const char *)g_sql.column_text(sel, 2);
        packed.append(t, g_sql.column_bytes(sel, 0)).append(1, '.').append(i, g_sql.column_bytes(sel, 1));
        offs.push_back(packed.size());
        const std::string addr(a, g_sql.column_bytes(sel, 2));
        auto it = interned.find(addr);
        if (it == interned.end()) {
            uint32_t j = RIO_NONE;
            st = rio_cuda_node_intern(d->h, addr.c_str(), &j);
            if (st != RIO_OK) { gpu_fail(d, st); break; }
            it = interned.emplace(addr, j).first;
        }
        idx.push_back(it->second);
        if (idx.size() >= batch && (st = flush()) != RIO_OK) { gpu_fail(d, st); break; }
    }
    if (st == RIO_OK && (st = flush()) != RIO_OK) gpu_fail(d, st);
    g_sql.finalize(sel);
    if (out_rows) *out_rows = total;
    return st;
}

/* update (sqlite.rs:68-85): address == NULL stores "no placement" -- LocalObjectPlacement's update(None) removes the key
 * (local.rs:34-38); the table follows that rule (a NULL row would make the reference's own lookup panic, sqlite.rs:99). */
rio_status rio_cuda_durable_update(rio_durable *d, const char *type, size_t type_len, const char *id, size_t id_len, const char *address, size_t address_len) {
    if (!d || !type || !id) { t_durable_error = "null argument"; return RIO_ERR_UNKNOWN; }
    std::lock_guard<std::mutex> g(d->mu);
    const bool ok = address ? run(d, d->st_upsert, type, type_len, id, id_len, address, address_len, 3) : run(d, d->st_remove, type, type_len, id, id_len, nullptr, 0, 2);
    if (!ok) return fail(d, RIO_ERR_UPSTREAM, "update");
    const rio_status st = rio_cuda_update_str(d->h, type, type_len, id, id_len, address, address_len);
    return st == RIO_OK ? st : gpu_fail(d, st);
}

/* lookup (sqlite.rs:86-100): answered by the GPU directory, the cache of the table */
rio_status rio_cuda_durable_lookup(rio_durable *d, const char *type, size_t type_len, const char *id, size_t id_len, char *buf, size_t cap, size_t *out_len) {
    if (!d) { t_durable_error = "null handle"; return RIO_ERR_UNKNOWN; }
    const rio_status st = rio_cuda_lookup_str(d->h, type, type_len, id, id_len, buf, cap, out_len);
    return st == RIO_OK ? st : gpu_fail(d, st);
}

/* clean_server (sqlite.rs:101-112): one DELETE through idx_object_placement_server_address, one scan on the GPU */
rio_status rio_cuda_durable_clean_server(rio_durable *d, const char *address, size_t address_len) {
    if (!d || !address) { t_durable_error = "null argument"; return RIO_ERR_UNKNOWN; }
    std::lock_guard<std::mutex> g(d->mu);
    if (!run(d, d->st_clean, address, address_len, nullptr, 0, nullptr, 0, 1)) return fail(d, RIO_ERR_UPSTREAM, "clean_server");
    const rio_status st = rio_cuda_clean_server_str(d->h, address, address_len);
    return st == RIO_OK ? st : gpu_fail(d, st);
}

/* remove (sqlite.rs:114-126) */
rio_status rio_cuda_durable_remove(rio_durable *d, const char *type, size_t type_len, const char *id, size_t id_len) {
    return rio_cuda_durable_update(d, type, type_len, id, id_len, nullptr, 0);
}

/* Batched update: n ids with their addresses (NULL = remove), ONE transaction on the table, one batched upsert on the GPU.
 * Duplicate ids inside the batch resolve like the GPU's rule: the last occurrence wins (SQL statements run in array order). */
rio_status rio_cuda_durable_update_batch(rio_durable *d, const char *const *types, const char *const *ids, const char *const *addresses, size_t n) {
    if (!d || (n && (!types || !ids || !addresses))) { t_durable_error = "null argument"; return RIO_ERR_UNKNOWN; }
    if (!n) return RIO_OK;
    std::lock_guard<std::mutex> g(d->mu);
    std::vector<uint64_t> keys(n);
    std::vector<uint32_t> idx(n, RIO_NONE);
    std::unordered_map<std::string, uint32_t> interned;
    for (size_t k = 0; k < n; k++) {
        if (!types[k] || !ids[k]) { t_durable_error = "null id"; return RIO_ERR_UNKNOWN; }
        keys[k] = rio_cuda_object_key(types[k], strlen(types[k]), ids[k], strlen(ids[k]));
        if (addresses[k]) {
            auto it = interned.find(addresses[k]);
            if (it == interned.end()) {
                uint32_t j = RIO_NONE;
                const rio_status st = rio_cuda_node_intern(d->h, addresses[k], &j);
                if (st != RIO_OK) return gpu_fail(d, st);
                it = interned.emplace(addresses[k], j).first;
            }
            idx[k] = it->second;
        }
    }
    if (!exec(d, "BEGIN")) return fail(d, RIO_ERR_UPSTREAM, "begin");
    for (size_t k = 0; k < n; k++) {
        const bool ok = addresses[k] ? run(d, d->st_upsert, types[k], strlen(types[k]), ids[k], strlen(ids[k]), addresses[k], strlen(addresses[k]), 3)
                                     : run(d, d->st_remove, types[k], strlen(types[k]), ids[k], strlen(ids[k]), nullptr, 0, 2);
        if (!ok) { const rio_status st = fail(d, RIO_ERR_UPSTREAM, "batched update"); exec(d, "ROLLBACK"); return st; }
    }
    if (!exec(d, "COMMIT")) { const rio_status st = fail(d, RIO_ERR_UPSTREAM, "commit"); exec(d, "ROLLBACK"); return st; }
    const rio_status st = rio_cuda_upsert_batch(d->h, keys.data(), idx.data(), n);
    return st == RIO_OK ? st : gpu_fail(d, st);
}

/* Service::get_or_create_placement for a batch of ids (service.rs:193-254), written through: the GPU resolves the batch
 * (rio_cuda_place_batch), then ONE transaction deletes the rows of every inactive server the batch met (clean_server,
 * service.rs:233-237) and upserts the rows whose placement changed.  out_idx receives the node index of every id. */
rio_status rio_cuda_durable_place_batch(rio_durable *d, const char *const *types, const char *const *ids, size_t n, uint32_t policy, uint32_t self_idx,
                                        uint32_t *out_idx) {
    if (!d || (n && (!types || !ids || !out_idx))) { t_durable_error = "null argument"; return RIO_ERR_UNKNOWN; }
    if (!n) return RIO_OK;
    std::lock_guard<std::mutex> g(d->mu);
    std::vector<uint64_t> keys(n);
    std::vector<uint32_t> before(n);
    for (size_t k = 0; k < n; k++) {
        if (!types[k] || !ids[k]) { t_durable_error = "null id"; return RIO_ERR_UNKNOWN; }
        keys[k] = rio_cuda_object_key(types[k], strlen(types[k]), ids[k], strlen(ids[k]));
    }
    rio_status st = rio_cuda_lookup_batch(d->h, keys.data(), n, before.data());
    if (st == RIO_OK) st = rio_cuda_place_batch(d->h, keys.data(), n, policy, self_idx, out_idx);
    if (st != RIO_OK) return gpu_fail(d, st);
    // the servers place_batch cleaned: recorded on some id of the batch and not active (malformed records are dropped one by one)
    std::unordered_map<uint32_t, bool> cleaned;
    std::string abuf;
    // address of an interned node, whatever its length (two calls: the length, then the bytes)
    auto address_of = [&](uint32_t j) -> bool {
        size_t len = 0;
        if (rio_cuda_node_address(d->h, j, nullptr, 0, &len) != RIO_OK) return false;
        abuf.assign(len, '\0');
        return len == 0 || rio_cuda_node_address(d->h, j, &abuf[0], len, &len) == RIO_OK;
    };
    if (!exec(d, "BEGIN")) return fail(d, RIO_ERR_UPSTREAM, "begin");
    bool ok = true;
    for (size_t k = 0; k < n && ok; k++) {
        const uint32_t b = before[k];
        if (b == RIO_NONE || cleaned.count(b)) continue;
        int32_t active = 0, malformed = 0;
        uint32_t weight = 0;
        if (rio_cuda_node_state(d->h, b, &active, &weight, &malformed) != RIO_OK) { ok = false; break; }
        cleaned[b] = true;
        if (active || malformed) continue;
        if (!address_of(b)) { ok = false; break; }
        ok = run(d, d->st_clean, abuf.data(), abuf.size(), nullptr, 0, nullptr, 0, 1);
    }
    for (size_t k = 0; k < n && ok; k++) {
        // RIO_NONE: no live server to place on (the solver policies with an empty live set) -- nothing to record; a row on a
        // server that was cleaned is already gone
        if (before[k] == out_idx[k] || out_idx[k] == RIO_NONE) continue;
        if (!address_of(out_idx[k])) { ok = false; break; }
        ok = run(d, d->st_upsert, types[k], strlen(types[k]), ids[k], strlen(ids[k]), abuf.data(), abuf.size(), 3);
    }
    if (!ok || !exec(d, "COMMIT")) { const rio_status e = fail(d, RIO_ERR_UPSTREAM, "write-through of place_batch"); exec(d, "ROLLBACK"); return e; }
    return RIO_OK;
}

}  // extern "C"
