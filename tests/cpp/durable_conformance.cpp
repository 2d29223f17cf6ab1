// durable_conformance.cpp -- the reference's SqliteObjectPlacement tests restated against the native durable write-through
// (include/rio_cuda.h rio_cuda_durable_*): the GPU directory answers, the reference's SQLite table is the source of truth.
// Each function cites the reference test it restates; exit code 0 = all passed.  Built and run by tests/test_gpu_cpp.py.
// usage: durable_conformance <scratch directory>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <optional>
#include <string>
#include <vector>

#include "../../include/rio_cuda.h"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s  [%s | %s]\n", __FILE__, __LINE__, #c, rio_cuda_durable_last_error(), rio_cuda_last_error(nullptr)); std::exit(1); } } while (0)

static rio_placement *engine() {
    rio_placement *h = nullptr;
    rio_config cfg{sizeof(rio_config), -1, 0, 0, 0};
    CHECK(rio_cuda_create(&cfg, &h) == RIO_OK);
    return h;
}
static std::optional<std::string> lookup(rio_durable *d, const char *t, const char *i) {
    char buf[256];
    size_t n = 0;
    CHECK(rio_cuda_durable_lookup(d, t, strlen(t), i, strlen(i), buf, sizeof buf, &n) == RIO_OK);
    if (n == (size_t)-1) return std::nullopt;
    return std::string(buf, n);
}
static void update(rio_durable *d, const char *t, const char *i, const char *a) { CHECK(rio_cuda_durable_update(d, t, strlen(t), i, strlen(i), a, a ? strlen(a) : 0) == RIO_OK); }

// rio-rs/src/object_placement/sqlite.rs:149-193 (test_sanity), statement for statement
static void test_sanity(const std::string &dir) {
    rio_placement *h = engine();
    rio_durable *d = nullptr;
    CHECK(rio_cuda_durable_open(h, (dir + "/sanity.sqlite3").c_str(), &d) == RIO_OK);   // prepare()
    CHECK(!lookup(d, "Test", "1").has_value());
    update(d, "Test", "1", "0.0.0.0:5000");
    CHECK(lookup(d, "Test", "1") == std::string("0.0.0.0:5000"));
    update(d, "Test", "1", "0.0.0.0:5001");                                              // overwrite
    CHECK(lookup(d, "Test", "1") == std::string("0.0.0.0:5001"));
    CHECK(rio_cuda_durable_clean_server(d, "0.0.0.0:5001", 12) == RIO_OK);
    CHECK(!lookup(d, "Test", "1").has_value());
    rio_cuda_durable_close(d);
    rio_cuda_destroy(h);
}

// rio-rs/tests/object_placement_backend.rs:11-34 through the durable provider
static void no_placement_and_save_and_load(const std::string &dir) {
    rio_placement *h = engine();
    rio_durable *d = nullptr;
    CHECK(rio_cuda_durable_open(h, (dir + "/backend.sqlite3").c_str(), &d) == RIO_OK);
    CHECK(!lookup(d, "obj", "1").has_value());
    update(d, "obj", "1", "0.0.0.0:8888");
    CHECK(lookup(d, "obj", "1") == std::string("0.0.0.0:8888"));
    CHECK(rio_cuda_durable_clean_server(d, "0.0.0.0:8888", 12) == RIO_OK);
    CHECK(!lookup(d, "obj", "1").has_value());
    rio_cuda_durable_close(d);
    rio_cuda_destroy(h);
}

// what the reference gets from SQLite for free: a restarted server finds every placement again
static void restart_recovers_the_directory(const std::string &dir) {
    const std::string path = dir + "/restart.sqlite3";
    const int n = 5000;
    std::vector<std::string> ids, addrs;
    for (int i = 0; i < n; i++) { ids.push_back(std::to_string(i)); addrs.push_back("10.0.0." + std::to_string(i % 7) + ":5000"); }
    {
        rio_placement *h = engine();
        rio_durable *d = nullptr;
        CHECK(rio_cuda_durable_open(h, path.c_str(), &d) == RIO_OK);
        std::vector<const char *> t(n, "Obj"), i, a;
        for (int k = 0; k < n; k++) { i.push_back(ids[k].c_str()); a.push_back(addrs[k].c_str()); }
        CHECK(rio_cuda_durable_update_batch(d, t.data(), i.data(), a.data(), n) == RIO_OK);                 // one transaction
        update(d, "Obj", "7", "10.0.0.6:5000");
        CHECK(rio_cuda_durable_remove(d, "Obj", 3, "8", 1) == RIO_OK);
        CHECK(rio_cuda_durable_clean_server(d, "10.0.0.2:5000", 13) == RIO_OK);
        // duplicates inside one batch: the last occurrence wins on both sides
        const char *dt[3] = {"Dup", "Dup", "Dup"}, *di[3] = {"x", "x", "y"}, *da[3] = {"10.0.0.1:5000", nullptr, "10.0.0.1:5000"};
        CHECK(rio_cuda_durable_update_batch(d, dt, di, da, 3) == RIO_OK);
        CHECK(!lookup(d, "Dup", "x").has_value() && lookup(d, "Dup", "y").has_value());
        rio_cuda_durable_close(d);
        rio_cuda_destroy(h);   // "crash": the GPU directory is gone
    }
    rio_placement *h = engine();
    rio_durable *d = nullptr;
    CHECK(rio_cuda_durable_open(h, path.c_str(), &d) == RIO_OK);
    CHECK(!lookup(d, "Obj", "1").has_value());                      // empty cache before recovery
    uint64_t rows = 0;
    CHECK(rio_cuda_durable_recover(d, &rows) == RIO_OK);
    uint64_t want_rows = 1;                                          // ("Dup","y")
    for (int k = 0; k < n; k++) {
        std::optional<std::string> want = addrs[k];
        if (k == 7) want = "10.0.0.6:5000";
        if (k == 8 || (want && *want == "10.0.0.2:5000")) want.reset();
        want_rows += want.has_value();
        if (k % 11 == 0 || k < 20) CHECK(lookup(d, "Obj", ids[k].c_str()) == want);
    }
    CHECK(rows == want_rows);
    uint64_t placed = 0, slots = 0;
    CHECK(rio_cuda_directory_len(h, &placed, &slots) == RIO_OK && placed == rows);
    CHECK(!lookup(d, "Dup", "x").has_value() && lookup(d, "Dup", "y") == std::string("10.0.0.1:5000"));
    rio_cuda_durable_close(d);
    rio_cuda_destroy(h);
}

// get_or_create_placement for a batch (service.rs:193-254), written through: claims, keeps, and the clean_server of a dead owner
static void place_batch_is_written_through(const std::string &dir) {
    const std::string path = dir + "/place.sqlite3";
    const char *nodes[3] = {"0.0.0.0:5000", "0.0.0.0:5001", "0.0.0.0:5002"};
    const int n = 300;
    std::vector<std::string> ids;
    for (int i = 0; i < n; i++) ids.push_back(std::to_string(i));
    std::vector<const char *> t(n, "MockService"), i;
    for (auto &s : ids) i.push_back(s.c_str());
    std::vector<uint32_t> out(n);
    {
        rio_placement *h = engine();
        uint32_t nidx[3];
        CHECK(rio_cuda_set_nodes(h, nodes, nullptr, nullptr, 3, 0, nidx) == RIO_OK);
        rio_durable *d = nullptr;
        CHECK(rio_cuda_durable_open(h, path.c_str(), &d) == RIO_OK);
        CHECK(rio_cuda_durable_place_batch(d, t.data(), i.data(), 200, RIO_PLACE_SELF, nidx[0], out.data()) == RIO_OK);       // 0..199 claimed by server 0
        CHECK(rio_cuda_durable_place_batch(d, t.data() + 100, i.data() + 100, 200, RIO_PLACE_SELF, nidx[1], out.data()) == RIO_OK);   // 100..199 kept, 200..299 -> server 1
        CHECK(out[0] == nidx[0] && out[150] == nidx[1]);
        CHECK(rio_cuda_node_set_active(h, nidx[0], 0) == RIO_OK);                                                               // server 0 dies
        CHECK(rio_cuda_durable_place_batch(d, t.data(), i.data(), 50, RIO_PLACE_SELF, nidx[2], out.data()) == RIO_OK);         // 0..49 re-placed on server 2; clean_server(server 0)
        for (int k = 0; k < 50; k++) CHECK(out[k] == nidx[2]);
        rio_cuda_durable_close(d);
        rio_cuda_destroy(h);
    }
    rio_placement *h = engine();
    rio_durable *d = nullptr;
    CHECK(rio_cuda_durable_open(h, path.c_str(), &d) == RIO_OK);
    uint64_t rows = 0;
    CHECK(rio_cuda_durable_recover(d, &rows) == RIO_OK);
    CHECK(rows == 50 + 100);                                             // 0..49 on server 2, 200..299 on server 1; 50..199 went with server 0
    CHECK(lookup(d, "MockService", "10") == std::string("0.0.0.0:5002"));
    CHECK(!lookup(d, "MockService", "120").has_value());
    CHECK(lookup(d, "MockService", "250") == std::string("0.0.0.0:5001"));
    rio_cuda_durable_close(d);
    rio_cuda_destroy(h);
}

int main(int argc, char **argv) {
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    test_sanity(dir);
    no_placement_and_save_and_load(dir);
    restart_recovers_the_directory(dir);
    place_batch_is_written_through(dir);
    std::printf("durable: all passed\n");
    return 0;
}
