// durable_fuzz.cpp -- random operation sequences on the native durable provider (csrc/durable.cu over the engine, libsqlite3 dlopen'ed)
// against a shadow std::map with SqliteObjectPlacement's / LocalObjectPlacement's semantics (sqlite.rs:68-126, local.rs:12-68) and the
// placement policy of Service::get_or_create_placement (service.rs:193-254) for place_batch; every few hundred steps the provider is
// "crashed" (engine destroyed, table kept) and recovered, and must come back with exactly the shadow's rows.  Linked with engine.cu +
// the host-sim doubles under ASan + UBSan by tests/test_engine_host_sim.py; runs unchanged against the real engine.
// usage: durable_fuzz <scratch directory> [seed] [steps]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <set>
#include <string>
#include <vector>

#include "../../include/rio_cuda.h"

static int g_fail = 0;
#define EXPECT(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s  [%s | %s]\n", __FILE__, __LINE__, #c, rio_cuda_durable_last_error(), rio_cuda_last_error(nullptr)); if (++g_fail > 3) std::exit(1); } } while (0)

typedef std::pair<std::string, std::string> Id;

int main(int argc, char **argv) {
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    const uint64_t seed = argc > 2 ? strtoull(argv[2], nullptr, 10) : 1;
    const int steps = argc > 3 ? atoi(argv[3]) : 1500;
    const std::string path = dir + "/fuzz_" + std::to_string(seed) + ".sqlite3";
    std::remove(path.c_str());
    std::mt19937_64 rng(seed);
    auto R = [&](uint64_t n) { return rng() % n; };

    std::vector<std::string> servers;
    for (int j = 0; j < 6; j++) servers.push_back("0.0.0.0:" + std::to_string(5000 + j));
    std::set<std::string> active(servers.begin(), servers.end());
    const char *strays[] = {"9.9.9.9:1", "garbage", "old-host:77"};          // addresses update() may record without their being members
    std::map<Id, std::string> shadow;

    rio_placement *h = nullptr;
    rio_durable *d = nullptr;
    std::vector<uint32_t> nidx(servers.size());
    auto open_all = [&] {
        rio_config cfg{sizeof(rio_config), -1, 1024, 0, 0};
        EXPECT(rio_cuda_create(&cfg, &h) == RIO_OK);
        std::vector<const char *> p;
        for (auto &a : servers) p.push_back(a.c_str());
        EXPECT(rio_cuda_set_nodes(h, p.data(), nullptr, nullptr, (uint32_t)p.size(), 0, nidx.data()) == RIO_OK);
        for (size_t j = 0; j < servers.size(); j++) EXPECT(rio_cuda_node_set_active(h, nidx[j], active.count(servers[j]) ? 1 : 0) == RIO_OK);
        EXPECT(rio_cuda_durable_open(h, path.c_str(), &d) == RIO_OK);
    };
    auto lookup = [&](const Id &id) -> std::string {
        std::vector<char> buf(128);
        size_t n = 0;
        EXPECT(rio_cuda_durable_lookup(d, id.first.data(), id.first.size(), id.second.data(), id.second.size(), buf.data(), buf.size(), &n) == RIO_OK);
        return n == (size_t)-1 ? std::string("<none>") : std::string(buf.data(), n);
    };
    auto rand_id = [&] { return Id(R(3) ? "Obj" : "Other", std::to_string(R(120))); };
    auto rand_addr = [&]() -> std::string { return R(5) == 0 ? strays[R(3)] : servers[R(servers.size())]; };
    auto clean = [&](const std::string &a) { for (auto it = shadow.begin(); it != shadow.end();) it = it->second == a ? shadow.erase(it) : std::next(it); };
    auto malformed = [](const std::string &a) { const size_t c = a.find(':'); return c == std::string::npos || c == 0 || c + 1 >= a.size(); };

    open_all();
    for (int step = 0; step < steps && !g_fail; step++) {
        switch (R(9)) {
        case 0: case 1: {
            const Id id = rand_id();
            const std::string a = rand_addr();
            EXPECT(rio_cuda_durable_update(d, id.first.data(), id.first.size(), id.second.data(), id.second.size(), a.data(), a.size()) == RIO_OK);
            shadow[id] = a;
        } break;
        case 2: {
            const Id id = rand_id();
            if (R(2)) EXPECT(rio_cuda_durable_remove(d, id.first.data(), id.first.size(), id.second.data(), id.second.size()) == RIO_OK);
            else EXPECT(rio_cuda_durable_update(d, id.first.data(), id.first.size(), id.second.data(), id.second.size(), nullptr, 0) == RIO_OK);   // update(None)
            shadow.erase(id);
        } break;
        case 3: {
            const std::string a = R(6) == 0 ? "never-seen:1" : rand_addr();
            EXPECT(rio_cuda_durable_clean_server(d, a.data(), a.size()) == RIO_OK);
            clean(a);
        } break;
        case 4: {   // one transaction, array order, duplicates allowed, NULL = remove
            const size_t n = 1 + R(30);
            std::vector<Id> ids;
            std::vector<std::string> addrs;
            std::vector<char> is_null;
            for (size_t k = 0; k < n; k++) { ids.push_back(rand_id()); addrs.push_back(rand_addr()); is_null.push_back(R(5) == 0); }
            std::vector<const char *> t, i, a;
            for (size_t k = 0; k < n; k++) { t.push_back(ids[k].first.c_str()); i.push_back(ids[k].second.c_str()); a.push_back(is_null[k] ? nullptr : addrs[k].c_str()); }
            EXPECT(rio_cuda_durable_update_batch(d, t.data(), i.data(), a.data(), n) == RIO_OK);
            for (size_t k = 0; k < n; k++) { if (is_null[k]) shadow.erase(ids[k]); else shadow[ids[k]] = addrs[k]; }
        } break;
        case 5: case 6: {   // get_or_create_placement for a batch on one of the live servers (service.rs:193-254)
            if (active.empty()) break;
            auto it = active.begin();
            std::advance(it, R(active.size()));
            const std::string self = *it;
            uint32_t self_idx = 0;
            for (size_t j = 0; j < servers.size(); j++) if (servers[j] == self) self_idx = nidx[j];
            const size_t n = 1 + R(40);
            std::vector<Id> ids;
            for (size_t k = 0; k < n; k++) ids.push_back(rand_id());
            std::vector<const char *> t, i;
            for (auto &x : ids) { t.push_back(x.first.c_str()); i.push_back(x.second.c_str()); }
            std::vector<uint32_t> out(n);
            EXPECT(rio_cuda_durable_place_batch(d, t.data(), i.data(), n, RIO_PLACE_SELF, self_idx, out.data()) == RIO_OK);
            for (auto &id : ids) {
                auto f = shadow.find(id);
                if (f != shadow.end()) {
                    if (!malformed(f->second) && active.count(f->second)) continue;        // :226-231 keep
                    const std::string dead = f->second;
                    if (malformed(dead)) shadow.erase(f); else clean(dead);               // :213-222 / :233-237
                }
                shadow[id] = self;                                                         // :244-252
            }
            for (size_t k = 0; k < n; k++) EXPECT(lookup(ids[k]) == shadow[ids[k]]);
        } break;
        case 7: {   // a member dies or comes back
            const size_t j = R(servers.size());
            const bool on = R(2);
            if (!on && active.size() == 1 && active.count(servers[j])) break;
            EXPECT(rio_cuda_node_set_active(h, nidx[j], on) == RIO_OK);
            if (on) active.insert(servers[j]); else active.erase(servers[j]);
        } break;
        case 8: {   // crash + restart: the GPU directory is gone, the table is the source of truth
            if (R(6)) break;
            rio_cuda_durable_close(d);
            rio_cuda_destroy(h);
            open_all();
            uint64_t rows = 0;
            EXPECT(rio_cuda_durable_recover(d, &rows) == RIO_OK);
            EXPECT(rows == shadow.size());
            uint64_t placed = 0;
            EXPECT(rio_cuda_directory_len(h, &placed, nullptr) == RIO_OK && placed == shadow.size());
            for (auto &kv : shadow) EXPECT(lookup(kv.first) == kv.second);
        } break;
        }
        for (int q = 0; q < 4 && !g_fail; q++) {   // spot checks after every step
            const Id id = rand_id();
            auto f = shadow.find(id);
            const std::string want = f == shadow.end() ? "<none>" : f->second;
            const std::string got = lookup(id);
            if (got != want) { std::fprintf(stderr, "step %d: (%s,%s) is %s, expected %s\n", step, id.first.c_str(), id.second.c_str(), got.c_str(), want.c_str()); g_fail++; }
        }
    }
    rio_cuda_durable_close(d);
    rio_cuda_destroy(h);
    std::remove(path.c_str());
    if (g_fail) return 1;
    std::printf("durable fuzz: all passed (seed %llu, %d steps, %zu rows at the end)\n", (unsigned long long)seed, steps, shadow.size());
    return 0;
}
