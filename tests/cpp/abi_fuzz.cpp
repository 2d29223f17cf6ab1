// abi_fuzz.cpp -- random sequences of C-ABI calls (include/rio_cuda.h) with ordinary, boundary and deliberately WRONG arguments against
// the engine's host code (csrc/engine.cu + resolver.cu, linked with the host-sim doubles under ASan + UBSan by
// tests/test_engine_host_sim.py).  The contract of the boundary: every call returns RIO_OK / RIO_ERR_UPSTREAM /
// RIO_ERR_UNKNOWN, no exception crosses, nothing is read or written out of bounds, an error leaves the handle usable -- and through all
// of it the directory keeps answering like LocalObjectPlacement (local.rs:12-68): a shadow std::map of the ids touched only by the
// string-level trait calls is compared after every step (entries a directory-wide re-placement may legitimately rewrite are
// forgotten, not guessed).
// usage: abi_fuzz [seed] [steps]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <optional>
#include <random>
#include <string>
#include <vector>

#include "../../include/rio_cuda.h"
#include "../../include/rio_cuda_dev.h"

static int g_fail = 0;
#define EXPECT(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s  [%s]\n", __FILE__, __LINE__, #c, rio_cuda_last_error(nullptr)); if (++g_fail > 5) std::exit(1); } } while (0)
static bool status_ok(rio_status s) { return s == RIO_OK || s == RIO_ERR_UPSTREAM || s == RIO_ERR_UNKNOWN; }

int main(int argc, char **argv) {
    const uint64_t seed = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
    const int steps = argc > 2 ? atoi(argv[2]) : 3000;
    std::mt19937_64 rng(seed);
    auto R = [&](uint64_t n) { return n ? rng() % n : 0; };

    rio_placement *h = nullptr;
    rio_config cfg{sizeof(rio_config), -1, 1024, 0, 0};
    EXPECT(rio_cuda_create(&cfg, &h) == RIO_OK);
    EXPECT(rio_cuda_create(nullptr, nullptr) != RIO_OK);

    std::vector<std::string> cluster;                       // addresses given to set_nodes / node_upsert
    std::vector<uint32_t> cluster_idx;
    std::map<std::string, std::optional<std::string>> shadow;   // "S.<n>" ids: value = address, nullopt = no placement; absent = unknown
    std::vector<rio_objset *> sets;
    rio_resolver *resolver = nullptr;
    uint32_t n_interned = 0;
    auto refresh_count = [&] { uint32_t t = 0, l = 0; EXPECT(rio_cuda_node_count(h, &t, &l) == RIO_OK); EXPECT(l <= t); n_interned = t; };
    auto any_idx = [&]() -> uint32_t { const uint64_t k = R(10); return k == 0 ? RIO_NONE : k == 1 ? n_interned + (uint32_t)R(3) : (uint32_t)R(n_interned ? n_interned : 1); };
    auto cluster_node = [&]() -> uint32_t { return cluster_idx.empty() ? 0u : cluster_idx[R(cluster_idx.size())]; };
    auto keys_of = [&](size_t n, uint64_t salt) { std::vector<uint64_t> k(n); for (size_t i = 0; i < n; i++) k[i] = (R(50) == 0) ? ~0ull - R(2) : (salt << 40) ^ rng(); return k; };

    for (int step = 0; step < steps && g_fail == 0; step++) {
        const uint64_t op = R(26);
        switch (op) {
        case 0: {   // set_nodes: sizes 0..300, weights incl. 0, sometimes features
            const uint32_t M = (uint32_t)(R(6) == 0 ? 0 : R(4) == 0 ? 100 + R(200) : 1 + R(12));
            const uint32_t K = R(4) == 0 ? (R(2) ? 16u : 4u) : 0u;
            cluster.clear();
            for (uint32_t j = 0; j < M; j++) cluster.push_back("10.0." + std::to_string(R(4)) + "." + std::to_string(j) + ":5000");
            std::vector<const char *> p;
            for (auto &a : cluster) p.push_back(a.c_str());
            std::vector<uint32_t> w(M), out(M ? M : 1);
            for (auto &x : w) x = (uint32_t)(R(8) == 0 ? 0 : R(16) == 0 ? 0xFFFFFFFFu - R(2) : 1 + R(16));   // dead, near 2^32, ordinary
            std::vector<float> f((size_t)M * (K ? K : 1));
            for (auto &x : f) x = (float)((double)R(2001) / 1000.0 - 1.0);
            EXPECT(rio_cuda_set_nodes(h, M ? p.data() : nullptr, R(3) ? w.data() : nullptr, K ? f.data() : nullptr, M, K, out.data()) == RIO_OK);
            cluster_idx.assign(out.begin(), out.begin() + M);
            EXPECT(rio_cuda_set_nodes(h, nullptr, nullptr, nullptr, 3, 0, nullptr) != RIO_OK);   // addrs NULL with M > 0
            refresh_count();
        } break;
        case 1: {   // node_upsert: well-formed, malformed, empty
            const char *forms[] = {"10.9.9.%d:5000", "garbage%d", ":%d", "host%d:", "a:b:c%d", ""};
            char buf[64];
            std::snprintf(buf, sizeof buf, forms[R(6)], (int)R(20));
            uint32_t idx = 0;
            EXPECT(rio_cuda_node_upsert(h, buf, (uint32_t)R(5), nullptr, 0, &idx) == RIO_OK);
            EXPECT(rio_cuda_node_upsert(h, nullptr, 1, nullptr, 0, &idx) != RIO_OK);
            cluster.push_back(buf); cluster_idx.push_back(idx);
            refresh_count();
        } break;
        case 2: { const uint32_t i = any_idx(); const rio_status st = rio_cuda_node_set_active(h, i, (int32_t)R(2)); EXPECT(status_ok(st) && (st == RIO_OK) == (i < n_interned)); } break;
        case 3: case 4: {   // update / remove through the trait's string calls
            const std::string id = std::to_string(R(400));
            if (R(4) == 0) {
                EXPECT(rio_cuda_remove_str(h, "S", 1, id.data(), id.size()) == RIO_OK);
                shadow["S." + id] = std::nullopt;
            } else {
                const std::string a = "sh-" + std::to_string(R(12)) + ":1";
                EXPECT(rio_cuda_update_str(h, "S", 1, id.data(), id.size(), a.data(), a.size()) == RIO_OK);
                shadow["S." + id] = a;
            }
            refresh_count();
        } break;
        case 5: {   // clean_server of a shadow address or an unknown one
            const std::string a = R(5) == 0 ? "never-seen:9" : "sh-" + std::to_string(R(12)) + ":1";
            EXPECT(rio_cuda_clean_server_str(h, a.data(), a.size()) == RIO_OK);
            for (auto &kv : shadow) if (kv.second && *kv.second == a) kv.second = std::nullopt;
        } break;
        case 6: {   // batched directory calls on foreign keys, all sizes incl. 0 and NULL buffers
            const size_t n = R(5) == 0 ? 0 : R(3) == 0 ? 3000 + R(3000) : 1 + R(40);
            std::vector<uint64_t> k = keys_of(n, 7);
            std::vector<uint32_t> idx(n ? n : 1), out(n ? n : 1);
            for (auto &x : idx) x = any_idx();
            EXPECT(rio_cuda_upsert_batch(h, k.data(), idx.data(), n) == RIO_OK);
            EXPECT(rio_cuda_lookup_batch(h, k.data(), n, out.data()) == RIO_OK);
            if (n) EXPECT(rio_cuda_lookup_batch(h, nullptr, n, out.data()) != RIO_OK);
            EXPECT(rio_cuda_remove_batch(h, k.data(), n / 2) == RIO_OK);
        } break;
        case 7: {   // place_batch: every policy incl. an invalid one, self_idx valid or not
            const size_t n = R(4) == 0 ? 0 : 1 + R(200);
            std::vector<uint64_t> k = keys_of(n, 11);
            std::vector<uint32_t> out(n ? n : 1);
            const uint32_t policy = (uint32_t)R(4), self = R(6) == 0 ? any_idx() : cluster_node();
            const rio_status st = rio_cuda_place_batch(h, k.data(), n, policy, self, out.data());
            EXPECT(status_ok(st));
            if (st == RIO_OK && n && policy == RIO_PLACE_SELF) for (size_t i = 0; i < n; i++) EXPECT(out[i] != RIO_NONE);
            if (n && policy == 3) EXPECT(st != RIO_OK);
            // the two reserved-looking keys are shared with case 6, which may have recorded them on ANY interned address: meeting a
            // non-active one makes place_batch clean that server (service.rs:233-237), shadow ids included -- forget, do not guess
            for (uint64_t key : k) if (key >= ~0ull - 1) { shadow.clear(); break; }
        } break;
        case 8: {   // check_address_batch
            const size_t n = R(4) == 0 ? 0 : 1 + R(100);
            std::vector<uint32_t> a(n ? n : 1);
            for (auto &x : a) x = any_idx();
            std::vector<uint8_t> v(n ? n : 1, 9);
            uint64_t cleaned = 0;
            const rio_status st = rio_cuda_check_address_batch(h, a.data(), n, cluster_node(), v.data(), &cleaned);
            EXPECT(status_ok(st));
            if (st == RIO_OK) for (size_t i = 0; i < n; i++) EXPECT(v[i] <= 3);
            if (st == RIO_OK && cleaned) shadow.clear();          // a non-active shadow address may have been cleaned: forget, do not guess
        } break;
        case 9: {   // assign_batch / assign_bounded_batch from host buffers
            const size_t n = R(4) == 0 ? 0 : 1 + R(3000);
            std::vector<uint64_t> k = keys_of(n, 13);
            std::vector<uint32_t> out(n ? n : 1);
            EXPECT(status_ok(rio_cuda_assign_batch(h, k.data(), nullptr, n, out.data())));
            if (R(3) == 0) {   // object features without keys (K of the handle or not)
                std::vector<float> f((n ? n : 1) * 16, 0.25f);
                EXPECT(status_ok(rio_cuda_assign_batch(h, nullptr, f.data(), n, out.data())));
            }
            if (n) EXPECT(rio_cuda_assign_batch(h, nullptr, nullptr, n, out.data()) != RIO_OK);
            uint32_t passes = 0;
            const rio_status st = rio_cuda_assign_bounded_batch(h, k.data(), n, R(3) ? 0 : n * 3, (uint32_t)R(8), (uint32_t)R(5), (uint32_t)R(6), out.data(), &passes);
            EXPECT(status_ok(st));
            if (st == RIO_OK && n) EXPECT(passes >= 1 && passes <= 5);
        } break;
        case 10: { const rio_status st = rio_cuda_set_solver(h, (uint32_t)R(4), (uint32_t)R(17)); EXPECT(status_ok(st)); } break;
        case 11: {   // directory-wide rebalance: preconditions are checked, not assumed
            uint64_t moved = 0;
            const rio_status st = rio_cuda_rebalance(h, (uint32_t)R(4), R(4) == 0 ? any_idx() : cluster_node(), &moved);
            EXPECT(status_ok(st));
            if (st == RIO_OK && moved) shadow.clear();            // may have re-placed anything: forget, do not guess
        } break;
        case 12: {   // new object set
            if (sets.size() < 4) {
                rio_objset *s = nullptr;
                const uint64_t cap = R(6) == 0 ? 0 : 1 + R(5000);
                const rio_status st = rio_cuda_set_create(h, cap, &s);
                EXPECT(status_ok(st) && (st == RIO_OK) == (cap > 0));
                if (s) { sets.push_back(s); EXPECT(rio_cuda_set_synth_keys(s, R(1000), cap, R(9)) == RIO_OK); }
            }
        } break;
        case 13: case 14: {   // set operations in any order, incl. the wrong one
            if (sets.empty()) break;
            rio_objset *s = sets[R(sets.size())];
            uint64_t n = 0;
            EXPECT(rio_cuda_set_size(s, &n) == RIO_OK);
            uint32_t passes = 0;
            uint64_t moved = 0;
            switch (R(10)) {
            case 0: EXPECT(status_ok(rio_cuda_set_assign(s, (uint32_t)R(5) == 0))); break;
            case 1: EXPECT(status_ok(rio_cuda_set_assign_bounded(s, R(3) ? 0 : n, (uint32_t)(1 + R(200)), (uint32_t)(1 + R(100)), (uint32_t)(1 + R(5)), &passes))); break;
            case 2: EXPECT(status_ok(rio_cuda_set_assign_bounded_begin(s, 0, 5, 4, (uint32_t)(1 + R(4))))); break;
            case 3: EXPECT(status_ok(rio_cuda_set_assign_bounded_end(s, &passes))); break;                       // with or without a begin before it
            case 4: EXPECT(status_ok(rio_cuda_set_rebalance(s, (uint32_t)(1 + R(2)), cluster_node(), &moved))); break;
            case 5: { std::vector<uint32_t> c(n_interned + 4); EXPECT(status_ok(rio_cuda_set_counters(s, c.data(), R(4) ? (uint32_t)c.size() : 0))); } break;
            case 6: { std::vector<uint32_t> o(n + 1); std::vector<uint64_t> kk(n + 1); EXPECT(rio_cuda_set_read(s, 0, n, kk.data(), o.data()) == RIO_OK); EXPECT(rio_cuda_set_read(s, 1, n, nullptr, o.data()) != RIO_OK); } break;
            case 7: { const rio_status st = rio_cuda_set_commit(s); EXPECT(status_ok(st)); } break;
            case 9: {   // affinity features of the right and of the wrong width
                const uint32_t K = R(3) == 0 ? 4u : 16u;
                std::vector<float> f((size_t)(n ? n : 1) * K);
                for (auto &x : f) x = (float)((double)R(2001) / 1000.0 - 1.0);
                EXPECT(status_ok(rio_cuda_set_load_feats(s, f.data(), K)));
                EXPECT(rio_cuda_set_load_feats(s, nullptr, K) != RIO_OK);
                EXPECT(status_ok(rio_cuda_set_assign(s, 1)));
            } break;
            case 8: { std::vector<uint64_t> kk = keys_of(n / 2 + 1, 17); EXPECT(status_ok(rio_cuda_set_load_keys(s, kk.data(), kk.size()))); EXPECT(rio_cuda_set_load_keys(s, kk.data(), n + 100000) != RIO_OK); } break;
            }
        } break;
        case 15: {   // drop a set (possibly with a bounded call still open on it)
            if (!sets.empty() && R(3) == 0) { const size_t i = R(sets.size()); rio_cuda_set_destroy(sets[i]); sets.erase(sets.begin() + i); }
        } break;
        case 16: {   // per-id calls through the coalescing front end
            if (!resolver) { EXPECT(status_ok(rio_cuda_resolver_create(h, (uint32_t)R(3), cluster_node(), 64, 20, &resolver))); break; }
            uint32_t idx = 0;
            EXPECT(status_ok(rio_cuda_resolver_resolve(resolver, (23ull << 40) ^ rng(), &idx)));
            EXPECT(status_ok(rio_cuda_resolver_lookup(resolver, rng(), &idx)));
            if (R(10) == 0) { rio_cuda_resolver_destroy(resolver); resolver = nullptr; }
        } break;
        case 17: {   // small out-buffers and NULL optionals
            uint32_t c[2];
            EXPECT(status_ok(rio_cuda_load_counters(h, c, 2)));
            uint64_t placed = 0, slots = 0;
            EXPECT(rio_cuda_directory_len(h, &placed, nullptr) == RIO_OK && rio_cuda_directory_len(h, nullptr, &slots) == RIO_OK && placed <= slots);
            EXPECT(rio_cuda_directory_reserve(h, R(20000)) == RIO_OK);
            char name[8];
            EXPECT(rio_cuda_device_info(h, nullptr, nullptr, nullptr, name, sizeof name) == RIO_OK && strlen(name) < sizeof name);
            size_t len = 0;
            EXPECT(status_ok(rio_cuda_node_address(h, any_idx(), nullptr, 0, &len)));
            char one[1];
            EXPECT(status_ok(rio_cuda_node_address(h, cluster_node(), one, 1, &len)));
        } break;
        case 18: {   // hash_ids == object_key
            const size_t n = 1 + R(50);
            std::string packed;
            std::vector<uint64_t> off{0}, want, got(n);
            for (size_t i = 0; i < n; i++) { const std::string t(1 + R(12), (char)('a' + R(26))), id = std::to_string(rng() % 100000); packed += t + "." + id; off.push_back(packed.size()); want.push_back(rio_cuda_object_key(t.data(), t.size(), id.data(), id.size())); }
            packed.append(16, '\0');
            EXPECT(rio_cuda_hash_ids(h, packed.data(), off.data(), n, got.data()) == RIO_OK);
            EXPECT(got == want);
        } break;
        case 19: {   // peer-memory window of a 1-rank world, exported twice
            uint8_t hd[RIO_IPC_HANDLE_BYTES];
            EXPECT(status_ok(rio_cuda_comm_ipc_export(h, 1, (uint32_t)(1 + R(2000)), hd)));
            EXPECT(status_ok(rio_cuda_comm_ipc_attach(h, 0, 1, hd)));
            EXPECT(rio_cuda_comm_ipc_export(h, 17, 8, hd) != RIO_OK);
            EXPECT(rio_cuda_comm_ipc_attach(h, 3, 2, hd) != RIO_OK);
            std::vector<uint32_t> c(n_interned + 1, 1);
            EXPECT(status_ok(rio_cuda_comm_sum_counters(h, c.data(), n_interned)));
        } break;
        case 20: {   // dev hooks: duplicated seeds (exact ties), one class per node
            if (n_interned) EXPECT(status_ok(rio_dev_set_node_seed(h, any_idx(), R(3) ? rng() : 12345)));
            EXPECT(rio_dev_set_table_options(h, (uint32_t)R(2)) == RIO_OK);
        } break;
        case 21: {   // lookups with a tiny buffer: the length is reported, nothing past `cap` is written
            const std::string id = std::to_string(R(400));
            char guard[8] = {'x', 'x', 'x', 'x', 'x', 'x', 'x', 'x'};
            size_t len = 0;
            EXPECT(rio_cuda_lookup_str(h, "S", 1, id.data(), id.size(), guard, 2, &len) == RIO_OK);
            EXPECT(guard[2] == 'x' && guard[7] == 'x');
            EXPECT(rio_cuda_lookup_str(h, nullptr, 0, id.data(), id.size(), guard, 2, &len) != RIO_OK);
        } break;
        default: break;   // 22..25: only the shadow check below
        }
        // LocalObjectPlacement's answers for the ids only the trait's string calls touch
        if (step % 7 == 0 || op == 3 || op == 4 || op == 5)
            for (auto &kv : shadow) {
                char buf[64];
                size_t len = 0;
                const std::string id = kv.first.substr(2);
                EXPECT(rio_cuda_lookup_str(h, "S", 1, id.data(), id.size(), buf, sizeof buf, &len) == RIO_OK);
                if (!kv.second) EXPECT(len == (size_t)-1);
                else EXPECT(len == kv.second->size() && !memcmp(buf, kv.second->data(), len));
                if (g_fail) { std::fprintf(stderr, "  step %d op %llu id %s\n", step, (unsigned long long)op, kv.first.c_str()); break; }
            }
    }
    if (resolver) rio_cuda_resolver_destroy(resolver);
    for (auto *s : sets) rio_cuda_set_destroy(s);
    EXPECT(rio_cuda_sync(h) == RIO_OK);
    rio_cuda_destroy(h);
    rio_cuda_destroy(nullptr);
    if (g_fail) return 1;
    std::printf("abi fuzz: all passed (seed %llu, %d steps)\n", (unsigned long long)seed, steps);
    return 0;
}
