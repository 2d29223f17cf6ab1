// abi_threads.cpp -- "all exported functions are re-entrant and thread-safe per handle" (include/rio_cuda.h, SURVEY 8b: one tokio task
// per connection hits the same provider): T threads issue random C-ABI calls on ONE handle -- the trait's string calls, batched
// directory calls, place_batch / check_address_batch, membership changes, host-buffer assignments, a private resident set each, and one
// shared resolver -- against the engine's host code linked with the host-sim doubles under ThreadSanitizer
// (tests/test_engine_host_sim.py).  Checked: no data race, no crash, every status is one of the three codes, and every thread reads
// back its own writes to its own ids.
// usage: abi_threads [threads] [steps per thread]
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rio_cuda.h"
#include "../../include/rio_cuda_dev.h"

static std::atomic<int> g_fail{0};
#define EXPECT(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s  [%s]\n", __FILE__, __LINE__, #c, rio_cuda_last_error(nullptr)); g_fail++; } } while (0)
static bool status_ok(rio_status s) { return s == RIO_OK || s == RIO_ERR_UPSTREAM || s == RIO_ERR_UNKNOWN; }

int main(int argc, char **argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 8, steps = argc > 2 ? atoi(argv[2]) : 400;
    rio_placement *h = nullptr;
    rio_config cfg{sizeof(rio_config), -1, 1024, 0, 0};
    EXPECT(rio_cuda_create(&cfg, &h) == RIO_OK);
    std::vector<std::string> addrs;
    for (int j = 0; j < 12; j++) addrs.push_back("10.0.0." + std::to_string(j) + ":5000");
    std::vector<const char *> p;
    for (auto &a : addrs) p.push_back(a.c_str());
    std::vector<uint32_t> nidx(addrs.size());
    EXPECT(rio_cuda_set_nodes(h, p.data(), nullptr, nullptr, (uint32_t)p.size(), 0, nidx.data()) == RIO_OK);
    rio_resolver *r = nullptr;
    EXPECT(rio_cuda_resolver_create(h, RIO_PLACE_HRW2, 0, 64, 20, &r) == RIO_OK);

    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            std::mt19937_64 rng(1000 + t);
            auto R = [&](uint64_t n) { return rng() % n; };
            rio_objset *s = nullptr;
            EXPECT(rio_cuda_set_create(h, 4000, &s) == RIO_OK);
            EXPECT(rio_cuda_set_synth_keys(s, (uint64_t)t * 4000, 4000, 3) == RIO_OK);
            const std::string type = "T" + std::to_string(t);          // ids of this thread only; addresses nobody cleans
            const std::string mine = "own-" + std::to_string(t) + ":1";
            for (int k = 0; k < steps && !g_fail; k++) {
                const std::string id = std::to_string(R(50));
                switch (R(12)) {
                case 0: case 1: {
                    EXPECT(rio_cuda_update_str(h, type.data(), type.size(), id.data(), id.size(), mine.data(), mine.size()) == RIO_OK);
                    char buf[32];
                    size_t len = 0;
                    EXPECT(rio_cuda_lookup_str(h, type.data(), type.size(), id.data(), id.size(), buf, sizeof buf, &len) == RIO_OK);
                    EXPECT(len == mine.size() && !memcmp(buf, mine.data(), len));
                } break;
                case 2: {
                    EXPECT(rio_cuda_remove_str(h, type.data(), type.size(), id.data(), id.size()) == RIO_OK);
                    size_t len = 0;
                    EXPECT(rio_cuda_lookup_str(h, type.data(), type.size(), id.data(), id.size(), nullptr, 0, &len) == RIO_OK && len == (size_t)-1);
                } break;
                case 3: {
                    std::vector<uint64_t> keys(1 + R(300));
                    for (auto &x : keys) x = ((uint64_t)(t + 1) << 48) ^ rng();
                    std::vector<uint32_t> out(keys.size());
                    EXPECT(status_ok(rio_cuda_place_batch(h, keys.data(), keys.size(), (uint32_t)(1 + R(2)), 0, out.data())));
                    EXPECT(rio_cuda_lookup_batch(h, keys.data(), keys.size(), out.data()) == RIO_OK);
                } break;
                case 4: { const uint32_t j = nidx[R(nidx.size())]; EXPECT(rio_cuda_node_set_active(h, j, (int32_t)R(2)) == RIO_OK); } break;   // a flapping member
                case 5: { uint32_t idx = 0; EXPECT(rio_cuda_node_upsert(h, addrs[R(addrs.size())].c_str(), (uint32_t)(1 + R(8)), nullptr, 0, &idx) == RIO_OK); } break;
                case 6: {
                    std::vector<uint64_t> keys(1 + R(2000));
                    for (auto &x : keys) x = rng();
                    std::vector<uint32_t> out(keys.size());
                    uint32_t passes = 0;
                    EXPECT(status_ok(rio_cuda_assign_bounded_batch(h, keys.data(), keys.size(), 0, 5, 4, 4, out.data(), &passes)));
                } break;
                case 7: { uint32_t passes = 0; EXPECT(status_ok(rio_cuda_set_assign_bounded(s, 0, 101, 100, 4, &passes))); } break;
                case 8: {
                    EXPECT(status_ok(rio_cuda_set_assign_bounded_begin(s, 0, 5, 4, 4)));
                    uint32_t passes = 0;
                    EXPECT(status_ok(rio_cuda_set_assign_bounded_end(s, &passes)));
                } break;
                case 9: { uint32_t o = 0; EXPECT(status_ok(rio_cuda_resolver_resolve(r, ((uint64_t)(t + 1) << 52) ^ rng(), &o))); } break;
                case 10: {
                    std::vector<uint32_t> a(1 + R(40)), c(64);
                    for (auto &x : a) x = nidx[R(nidx.size())];
                    std::vector<uint8_t> v(a.size());
                    EXPECT(status_ok(rio_cuda_check_address_batch(h, a.data(), a.size(), nidx[0], v.data(), nullptr)));
                    EXPECT(status_ok(rio_cuda_set_solver(h, (uint32_t)(1 + R(2)), 0)));
                    EXPECT(status_ok(rio_cuda_load_counters(h, c.data(), (uint32_t)c.size())));
                } break;
                case 11: { uint64_t moved = 0; EXPECT(status_ok(rio_cuda_set_rebalance(s, (uint32_t)(1 + R(2)), nidx[R(nidx.size())], &moved))); } break;
                }
            }
            rio_cuda_set_destroy(s);
        });
    for (auto &x : th) x.join();
    rio_cuda_resolver_destroy(r);
    EXPECT(rio_cuda_sync(h) == RIO_OK);
    rio_cuda_destroy(h);
    if (g_fail) return 1;
    std::printf("abi threads: all passed (%d threads x %d steps)\n", T, steps);
    return 0;
}
