// resolver_stress.cpp -- the micro-batching front end (csrc/resolver.cu, rio_cuda_resolver_*) under concurrent per-id callers.
// Runs against any implementation of the C ABI: the real engine on the GPU box, or the in-memory test double
// (tests/cpp/model_backend.cpp) on a box without a GPU -- there also under ThreadSanitizer (tests/test_host_layers_cpu.py).
// What is checked is the QUEUE: every caller gets the answer that belongs to ITS id (no crossed slots), updates are visible to the
// caller's own later lookups (a caller blocks until its batch is applied), errors of a batch reach every waiter, calls are
// coalesced, and create / destroy with callers in flight neither hangs nor races.
// usage: resolver_stress [threads] [ids per thread]
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rio_cuda.h"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s [%s]\n", __FILE__, __LINE__, #c, rio_cuda_resolver_last_error()); std::exit(1); } } while (0)

static std::string lookup(rio_resolver *r, const std::string &id) {
    char buf[64];
    size_t n = 0;
    CHECK(rio_cuda_resolver_lookup_str(r, "Obj", 3, id.data(), id.size(), buf, sizeof buf, &n) == RIO_OK);
    return n == (size_t)-1 ? std::string("<none>") : std::string(buf, n);
}

int main(int argc, char **argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 16, N = argc > 2 ? atoi(argv[2]) : 300;
    rio_placement *h = nullptr;
    rio_config cfg{sizeof(rio_config), -1, 0, 0, 0};
    CHECK(rio_cuda_create(&cfg, &h) == RIO_OK);
    const char *nodes[4] = {"0.0.0.0:5000", "0.0.0.0:5001", "0.0.0.0:5002", "0.0.0.0:5003"};
    uint32_t nidx[4];
    CHECK(rio_cuda_set_nodes(h, nodes, nullptr, nullptr, 4, 0, nidx) == RIO_OK);
    rio_resolver *r = nullptr;
    CHECK(rio_cuda_resolver_create(h, RIO_PLACE_SELF, nidx[0], 256, 50, &r) == RIO_OK);

    std::atomic<int> wrong{0};
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            for (int k = 0; k < N; k++) {
                const std::string id = std::to_string(t) + "-" + std::to_string(k);
                const char *mine = nodes[1 + (t + k) % 3];
                if (lookup(r, id) != "<none>") wrong++;                                                    // nobody placed it yet
                CHECK(rio_cuda_resolver_update_str(r, "Obj", 3, id.data(), id.size(), mine, strlen(mine)) == RIO_OK);
                if (lookup(r, id) != mine) wrong++;                                                        // my own write, visible to me
                if (k % 3 == 0) {
                    CHECK(rio_cuda_resolver_update_str(r, "Obj", 3, id.data(), id.size(), nullptr, 0) == RIO_OK);   // update(None) == remove
                    if (lookup(r, id) != "<none>") wrong++;
                    char buf[64];
                    size_t n = 0;                                                                          // unplaced -> the serving node claims it
                    CHECK(rio_cuda_resolver_resolve_str(r, "Obj", 3, id.data(), id.size(), buf, sizeof buf, &n) == RIO_OK);
                    if (std::string(buf, n) != nodes[0]) wrong++;
                } else {
                    char buf[64];
                    size_t n = 0;                                                                          // placed on a live node -> kept
                    CHECK(rio_cuda_resolver_resolve_str(r, "Obj", 3, id.data(), id.size(), buf, sizeof buf, &n) == RIO_OK);
                    if (std::string(buf, n) != mine) wrong++;
                }
            }
        });
    for (auto &x : th) x.join();
    CHECK(wrong.load() == 0);
    uint64_t calls = 0, batches = 0, largest = 0;
    CHECK(rio_cuda_resolver_stats(r, &calls, &batches, &largest) == RIO_OK);
    CHECK(calls >= (uint64_t)T * N * 4 && batches >= 1 && batches <= calls && largest >= 1);
    if (T >= 8) CHECK(batches < calls);   // with many callers at least some calls share a batch

    // an engine error inside a batch reaches the caller that waited for it, with the message
    rio_resolver *bad = nullptr;
    CHECK(rio_cuda_resolver_create(h, RIO_PLACE_SELF, 0xFFFFFF00u, 16, 10, &bad) == RIO_OK);   // self_idx is not a node: place_batch will refuse
    uint32_t idx = 0;
    CHECK(rio_cuda_resolver_resolve(bad, 12345, &idx) != RIO_OK);
    CHECK(strlen(rio_cuda_resolver_last_error()) > 0);
    rio_cuda_resolver_destroy(bad);

    // callers hammering the queue right up to the shutdown: every call completes, the worker drains and joins, nothing hangs
    std::atomic<bool> stop{false};
    std::vector<std::thread> late;
    for (int t = 0; t < 4; t++)
        late.emplace_back([&] {
            uint32_t o = 0;
            uint64_t k = 0;
            while (!stop.load()) { if (rio_cuda_resolver_lookup(r, ++k, &o) != RIO_OK) break; }
        });
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
    stop.store(true);
    for (auto &x : late) x.join();
    rio_cuda_resolver_destroy(r);
    rio_cuda_destroy(h);
    std::printf("resolver: all passed (%d threads x %d ids: %llu calls in %llu batches, largest %llu)\n", T, N, (unsigned long long)calls, (unsigned long long)batches,
                (unsigned long long)largest);
    return 0;
}
