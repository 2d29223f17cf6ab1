// bench_c1.cpp -- C1 (BASELINE.json configs[0]) through the C ABI with REAL threads (a Python harness serialises its callers on
// the interpreter lock): per-id lookups over 1 k ids / 4 nodes, direct (one GPU round trip per call behind the handle's mutex)
// and through the coalescing front end (rio_cuda_resolver_lookup) at 1 / 16 / 64 threads.  Prints one JSON object.
// build: g++ -O2 -std=c++17 -pthread tools/bench_c1.cpp -Iinclude -Lrio_rs_b200 -lrio_cuda -Wl,-rpath,$PWD/rio_rs_b200 -o tools/bench_c1
#include <chrono>
#include <cstdio>
#include <string>
#include <thread>
#include <vector>

#include "rio_cuda.h"

static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <class F>
static double run_threads(int T, int reps, int n, F &&fn) {
    std::vector<std::thread> th;
    const double t0 = now_us();
    for (int t = 0; t < T; t++) th.emplace_back([&, t] { for (int r = 0; r < reps; r++) for (int k = t; k < n; k += T) fn(k); });
    for (auto &x : th) x.join();
    return (now_us() - t0) / ((double)n * reps);
}

int main() {
    const int n = 1000, M = 4;
    rio_placement *h = nullptr;
    rio_config cfg{sizeof(rio_config), -1, 0, 0, 0};
    if (rio_cuda_create(&cfg, &h) != RIO_OK) { fprintf(stderr, "create: %s\n", rio_cuda_last_error(nullptr)); return 1; }
    std::vector<std::string> addr;
    std::vector<const char *> ap;
    for (int j = 0; j < M; j++) addr.push_back("10.0.0." + std::to_string(j) + ":5000");
    for (auto &a : addr) ap.push_back(a.c_str());
    rio_cuda_set_nodes(h, ap.data(), nullptr, nullptr, M, 0, nullptr);
    std::vector<uint64_t> keys(n);
    std::vector<uint32_t> idx(n), out(n);
    for (int i = 0; i < n; i++) { std::string id = std::to_string(i); keys[i] = rio_cuda_object_key("Obj", 3, id.data(), id.size()); idx[i] = i % M; }
    rio_cuda_upsert_batch(h, keys.data(), idx.data(), n);
    int bad = 0;
    auto direct = [&](int k) { uint32_t o = RIO_NONE; rio_cuda_lookup_batch(h, &keys[k], 1, &o); if (o != (uint32_t)(k % M)) bad++; };
    direct(0);
    const double d1 = run_threads(1, 2, n, direct), d16 = run_threads(16, 4, n, direct);
    rio_resolver *r = nullptr;
    rio_cuda_resolver_create(h, RIO_PLACE_SELF, 0, 256, 20, &r);
    auto coal = [&](int k) { uint32_t o = RIO_NONE; rio_cuda_resolver_lookup(r, keys[k], &o); if (o != (uint32_t)(k % M)) bad++; };
    coal(0);
    const double c1 = run_threads(1, 2, n, coal), c16 = run_threads(16, 8, n, coal), c64 = run_threads(64, 16, n, coal), c256 = run_threads(256, 32, n, coal);
    uint64_t calls = 0, batches = 0, largest = 0;
    rio_cuda_resolver_stats(r, &calls, &batches, &largest);
    rio_cuda_resolver_destroy(r);
    printf("{\"direct_us_per_lookup_1_thread\": %.3f, \"direct_us_per_lookup_16_threads\": %.3f, \"coalesced_us_per_lookup_1_thread\": %.3f, "
           "\"coalesced_us_per_lookup_16_threads\": %.3f, \"coalesced_us_per_lookup_64_threads\": %.3f, \"coalesced_us_per_lookup_256_threads\": %.3f, "
           "\"coalescing\": {\"calls\": %llu, \"batches\": %llu, \"largest_batch\": %llu}, \"wrong_answers\": %d}\n",
           d1, d16, c1, c16, c64, c256, (unsigned long long)calls, (unsigned long long)batches, (unsigned long long)largest, bad);
    rio_cuda_destroy(h);
    return bad ? 2 : 0;
}
