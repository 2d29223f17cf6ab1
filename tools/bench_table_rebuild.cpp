// bench_table_rebuild.cpp -- host time of ONE membership event (node_set_active -> table rebuild: class-sorted records, HRW2 blob, policy
// state, staging copy) at M = 1024, trie_bits 12, measured on the host-sim build of the engine (tests/cpp/hostsim: no GPU, "uploads" are
// memcpy) -- i.e. the CPU work a rio_cuda_rebalance / set_rebalance pays before its kernel can start (C5: 8 such events).
//   g++ -std=c++17 -O3 -Itests/cpp/hostsim -I. -x c++ rio_rs_b200/csrc/engine.cu rio_rs_b200/csrc/resolver.cu rio_rs_b200/csrc/durable.cu \
//       tests/cpp/hostsim/launchers.cpp tools/bench_table_rebuild.cpp -o /tmp/bench_table_rebuild -ldl -lpthread && /tmp/bench_table_rebuild
#include <chrono>
#include <cstdio>
#include <string>
#include <vector>
#include "include/rio_cuda.h"
int main() {
    rio_placement *h = nullptr; rio_config cfg{sizeof(rio_config), -1, 1024, 0, 0};
    rio_cuda_create(&cfg, &h);
    const uint32_t M = 1024;
    std::vector<std::string> a; for (uint32_t j = 0; j < M; j++) a.push_back("10.0." + std::to_string(j >> 8) + "." + std::to_string(j & 255) + ":5000");
    std::vector<const char*> p; for (auto &x : a) p.push_back(x.c_str());
    std::vector<uint32_t> w(M), idx(M); for (uint32_t j = 0; j < M; j++) w[j] = 1 + (j * 7) % 16;
    rio_cuda_set_nodes(h, p.data(), w.data(), nullptr, M, 0, idx.data());
    rio_cuda_set_solver(h, RIO_SOLVER_HRW2, 12);
    uint64_t key = 1; uint32_t out = 0;
    rio_cuda_assign_batch(h, &key, nullptr, 1, &out);
    for (int rep = 0; rep < 3; rep++) {
        auto t0 = std::chrono::steady_clock::now();
        const int K = 200;
        for (int k = 0; k < K; k++) { rio_cuda_node_set_active(h, idx[17 + k % 50], k & 1); rio_cuda_assign_batch(h, &key, nullptr, 1, &out); }
        double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / K;
        printf("membership event -> table rebuild (M = %u, trie_bits 12) + 1-object assign: %.1f us per event\n", M, us);
    }
    rio_cuda_destroy(h);
}
