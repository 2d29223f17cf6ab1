// client.cpp -- librio_client.so: client-side deterministic first hop (include/rio_client.h).
// Plain C++ (g++), no CUDA: a client resolves one id at a time on its own CPU.  Shares spec.cuh with the kernels.
#include "../../include/rio_client.h"

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "spec.cuh"
#include "trie_table.hpp"

using namespace rio;

struct rio_client_ring {
    std::vector<std::string> addr;
    struct Node { uint32_t s0, m, s2, invw; };   // invw == 0: not live
    std::vector<Node> nodes;
    std::vector<uint64_t> seed;
    std::vector<uint32_t> weight;
    // HRW2: the table blob -- thresholds of the binary trie over node positions in heap order, one leaf word per
    // bucket, member-keyed chain records -- built by the SAME function the servers use for the table their kernels walk
    // (trie_table.hpp), and walked here by the host form of the kernel's walk
    uint32_t policy = RIO_CLIENT_POLICY_HRW, bits = 12;
    TrieBlob trie;
    std::vector<ContestRec> level;
};

namespace {

uint32_t first_hop(const rio_client_ring &r, uint64_t key) {
    const ObjHash o = obj_hash(key);
    uint64_t best_sc = 0;
    uint32_t best_u = 0, best = kNone;
    for (uint32_t j = 0; j < (uint32_t)r.nodes.size(); j++) {          // flat loop over every node
        const rio_client_ring::Node &n = r.nodes[j];
        if (!n.invw) continue;
        const uint32_t u = pair_hash(o, n.s0, n.m, n.s2);
        const uint64_t sc = (uint64_t)elog(u) * n.invw;
        if (best == kNone || cand_better(sc, u, j, best_sc, best_u, best)) { best_sc = sc; best_u = u; best = j; }
    }
    return best;
}

void build_trie(rio_client_ring &r) {
    std::vector<TrieMember> members;
    for (uint32_t j = 0; j < (uint32_t)r.seed.size(); j++) if (r.weight[j]) members.push_back(TrieMember{r.seed[j], j, r.weight[j]});
    r.trie = build_trie_blob(members, r.bits);
    r.level = trie_level_constants(r.bits);
}

uint32_t first_hop_hrw2(const rio_client_ring &r, uint64_t key) {
    return trie_walk_host(r.trie.words.data(), r.trie.bits, r.level.data(), obj_hash(key));
}

uint32_t pick(const rio_client_ring &r, uint64_t key) { return r.policy == RIO_CLIENT_POLICY_HRW2 ? first_hop_hrw2(r, key) : first_hop(r, key); }

}  // namespace

extern "C" {

int32_t rio_client_ring_set_policy(rio_client_ring *ring, uint32_t policy, uint32_t trie_bits) {
    if (!ring || (policy != RIO_CLIENT_POLICY_HRW && policy != RIO_CLIENT_POLICY_HRW2) || trie_bits > 14) return RIO_CLIENT_ERR;
    try {
        ring->policy = policy;
        if (trie_bits) ring->bits = trie_bits;
        if (policy == RIO_CLIENT_POLICY_HRW2) build_trie(*ring);
    } catch (...) { return RIO_CLIENT_ERR; }
    return RIO_CLIENT_OK;
}

int32_t rio_client_ring_create(const char *const *addresses, const size_t *address_lens, const uint32_t *weights, uint32_t n, rio_client_ring **out) {
    if (!out || (n && (!addresses || !address_lens))) return RIO_CLIENT_ERR;
    rio_client_ring *r = new (std::nothrow) rio_client_ring();
    if (!r) return RIO_CLIENT_ERR;
    try {
        for (uint32_t j = 0; j < n; j++) {
            if (!addresses[j]) { delete r; return RIO_CLIENT_ERR; }
            r->addr.emplace_back(addresses[j], address_lens[j]);
            const uint64_t seed = mix64(fnv1a64(addresses[j], address_lens[j]));
            const uint64_t seed2 = mix64(seed ^ kSaltNode2);
            r->nodes.push_back({(uint32_t)seed, (uint32_t)(seed >> 32) | 1u, (uint32_t)seed2, inv_weight(weights ? weights[j] : 1u)});
            r->seed.push_back(seed);
            r->weight.push_back(weights ? weights[j] : 1u);
        }
    } catch (...) { delete r; return RIO_CLIENT_ERR; }
    *out = r;
    return RIO_CLIENT_OK;
}

void rio_client_ring_destroy(rio_client_ring *ring) { delete ring; }

uint32_t rio_client_ring_size(const rio_client_ring *ring) { return ring ? (uint32_t)ring->addr.size() : 0; }

int32_t rio_client_ring_address(const rio_client_ring *ring, uint32_t index, char *buf, size_t cap, size_t *out_len) {
    if (!ring || !out_len || index >= ring->addr.size()) return RIO_CLIENT_ERR;
    const std::string &a = ring->addr[index];
    *out_len = a.size();
    if (buf) for (size_t i = 0; i < a.size() && i < cap; i++) buf[i] = a[i];
    return RIO_CLIENT_OK;
}

uint64_t rio_client_object_key(const char *type, size_t type_len, const char *id, size_t id_len) {
    uint64_t h = fnv1a64(type, type_len);
    const char dot = '.';
    h = fnv1a64(&dot, 1, h);
    h = fnv1a64(id, id_len, h);
    return mix64(h);
}

int32_t rio_client_first_hop_key(const rio_client_ring *ring, uint64_t key, uint32_t *out_index) {
    if (!ring || !out_index) return RIO_CLIENT_ERR;
    *out_index = pick(*ring, key);
    return RIO_CLIENT_OK;
}

int32_t rio_client_first_hop(const rio_client_ring *ring, const char *type, size_t type_len, const char *id, size_t id_len, uint32_t *out_index) {
    if (!ring || !out_index || (!type && type_len) || (!id && id_len)) return RIO_CLIENT_ERR;
    *out_index = pick(*ring, rio_client_object_key(type, type_len, id, id_len));
    return RIO_CLIENT_OK;
}

int32_t rio_client_first_hop_batch(const rio_client_ring *ring, const uint64_t *keys, size_t n, uint32_t *out_index) {
    if (!ring || (n && (!keys || !out_index))) return RIO_CLIENT_ERR;
    for (size_t i = 0; i < n; i++) out_index[i] = pick(*ring, keys[i]);
    return RIO_CLIENT_OK;
}

}  // extern "C"
