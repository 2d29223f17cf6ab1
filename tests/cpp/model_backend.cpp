// model_backend.cpp -- TEST DOUBLE, never shipped, never linked into the product: the handful of include/rio_cuda.h entry points
// that the host-only layers of librio_cuda (csrc/durable.cu: write-through into the reference's SQLite schema; csrc/resolver.cu:
// the micro-batching front end) are built on, answered by an in-memory model with LocalObjectPlacement's semantics
// (rio-rs/src/object_placement/local.rs:12-68: a map, update(None) removes, clean_server = retain) and the placement policy of
// Service::get_or_create_placement (rio-rs/src/service.rs:193-254) for RIO_PLACE_SELF.
//
// Why it exists: those two layers are plain C++ over the PUBLIC C ABI, so on a box without a GPU they can be compiled as C++ against
// this double and driven by the same conformance harness the GPU box runs against the real engine (tests/cpp/durable_conformance.cpp),
// and the resolver's queue can be run under ThreadSanitizer (tests/test_host_layers_cpu.py).  The product has no CPU path: the real
// rio_cuda_create fails without a CUDA device, and nothing under rio_rs_b200/ or include/ refers to this file.
#include <atomic>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/rio_cuda.h"

namespace {

uint64_t mix64(uint64_t x) {   // DESIGN.md 3.1 (splitmix64 finaliser)
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}
uint64_t fnv1a(const char *p, size_t n, uint64_t h = 0xCBF29CE484222325ull) {
    for (size_t i = 0; i < n; i++) { h ^= (uint8_t)p[i]; h *= 0x100000001B3ull; }
    return h;
}
thread_local std::string t_err;

struct Node { std::string addr; bool active = false, malformed = false; };

}  // namespace

struct rio_placement {
    std::mutex mu;
    std::vector<Node> nodes;
    std::unordered_map<std::string, uint32_t> index;
    std::unordered_map<uint64_t, uint32_t> dir;   // key -> node index (absent = no placement)
    std::atomic<uint64_t> batched_calls{0};
    uint32_t intern(const std::string &a) {
        auto it = index.find(a);
        if (it != index.end()) return it->second;
        Node n;
        n.addr = a;
        const size_t c = a.find(':');
        n.malformed = c == std::string::npos || c == 0 || c + 1 >= a.size();   // service.rs:205-213
        nodes.push_back(n);
        index.emplace(a, (uint32_t)nodes.size() - 1);
        return (uint32_t)nodes.size() - 1;
    }
    void clean(uint32_t idx) {                                                   // local.rs:51-58
        for (auto it = dir.begin(); it != dir.end();) it = it->second == idx ? dir.erase(it) : std::next(it);
    }
};

extern "C" {

uint32_t rio_cuda_abi_version(void) { return RIO_ABI_VERSION; }
const char *rio_cuda_last_error(rio_placement *) { return t_err.c_str(); }

rio_status rio_cuda_create(const rio_config *, rio_placement **out) {
    if (!out) return RIO_ERR_UNKNOWN;
    *out = new rio_placement();
    return RIO_OK;
}
void rio_cuda_destroy(rio_placement *h) { delete h; }

uint64_t rio_cuda_object_key(const char *type, size_t type_len, const char *id, size_t id_len) {
    uint64_t h = fnv1a(type, type_len);
    h = fnv1a(".", 1, h);
    return mix64(fnv1a(id, id_len, h));
}

rio_status rio_cuda_hash_ids(rio_placement *h, const char *packed, const uint64_t *offsets, size_t n, uint64_t *out_keys) {
    if (!h || (n && (!packed || !offsets || !out_keys))) { t_err = "null buffer"; return RIO_ERR_UNKNOWN; }
    for (size_t i = 0; i < n; i++) out_keys[i] = mix64(fnv1a(packed + offsets[i], (size_t)(offsets[i + 1] - offsets[i])));
    return RIO_OK;
}

rio_status rio_cuda_set_nodes(rio_placement *h, const char *const *addrs, const uint32_t *, const float *, uint32_t M, uint32_t, uint32_t *out_idx) {
    std::lock_guard<std::mutex> g(h->mu);
    for (auto &n : h->nodes) n.active = false;
    for (uint32_t j = 0; j < M; j++) {
        const uint32_t idx = h->intern(addrs[j]);
        h->nodes[idx].active = true;
        if (out_idx) out_idx[j] = idx;
    }
    return RIO_OK;
}
rio_status rio_cuda_node_set_active(rio_placement *h, uint32_t idx, int32_t active) {
    std::lock_guard<std::mutex> g(h->mu);
    if (idx >= h->nodes.size()) { t_err = "node index out of range"; return RIO_ERR_UNKNOWN; }
    h->nodes[idx].active = active != 0;
    return RIO_OK;
}
rio_status rio_cuda_node_intern(rio_placement *h, const char *address, uint32_t *out_idx) {
    std::lock_guard<std::mutex> g(h->mu);
    *out_idx = h->intern(address);
    return RIO_OK;
}
rio_status rio_cuda_node_address(rio_placement *h, uint32_t idx, char *buf, size_t cap, size_t *out_len) {
    std::lock_guard<std::mutex> g(h->mu);
    if (idx >= h->nodes.size()) { t_err = "node index out of range"; return RIO_ERR_UNKNOWN; }
    const std::string &a = h->nodes[idx].addr;
    if (out_len) *out_len = a.size();
    if (buf && cap) memcpy(buf, a.data(), cap < a.size() ? cap : a.size());
    return RIO_OK;
}
rio_status rio_cuda_node_state(rio_placement *h, uint32_t idx, int32_t *active, uint32_t *weight, int32_t *malformed) {
    std::lock_guard<std::mutex> g(h->mu);
    if (idx >= h->nodes.size()) { t_err = "node index out of range"; return RIO_ERR_UNKNOWN; }
    if (active) *active = h->nodes[idx].active;
    if (weight) *weight = 1;
    if (malformed) *malformed = h->nodes[idx].malformed;
    return RIO_OK;
}

rio_status rio_cuda_lookup_batch(rio_placement *h, const uint64_t *keys, size_t n, uint32_t *out_idx) {
    std::lock_guard<std::mutex> g(h->mu);
    h->batched_calls++;
    for (size_t i = 0; i < n; i++) { auto it = h->dir.find(keys[i]); out_idx[i] = it == h->dir.end() ? RIO_NONE : it->second; }
    return RIO_OK;
}
rio_status rio_cuda_upsert_batch(rio_placement *h, const uint64_t *keys, const uint32_t *idx, size_t n) {   // array order: the last occurrence wins
    std::lock_guard<std::mutex> g(h->mu);
    h->batched_calls++;
    for (size_t i = 0; i < n; i++) { if (idx[i] == RIO_NONE) h->dir.erase(keys[i]); else h->dir[keys[i]] = idx[i]; }
    return RIO_OK;
}
rio_status rio_cuda_directory_len(rio_placement *h, uint64_t *out_placed, uint64_t *out_slots) {
    std::lock_guard<std::mutex> g(h->mu);
    if (out_placed) *out_placed = h->dir.size();
    if (out_slots) *out_slots = h->dir.bucket_count();
    return RIO_OK;
}

// service.rs:193-254 per id, in array order (RIO_PLACE_SELF only: the double has no solver)
rio_status rio_cuda_place_batch(rio_placement *h, const uint64_t *keys, size_t n, uint32_t policy, uint32_t self_idx, uint32_t *out_idx) {
    std::lock_guard<std::mutex> g(h->mu);
    h->batched_calls++;
    if (policy != RIO_PLACE_SELF || self_idx >= h->nodes.size()) { t_err = "the test double places with RIO_PLACE_SELF only"; return RIO_ERR_UNKNOWN; }
    for (size_t i = 0; i < n; i++) {
        auto it = h->dir.find(keys[i]);
        if (it != h->dir.end()) {
            const Node &nd = h->nodes[it->second];
            if (!nd.malformed && nd.active) { out_idx[i] = it->second; continue; }   // :226-231
            if (nd.malformed) h->dir.erase(it);                                       // :213-222
            else h->clean(it->second);                                                // :233-237
        }
        h->dir[keys[i]] = self_idx;                                                   // :244-252
        out_idx[i] = self_idx;
    }
    return RIO_OK;
}

rio_status rio_cuda_update_str(rio_placement *h, const char *type, size_t type_len, const char *id, size_t id_len, const char *address, size_t address_len) {
    std::lock_guard<std::mutex> g(h->mu);
    const uint64_t key = rio_cuda_object_key(type, type_len, id, id_len);
    if (!address) h->dir.erase(key);
    else h->dir[key] = h->intern(std::string(address, address_len));
    return RIO_OK;
}
rio_status rio_cuda_lookup_str(rio_placement *h, const char *type, size_t type_len, const char *id, size_t id_len, char *buf, size_t cap, size_t *out_len) {
    std::lock_guard<std::mutex> g(h->mu);
    auto it = h->dir.find(rio_cuda_object_key(type, type_len, id, id_len));
    if (it == h->dir.end()) { *out_len = (size_t)-1; return RIO_OK; }
    const std::string &a = h->nodes[it->second].addr;
    *out_len = a.size();
    if (buf && cap) memcpy(buf, a.data(), cap < a.size() ? cap : a.size());
    return RIO_OK;
}
rio_status rio_cuda_clean_server_str(rio_placement *h, const char *address, size_t address_len) {
    std::lock_guard<std::mutex> g(h->mu);
    auto it = h->index.find(std::string(address, address_len));
    if (it != h->index.end()) h->clean(it->second);
    return RIO_OK;
}

// test-only: how many batched engine calls the double has served (the resolver test counts coalescing with it)
uint64_t model_backend_batched_calls(rio_placement *h) { return h->batched_calls.load(); }

}  // extern "C"
