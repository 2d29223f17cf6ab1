// backend_conformance.cpp -- the reference's provider conformance tests, restated against the C++ mirror of the trait.
// Each function cites the reference test it restates; exit code 0 = all passed.  Built and run by tests/test_gpu_cpp.py.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../rio_rs_b200/host/gpu_object_placement.hpp"

using namespace rio_rs;

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); std::exit(1); } } while (0)

// rio-rs/tests/object_placement_backend.rs:11-16
static void no_placement(GpuObjectPlacement provider) {
    provider.prepare();
    auto server_addr = provider.lookup(ObjectId("obj", "1"));
    CHECK(!server_addr.has_value());
}

// rio-rs/tests/object_placement_backend.rs:18-34
static void save_and_load(GpuObjectPlacement provider) {
    provider.prepare();
    ObjectId obj_id("obj", "1");
    provider.update(ObjectPlacementItem(obj_id, std::string("0.0.0.0:8888")));
    auto server_addr = provider.lookup(ObjectId("obj", "1"));
    CHECK(server_addr.has_value() && *server_addr == "0.0.0.0:8888");
    provider.clean_server("0.0.0.0:8888");
    server_addr = provider.lookup(ObjectId("obj", "1"));
    CHECK(!server_addr.has_value());
}

// rio-rs/src/object_placement/local.rs:75-114
static void provider_is_clonable() {
    GpuObjectPlacement provider;
    GpuObjectPlacement cloned_provider = provider.clone();
    provider.update(ObjectPlacementItem(ObjectId("test", "1"), std::string("0.0.0.0:80")));
    CHECK(provider.lookup(ObjectId("test", "1")).has_value());
    CHECK(cloned_provider.lookup(ObjectId("test", "1")).has_value());
    cloned_provider.clean_server("0.0.0.0:80");
    CHECK(!provider.lookup(ObjectId("test", "1")).has_value());
    CHECK(!cloned_provider.lookup(ObjectId("test", "1")).has_value());
}

// rio-rs/src/object_placement/sqlite.rs:149-193
static void overwrite_then_clean() {
    GpuObjectPlacement p;
    p.update(ObjectPlacementItem(ObjectId("Test", "1"), std::string("0.0.0.0:5000")));
    p.update(ObjectPlacementItem(ObjectId("Test", "1"), std::string("0.0.0.0:5001")));
    CHECK(*p.lookup(ObjectId("Test", "1")) == "0.0.0.0:5001");
    p.clean_server("0.0.0.0:5001");
    CHECK(!p.lookup(ObjectId("Test", "1")).has_value());
    p.update(ObjectPlacementItem(ObjectId("Test", "1"), std::nullopt));   // update(None): local.rs:34-38
    p.remove(ObjectId("Test", "1"));                                     // idempotent remove: local.rs:60-68
    CHECK(!p.lookup(ObjectId("Test", "1")).has_value());
}

// Service::get_or_create_placement, batched (service.rs:193-254; behaviour of tests/object_allocation.rs:75-137)
static void placement_policy() {
    GpuObjectPlacement p;
    auto idx = p.set_nodes({"0.0.0.0:5000", "0.0.0.0:5001"});
    std::vector<uint64_t> keys;
    for (int i = 0; i < 1000; i++) keys.push_back(GpuObjectPlacement::object_key(ObjectId("MockService", std::to_string(i))));
    auto first = p.place_batch(keys, RIO_PLACE_SELF, idx[0]);
    for (auto v : first) CHECK(v == idx[0]);                     // unallocated -> claimed by the serving node
    auto again = p.place_batch(keys, RIO_PLACE_SELF, idx[1]);
    for (auto v : again) CHECK(v == idx[0]);                     // still owned by node 0 (-> Redirect upstream)
    p.node_set_active(idx[0], false);                            // owner dies
    std::vector<uint64_t> some(keys.begin(), keys.begin() + 10);
    auto moved = p.place_batch(some, RIO_PLACE_SELF, idx[1]);
    for (auto v : moved) CHECK(v == idx[1]);                     // first_server != second_server
    auto rest = p.lookup_many(keys);
    for (size_t i = 10; i < rest.size(); i++) CHECK(rest[i] == RIO_NONE);   // clean_server dropped the dead node's objects
    CHECK(*p.lookup(ObjectId("MockService", "3")) == "0.0.0.0:5001");
}

int main() {
    try {
        no_placement(GpuObjectPlacement());
        save_and_load(GpuObjectPlacement());
        provider_is_clonable();
        overwrite_then_clean();
        placement_policy();
    } catch (const ObjectPlacementError &e) {
        std::fprintf(stderr, "ObjectPlacementError(%s): %s\n", e.kind == ObjectPlacementError::Upstream ? "Upstream" : "Unknown", e.what());
        return 2;
    }
    std::puts("backend_conformance: all passed");
    return 0;
}
