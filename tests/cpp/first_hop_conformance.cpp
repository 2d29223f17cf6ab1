// C++ conformance harness for rio::FirstHop (rio_rs_b200/host/first_hop.hpp) -- CPU only.
// argv: a file with one "address weight" per line, then a file with "type id expected_address" lines produced by the
// oracle (tests/test_client_first_hop.py writes both).  Also replays the reference client's cache behaviour.
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../rio_rs_b200/host/first_hop.hpp"

#define CHECK(c) do { if (!(c)) { std::printf("FAILED line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main(int argc, char **argv) {
    if (argc != 3) return 2;
    std::vector<std::string> addrs;
    std::vector<uint32_t> w;
    { std::ifstream f(argv[1]); std::string a; uint32_t x; while (f >> a >> x) { addrs.push_back(a); w.push_back(x); } }
    rio::FirstHop fh(3);
    bool threw = false;
    try { fh.get_service_object_address("Obj", "1"); } catch (const rio::NoServersAvailable &) { threw = true; }
    CHECK(threw);                                                     // empty view: ClientError::NoServersAvailable
    fh.set_active_servers(addrs, w);
    size_t n = 0;
    { std::ifstream f(argv[2]); std::string t, i, want; while (f >> t >> i >> want) { CHECK(fh.get_service_object_address(t, i) == want); n++; } }
    CHECK(n > 100);
    const std::string owner = fh.get_service_object_address("Obj", "1");
    std::string other;
    for (const auto &a : addrs) if (a != owner) { other = a; break; }
    fh.record_redirect("Obj", "1", other);                            // tower_services.rs:158-168
    CHECK(fh.get_service_object_address("Obj", "1") == other);
    for (int k = 2; k < 6; k++) fh.record_redirect("Obj", std::to_string(k), other);
    CHECK(fh.get_service_object_address("Obj", "1") == owner);        // evicted by the LRU limit -> the hash again
    std::printf("all passed (%zu ids)\n", n);
    return 0;
}
