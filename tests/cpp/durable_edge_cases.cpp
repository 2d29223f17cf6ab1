// durable_edge_cases.cpp -- corner cases of the native write-through (csrc/durable.cu) that the reference's own tests do not reach:
// server addresses longer than any fixed buffer, NULL ids in a batch, a batch with nothing to place.  CPU only, against the test
// double (tests/cpp/model_backend.cpp); built and run by tests/test_host_layers_cpu.py.
// usage: durable_edge_cases <scratch directory>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rio_cuda.h"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s  [%s]\n", __FILE__, __LINE__, #c, rio_cuda_durable_last_error()); std::exit(1); } } while (0)

int main(int argc, char **argv) {
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    const std::string path = dir + "/edge.sqlite3";
    const std::string long_addr = std::string(700, 'h') + ".example:5000";       // 713 bytes: longer than every stack buffer there was
    std::vector<char> big(2048);
    {
        rio_placement *h = nullptr;
        rio_config cfg{sizeof(rio_config), -1, 0, 0, 0};
        CHECK(rio_cuda_create(&cfg, &h) == RIO_OK);
        const char *nodes[2] = {long_addr.c_str(), "0.0.0.0:5001"};
        uint32_t nidx[2];
        CHECK(rio_cuda_set_nodes(h, nodes, nullptr, nullptr, 2, 0, nidx) == RIO_OK);
        rio_durable *d = nullptr;
        CHECK(rio_cuda_durable_open(h, path.c_str(), &d) == RIO_OK);
        const char *t[3] = {"Obj", "Obj", "Obj"}, *i[3] = {"a", "b", "c"};
        uint32_t out[3] = {0, 0, 0};
        CHECK(rio_cuda_durable_place_batch(d, t, i, 3, RIO_PLACE_SELF, nidx[0], out) == RIO_OK);      // claimed by the long-named server
        CHECK(out[0] == nidx[0] && out[2] == nidx[0]);
        size_t n = 0;
        CHECK(rio_cuda_durable_lookup(d, "Obj", 3, "a", 1, big.data(), big.size(), &n) == RIO_OK && n == long_addr.size() && !memcmp(big.data(), long_addr.data(), n));
        CHECK(rio_cuda_node_set_active(h, nidx[0], 0) == RIO_OK);                                     // it dies: clean_server(long address) must hit its rows
        CHECK(rio_cuda_durable_place_batch(d, t, i, 1, RIO_PLACE_SELF, nidx[1], out) == RIO_OK);
        CHECK(out[0] == nidx[1]);
        const char *bad_t[2] = {"Obj", nullptr}, *bad_i[2] = {"x", "y"};
        CHECK(rio_cuda_durable_place_batch(d, bad_t, bad_i, 2, RIO_PLACE_SELF, nidx[1], out) == RIO_ERR_UNKNOWN);   // refused, not a crash
        CHECK(rio_cuda_durable_place_batch(d, t, i, 0, RIO_PLACE_SELF, nidx[1], out) == RIO_OK);      // empty batch
        rio_cuda_durable_close(d);
        rio_cuda_destroy(h);
    }
    rio_placement *h = nullptr;
    rio_config cfg{sizeof(rio_config), -1, 0, 0, 0};
    CHECK(rio_cuda_create(&cfg, &h) == RIO_OK);
    rio_durable *d = nullptr;
    CHECK(rio_cuda_durable_open(h, path.c_str(), &d) == RIO_OK);
    uint64_t rows = 0;
    CHECK(rio_cuda_durable_recover(d, &rows) == RIO_OK);
    CHECK(rows == 1);                                                             // "a" on the survivor; "b" and "c" went with the dead server's rows
    size_t n = 0;
    CHECK(rio_cuda_durable_lookup(d, "Obj", 3, "a", 1, big.data(), big.size(), &n) == RIO_OK && std::string(big.data(), n) == "0.0.0.0:5001");
    CHECK(rio_cuda_durable_lookup(d, "Obj", 3, "b", 1, big.data(), big.size(), &n) == RIO_OK && n == (size_t)-1);
    rio_cuda_durable_close(d);
    rio_cuda_destroy(h);
    std::printf("durable edge cases: all passed\n");
    return 0;
}
