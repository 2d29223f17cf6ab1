// trie_table_selftest.cpp -- csrc/trie_table.hpp against the plainest statement of the same table (a full sort by (pos, idx), the
// reference threshold division of spec.cuh), byte for byte, and contest_t3_fast against contest_t3 operand by operand.  The shipped
// builder avoids the sort and the 64-bit divisions (a table rebuild is on the critical path of every membership event); this is what
// keeps those shortcuts honest.  Built and run by tests/test_client_first_hop.py (CPU).
#include <cstdio>
#include <cstring>
#include <random>

#include "../../rio_rs_b200/csrc/trie_table.hpp"

using namespace rio;

// DESIGN.md 3.8 / 4.1 read literally
static std::vector<uint32_t> plain_blob(const std::vector<TrieMember> &members, uint32_t bits) {
    const uint32_t nb = 1u << bits;
    struct Mem { uint64_t pos, seed; uint32_t idx, w; };
    std::vector<Mem> mem;
    for (const TrieMember &m : members) if (m.weight) mem.push_back(Mem{mix64(m.seed ^ kSaltPos), m.seed, m.idx, m.weight});
    std::sort(mem.begin(), mem.end(), [](const Mem &a, const Mem &b) { return a.pos != b.pos ? a.pos < b.pos : a.idx < b.idx; });
    std::vector<std::vector<Mem>> bucket(nb);
    for (const Mem &m : mem) bucket[bits ? (uint32_t)(m.pos >> (64 - bits)) : 0u].push_back(m);
    std::vector<uint64_t> wsum((size_t)2 * nb, 0);
    for (uint32_t k = 0; k < nb; k++) for (const Mem &m : bucket[k]) wsum[nb + k] += m.w;
    for (uint32_t i = nb - 1; i >= 1; i--) wsum[i] = wsum[2 * i] + wsum[2 * i + 1];
    std::vector<uint32_t> words((size_t)2 * nb, 0);
    for (uint32_t i = 1; i < nb; i++) words[i] = contest_t3(wsum[2 * i], wsum[2 * i + 1]);
    while (words.size() % 4) words.push_back(0);
    const uint32_t off_crec = (uint32_t)words.size() * 4;
    for (uint32_t k = 0; k < nb; k++) {
        const std::vector<Mem> &b = bucket[k];
        if (b.empty()) { words[nb + k] = kNone; continue; }
        if (b.size() == 1) { words[nb + k] = b[0].idx; continue; }
        words[nb + k] = 0x80000000u | (uint32_t)(words.size() * 4);
        uint64_t rest = wsum[nb + k];
        for (size_t q = 0; q + 1 < b.size(); q++) {
            rest -= b[q].w;
            const ContestRec r = contest_rec(b[q].seed);
            const uint32_t here = (uint32_t)(words.size() * 4);
            const uint32_t rec[8] = {r.s0, r.m2, r.h2, contest_t3(b[q].w, rest), b[q].idx, q + 2 == b.size() ? b.back().idx : 0x80000000u | (here + 32u), 0u, 0u};
            words.insert(words.end(), rec, rec + 8);
        }
    }
    (void)off_crec;
    if (words.size() < 4) words.resize(4, 0);
    return words;
}

int main() {
    std::mt19937_64 rng(2026);
    uint64_t pairs = 0;
    const uint64_t edges[] = {0, 1, 2, 3, 15, 16, 17, 255, 65535, 65536, (1ull << 31) - 1, 1ull << 31, (1ull << 31) + 1, (1ull << 32) - 2, (1ull << 32) - 1, 1ull << 32, 1ull << 33, (1ull << 40) + 7};
    for (uint64_t a : edges) for (uint64_t b : edges) { if (contest_t3_fast(a, b) != contest_t3(a, b)) { std::printf("T3 mismatch %llu %llu\n", (unsigned long long)a, (unsigned long long)b); return 1; } pairs++; }
    for (int i = 0; i < 3000000; i++) {
        const uint64_t a = rng() >> (63 - rng() % 34), b = rng() >> (63 - rng() % 34);
        if (contest_t3_fast(a, b) != contest_t3(a, b)) { std::printf("T3 mismatch %llu %llu\n", (unsigned long long)a, (unsigned long long)b); return 1; }
        pairs++;
    }
    for (int i = 0; i < 1000000; i++) {   // quotients that sit on or next to an exact multiple
        const uint64_t s = 2 + (rng() >> 33), q = rng() % ((1ull << 31) + 1);
        const uint64_t wl = (uint64_t)(((unsigned __int128)q * s) >> 31) + (rng() % 3) - 1;
        if (wl == 0 || wl >= s) continue;
        if (contest_t3_fast(wl, s - wl) != contest_t3(wl, s - wl)) { std::printf("T3 mismatch near a multiple %llu %llu\n", (unsigned long long)wl, (unsigned long long)(s - wl)); return 1; }
        pairs++;
    }
    int cases = 0;
    for (uint32_t bits : {0u, 1u, 2u, 3u, 4u, 8u, 10u, 12u, 14u})
        for (uint32_t M : {0u, 1u, 2u, 3u, 17u, 64u, 300u, 1024u, 3000u})
            for (int rep = 0; rep < 4; rep++) {
                std::vector<TrieMember> m;
                for (uint32_t j = 0; j < M; j++) {
                    uint64_t seed = rng();
                    uint32_t w = rep == 0 ? 1 : (uint32_t)(rng() % 17);
                    if (rep == 2 && j % 7 == 0) w = 0xFFFFFFFFu - (uint32_t)(rng() % 3);
                    if (rep == 3) w = (uint32_t)(rng() >> 32);
                    if (j > 2 && j % 50 == 0) seed = m[j - 1].seed;          // equal positions: ordered by node index
                    m.push_back(TrieMember{seed, j, w});
                }
                std::shuffle(m.begin(), m.end(), rng);
                const TrieBlob b = build_trie_blob(m, bits);
                const std::vector<uint32_t> want = plain_blob(m, bits);
                if (b.words != want || b.blob_bytes != want.size() * 4 || b.bits != bits) { std::printf("blob mismatch: bits %u, %u members, variant %d\n", bits, M, rep); return 1; }
                cases++;
            }
    std::printf("trie table selftest: all passed (%llu threshold operand pairs, %d blobs)\n", (unsigned long long)pairs, cases);
    return 0;
}
