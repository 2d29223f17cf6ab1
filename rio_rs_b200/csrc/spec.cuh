// spec.cuh -- solver spec v1 for host and device code of librio_cuda.
// Written against the spec text, NOT shared with oracle/ (the oracle restates it independently in C).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define RIO_HD __host__ __device__ __forceinline__
#define RIO_D __device__ __forceinline__
#else
#define RIO_HD inline
#define RIO_D inline
#endif

namespace rio {

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr uint64_t kSaltObj = 0xD6E8FEB86659FD93ull;
constexpr uint64_t kSaltNode2 = 0xA0761D6478BD642Full;
constexpr uint64_t kSaltSpill = 0x2545F4914F6CDD1Dull;
constexpr uint64_t kGolden64 = 0x9E3779B97F4A7C15ull;
constexpr uint64_t kSaltPos = 0x8CB92BA72F3D8DD7ull;     // HRW2: node position in the trie
constexpr uint64_t kSaltLevel = 0x3C79AC492BA7B653ull;   // HRW2: pseudo-node seed of a trie level
constexpr uint32_t kPairC1 = 0x9E3779B1u;
constexpr uint32_t kLogK0 = 0x71376877u, kLogK1 = 0x44D58AB6u, kLogK2 = 0x2677DB2Eu, kLogK3 = 0x0B98D5FAu;
constexpr uint64_t kFnvBasis = 0xCBF29CE484222325ull, kFnvPrime = 0x100000001B3ull;

RIO_HD uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

RIO_HD uint32_t mulhi_u32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

RIO_HD uint32_t clz_u32(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__clz((int)u);
#else
    return u ? (uint32_t)__builtin_clz(u) : 32u;
#endif
}

// Q32 log2(1+F/2^32), monotone non-decreasing
RIO_HD uint32_t log2frac(uint32_t F) {
    uint32_t t2 = kLogK2 - mulhi_u32(F, kLogK3);
    uint32_t t1 = kLogK1 - mulhi_u32(F, t2);
    uint32_t g = kLogK0 - mulhi_u32(F, t1);
    uint32_t q = mulhi_u32(F, ~F);
    return F + mulhi_u32(q, g);
}

// E(u): Q26 fixed point -log2(u / 2^32), monotone non-increasing in u
RIO_HD uint32_t elog(uint32_t u) {
    uint32_t lz = clz_u32(u);
    uint32_t m = lz < 32 ? (u << lz) : 0u;
    return ((lz + 1u) << 26) - (log2frac(m << 1) >> 6);
}

struct ObjHash { uint32_t b, ab; };

RIO_HD ObjHash obj_hash(uint64_t key) {
    uint64_t h = mix64(key ^ kSaltObj);
    uint32_t a = (uint32_t)h;
    ObjHash o;
    o.b = (uint32_t)(h >> 32) | 1u;
    o.ab = a * o.b;
    return o;
}

// u(key, node), spec v3: s0 = lo32(seed), m = hi32(seed) | 1 (stored pre-or'ed in the node records),
// s2 = lo32(mix64(seed ^ kSaltNode2)).  SASS per pair: IMAD, IMAD.
RIO_HD uint32_t pair_hash(ObjHash o, uint32_t s0, uint32_t m, uint32_t s2) {
    const uint32_t p = s0 * o.b + o.ab;
    return p * m + s2;
}

// ---- HRW2: a contest is "v(key, seed) < T", v the pair hash reduced to 31 bits.  The kernels evaluate it
// as u = p * (2m) + (2h + 1) = 2v + 1 (always odd) against T3 = max(2T - 1, 0): u <= T3  <=>  v < T, for every T in [0, 2^31]
// -- both forced outcomes (T = 0: never, T = 2^31: always) are representable in one unsigned 32-bit compare.
struct ContestRec { uint32_t s0, m2, h2; };   // m2 = 2 * (hi32(seed) | 1), h2 = 2 * (s2 & 0x7FFFFFFF) + 1
RIO_HD ContestRec contest_rec(uint64_t seed) {
    const uint32_t s2 = (uint32_t)mix64(seed ^ kSaltNode2);
    ContestRec r;
    r.s0 = (uint32_t)seed;
    r.m2 = ((uint32_t)(seed >> 32) | 1u) << 1;
    r.h2 = (s2 << 1) | 1u;
    return r;
}
RIO_HD uint32_t contest_u(ObjHash o, uint32_t s0, uint32_t m2, uint32_t h2) {
    const uint32_t p = s0 * o.b + o.ab;
    return p * m2 + h2;
}
inline uint64_t level_seed(uint32_t level) { return mix64(kGolden64 * ((uint64_t)level + 1) ^ kSaltLevel); }
// T3 of a contest between a left part of weight wl and a right part of weight wr
inline uint32_t contest_t3(uint64_t wl, uint64_t wr) {
    if (wl == 0) return 0u;                      // never LEFT (also the empty subtree)
    if (wr == 0) return 0xFFFFFFFFu;             // T = 2^31: always LEFT
    const uint64_t t = wl < (1ull << 33) ? (wl << 31) / (wl + wr)                                    // the common case fits 64 bits
                                         : (uint64_t)((((unsigned __int128)wl) << 31) / (wl + wr));   // <= 2^31
    return t ? (uint32_t)(2 * t - 1) : 0u;
}

RIO_HD uint32_t inv_weight(uint32_t w) { return w ? 0xFFFFFFFFu / w : 0u; }

RIO_HD uint32_t spill_hash(uint64_t key, uint32_t round) {
    return (uint32_t)(mix64(key ^ (kSaltSpill + (uint64_t)round * kGolden64)) >> 32);
}

RIO_HD uint64_t synth_key(uint64_t i, uint64_t seed) { return mix64(kGolden64 * (i + 1) ^ seed); }

inline uint64_t fnv1a64(const char *p, size_t n, uint64_t h = kFnvBasis) {
    for (size_t i = 0; i < n; i++) { h ^= (uint8_t)p[i]; h *= kFnvPrime; }
    return h;
}

// lexicographic (score, ~u, idx) "a beats b"
RIO_HD bool cand_better(uint64_t sa, uint32_t ua, uint32_t ia, uint64_t sb, uint32_t ub, uint32_t ib) {
    return sa < sb || (sa == sb && (ua > ub || (ua == ub && ia < ib)));
}

}  // namespace rio
