"""Third, independent restatement of the solver spec in pure Python ints.

Used only to cross-check oracle/rio_oracle.c on small cases and to generate tests/golden/.
Slow by construction (Python loops) -- keep inputs tiny.
"""
M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF
NONE = 0xFFFFFFFF


def mix64(x):
    x &= M64
    x ^= x >> 30
    x = x * 0xBF58476D1CE4E5B9 & M64
    x ^= x >> 27
    x = x * 0x94D049BB133111EB & M64
    x ^= x >> 31
    return x


def fnv1a64(data):
    h = 0xCBF29CE484222325
    for b in data:
        h = (h ^ b) * 0x100000001B3 & M64
    return h


def object_key(type_, id_):
    return mix64(fnv1a64((type_ + "." + id_).encode()))


def node_seed(address):
    return mix64(fnv1a64(address.encode()))


def log2frac(F):
    K0, K1, K2, K3 = 0x71376877, 0x44D58AB6, 0x2677DB2E, 0x0B98D5FA
    t2 = K2 - (F * K3 >> 32)
    t1 = K1 - (F * t2 >> 32)
    g = K0 - (F * t1 >> 32)
    q = F * (~F & M32) >> 32
    return F + (q * g >> 32)


def elog(u):
    lz = 32 - u.bit_length()
    m = (u << lz) & M32 if lz < 32 else 0
    L = log2frac((m << 1) & M32)
    return ((lz + 1) << 26) - (L >> 6)


def pair_hash(key, seed):
    h = mix64(key ^ 0xD6E8FEB86659FD93)
    a = h & M32
    b = (h >> 32) | 1
    ab = a * b & M32
    s0, s1 = seed & M32, seed >> 32
    s2 = mix64(seed ^ 0xA0761D6478BD642F) & M32
    p = (s0 * b + ab) & M32
    return (p * (s1 | 1) + s2) & M32


def inv_weight(w):
    return M32 // w if w else 0


def hrw(key, seeds, weights, closed=()):
    best = None
    for j, (s, w) in enumerate(zip(seeds, weights)):
        if not w or j in closed:
            continue
        u = pair_hash(key, s)
        cand = (elog(u) * inv_weight(w), M32 - u, j)
        if best is None or cand < best:
            best = cand
    return NONE if best is None else best[2]


def spill_hash(key, rnd):
    return mix64(key ^ ((0x2545F4914F6CDD1D + rnd * 0x9E3779B97F4A7C15) & M64)) >> 32


def capacity(n_total, w, w_sum, num, den):
    if not w or not w_sum or not den:
        return 0
    return min(M32, -(-(num * n_total * w) // (den * w_sum)))


def assign_bounded(keys, seeds, weights, num=5, den=4, max_rounds=4):
    n, M = len(keys), len(seeds)
    W = sum(weights)
    cap = [capacity(n, w, W, num, den) for w in weights]
    idx = [hrw(k, seeds, weights) for k in keys]
    closed = set()
    passes = 1
    for r in range(1, max_rounds):
        c = [0] * M
        for j in idx:
            if j != NONE:
                c[j] += 1
        over = {j for j in range(M) if weights[j] and c[j] > cap[j]}
        closed |= over
        open_ = [j for j in range(M) if weights[j] and j not in closed]
        if not over or not open_:
            break
        thr = {j: ((c[j] - cap[j]) << 32) // c[j] for j in over}
        for i, k in enumerate(keys):
            j = idx[i]
            if j in over and spill_hash(k, r) < thr[j]:
                idx[i] = hrw(k, seeds, weights, closed)
        passes += 1
    c = [0] * M
    for j in idx:
        if j != NONE:
            c[j] += 1
    return idx, c, passes


def synth_key(i, seed):
    return mix64((0x9E3779B97F4A7C15 * (i + 1) & M64) ^ seed)


# ---- 3.8 HRW2: hierarchical weighted rendezvous with fan-out 2 -----------------------------------------------------
# Written from the spec text as a literal recursion over sets of members (no sorting tricks, no prefix sums, no heap):
# deliberately the slowest and most obvious of the three implementations.
SALT_POS = 0x8CB92BA72F3D8DD7
SALT_LVL = 0x3C79AC492BA7B653


def hrw2_v(key, seed):
    h = mix64(key ^ 0xD6E8FEB86659FD93)
    a = h & M32
    b = (h >> 32) | 1
    ab = a * b & M32
    p = ((seed & M32) * b + ab) & M32
    m = (seed >> 32) | 1
    s2 = mix64(seed ^ 0xA0761D6478BD642F) & M32
    return (p * m + (s2 & 0x7FFFFFFF)) & 0x7FFFFFFF


def hrw2_level_seed(level):
    return mix64((0x9E3779B97F4A7C15 * (level + 1) & M64) ^ SALT_LVL)


def hrw2_contest_left(key, contest_seed, wl, wr):
    t = (wl << 31) // (wl + wr) if wl + wr else 0
    return hrw2_v(key, contest_seed) < t


def hrw2(key, seeds, weights, closed=(), bits=12):
    members = [(mix64(s ^ SALT_POS), j) for j, (s, w) in enumerate(zip(seeds, weights)) if w and j not in closed]
    if not members:
        return NONE
    for level in range(bits):
        shift = 63 - level
        left = [m for m in members if not (m[0] >> shift) & 1]
        right = [m for m in members if (m[0] >> shift) & 1]
        wl = sum(weights[j] for _, j in left)
        wr = sum(weights[j] for _, j in right)
        members = left if hrw2_contest_left(key, hrw2_level_seed(level), wl, wr) else right
    members.sort()
    while len(members) > 1:
        (_, j), rest = members[0], members[1:]
        if hrw2_contest_left(key, seeds[j], weights[j], sum(weights[i] for _, i in rest)):
            return j
        members = rest
    return members[0][1]


def assign_bounded_hrw2(keys, seeds, weights, num=5, den=4, max_rounds=4, bits=12):
    n, M = len(keys), len(seeds)
    W = sum(weights)
    cap = [capacity(n, w, W, num, den) for w in weights]
    idx = [hrw2(k, seeds, weights, (), bits) for k in keys]
    closed = set()
    passes = 1
    for r in range(1, max_rounds):
        c = [0] * M
        for j in idx:
            if j != NONE:
                c[j] += 1
        over = {j for j in range(M) if weights[j] and c[j] > cap[j]}
        closed |= over
        open_ = [j for j in range(M) if weights[j] and j not in closed]
        if not over or not open_:
            break
        thr = {j: ((c[j] - cap[j]) << 32) // c[j] for j in over}
        for i, k in enumerate(keys):
            j = idx[i]
            if j in over and spill_hash(k, r) < thr[j]:
                idx[i] = hrw2(k, seeds, weights, closed, bits)
        passes += 1
    c = [0] * M
    for j in idx:
        if j != NONE:
            c[j] += 1
    return idx, c, passes
