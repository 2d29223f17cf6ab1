// k_affinity_umma.cu -- affinity-cost placement on the 5th-gen tensor cores (tcgen05 + TMEM), K = 16.
//
//   cost_ij = -sum_k Fobj[i,k] * Fnode[j,k]   ->  per-object argmin over the live nodes   
//
// This is the one dense contraction on the path, so it is the one place tensor cores are used.  fp32 inputs are split
// on the fly into three bf16 pieces (a = a_h + a_m + a_l exactly) and the product is rebuilt from six cross terms
// (hh, hm, mh, mm, hl, lh; the dropped terms are <= 2^-24 relative), each term ONE tcgen05.mma (M=128 objects x
// N=NT nodes x K=16) accumulating in fp32 in TMEM.  Nothing is materialised in HBM: the N x M grid lives only in TMEM.
//
// Warp roles (384 threads, one CTA per SM, persistent over 128-object row blocks):
//   warps 0-1  producers : load 128 object rows (fp32, two per thread), split to bf16 h/m/l, store the three K-major
//                          16-byte-interleaved operand blocks into shared memory, fence.proxy.async, arrive a_full
//   warp  2    MMA issuer: one lane issues 6 tcgen05.mma per node tile into one of two TMEM accumulators,
//                          tcgen05.commit -> tmem_full (and -> a_empty after the last tile of the row block)
//   warp  3    idle (keeps the epilogue on warps 4-11, whose warp%4 selects the TMEM lane quarter they may read)
//   warps 4-11 epilogue  : two warps per lane quarter, each owning half of every tile's columns: tcgen05.ld (64 columns
//                          per wait), 3-input max over groups of 8 columns, running (best value, best group); arrive
//                          tmem_empty; after the last tile the halves merge through shared memory and emit the winning group.
// k_affinity_resolve (second pass, CUDA cores, one warp per object) re-evaluates the 8 candidates of each winning group in
// fp32: node index + fp32 cost + per-node histogram.  It is a separate kernel because any LDS/LDG issued while the tensor
// core streams K=16 operands out of shared memory crawls (profiles/r01_umma_role_cycles_*.txt).
// Node operands (3 bf16 blocks, 96 B per node) stay resident in shared memory for the whole kernel.
#include "kernels.cuh"
#include "spec.cuh"

#include <cuda_bf16.h>
#include <cstdlib>

namespace rio {

namespace {

constexpr int kUmmaThreads = 384;   // 12 warps = 3 per SMSP: 170 registers per thread
constexpr int kProducerThreads = 64;  // warps 0-1, two object rows per thread
constexpr int kRows = 128;          // objects per row block == UMMA M
constexpr int kStages = 2;          // A-operand stages
constexpr uint32_t kABlockBytes = kRows * 32;            // one bf16 term of a row block: [2 k-chunks][128 rows][16 B]
constexpr uint32_t kAStageBytes = 3 * kABlockBytes;      // h, m, l
constexpr uint32_t kBarBytes = 128 + 2 * 128 * 8;   // mbarriers + TMEM slot, then two (best value, best group) merge slots of 128 rows

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle ("interleave") shared-memory operand descriptor: 8-row x 16-byte core matrices,
// LBO = byte distance between the two 16-byte K chunks, SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
           (1ull << 46);   // descriptor version 1 (Blackwell), base offset 0, SWIZZLE_NONE
}

// fp32 -> three bf16 pieces with a == h + m + l exactly
__device__ __forceinline__ void split3(float a, __nv_bfloat16 &h, __nv_bfloat16 &m, __nv_bfloat16 &l) {
    h = __float2bfloat16_rn(a);
    const float r1 = a - __bfloat162float(h);
    m = __float2bfloat16_rn(r1);
    l = __float2bfloat16_rn(r1 - __bfloat162float(m));
}
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 lo, __nv_bfloat16 hi) {
    return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
}

// Split one fp32 row of 16 features and store it as row `r` of the three operand blocks at `base`
// (block x at base + x*block_bytes; inside a block: [k-chunk][row][16 B], chunk stride = chunk_stride bytes).
__device__ __forceinline__ void store_row_split(unsigned char *base, uint32_t block_bytes, uint32_t chunk_stride, uint32_t r, const float (&f)[16]) {
    __nv_bfloat16 h[16], m[16], l[16];
#pragma unroll
    for (int k = 0; k < 16; k++) split3(f[k], h[k], m[k], l[k]);
#pragma unroll
    for (int kc = 0; kc < 2; kc++) {
        uint4 vh = make_uint4(pack2(h[kc * 8 + 0], h[kc * 8 + 1]), pack2(h[kc * 8 + 2], h[kc * 8 + 3]), pack2(h[kc * 8 + 4], h[kc * 8 + 5]), pack2(h[kc * 8 + 6], h[kc * 8 + 7]));
        uint4 vm = make_uint4(pack2(m[kc * 8 + 0], m[kc * 8 + 1]), pack2(m[kc * 8 + 2], m[kc * 8 + 3]), pack2(m[kc * 8 + 4], m[kc * 8 + 5]), pack2(m[kc * 8 + 6], m[kc * 8 + 7]));
        uint4 vl = make_uint4(pack2(l[kc * 8 + 0], l[kc * 8 + 1]), pack2(l[kc * 8 + 2], l[kc * 8 + 3]), pack2(l[kc * 8 + 4], l[kc * 8 + 5]), pack2(l[kc * 8 + 6], l[kc * 8 + 7]));
        const uint32_t off = kc * chunk_stride + r * 16;
        *reinterpret_cast<uint4 *>(base + 0 * block_bytes + off) = vh;
        *reinterpret_cast<uint4 *>(base + 1 * block_bytes + off) = vm;
        *reinterpret_cast<uint4 *>(base + 2 * block_bytes + off) = vl;
    }
}

__device__ __forceinline__ float max3f(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

struct UmmaParams {
    const float *fobj;        // n x 16
    uint64_t n;
    const float *fnode_c;     // m_pad x 16 fp32, live nodes compacted in node-index order, zero padded
    const uint32_t *nidx_map; // compacted position -> interned node index
    uint32_t n_live, m_pad;
    uint32_t *out_idx;
    float *out_cost;          // nullable
    uint32_t *counters;       // nullable
    uint32_t lbo_a, sbo_a, lbo_b, sbo_b;   // descriptor strides in bytes
    unsigned long long *timing;            // optional per-CTA cycle counters (16 per CTA) for tools/umma_timing.py; NULL in production
};

template <int NT, int LDW>
__global__ void __launch_bounds__(kUmmaThreads, 1) k_affinity_umma(UmmaParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *a_full = reinterpret_cast<uint64_t *>(smem);         // [kStages]
    uint64_t *a_empty = a_full + kStages;                          // [kStages]
    constexpr uint32_t NBUF = (NT == 128) ? 4 : 2;                  // TMEM accumulator buffers of NT columns each
    uint64_t *t_full = a_empty + kStages;                          // [NBUF]
    uint64_t *t_empty = t_full + 4;                                // [NBUF]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(t_empty + 4);
    float2 *sMerge = reinterpret_cast<float2 *>(smem + 128);          // [2][128]: (best value, best group bits) handed from column half B to half A
    unsigned char *sB = smem + kBarBytes;                          // 3 blocks of m_pad*32 bytes
    const uint32_t b_block_bytes = P.m_pad * 32;
    unsigned char *sA = sB + 3 * b_block_bytes;                    // kStages stages of 3 blocks

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t_begin = P.timing ? clock64() : 0;
    const uint32_t n_tiles = P.m_pad / NT;
    const uint64_t n_rb = (P.n + kRows - 1) / kRows;

    // ---- one-time setup: barriers, TMEM, node operands ----------------------------------------------------------
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(&a_full[s], kProducerThreads); mbar_init(&a_empty[s], 1); }
        for (uint32_t b = 0; b < NBUF; b++) { mbar_init(&t_full[b], 1); mbar_init(&t_empty[b], 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(NBUF * NT)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (uint32_t p = threadIdx.x; p < P.m_pad; p += blockDim.x) {
        float f[16];
        const float4 *row = reinterpret_cast<const float4 *>(P.fnode_c + (size_t)p * 16);
#pragma unroll
        for (int q = 0; q < 4; q++) { const float4 v = __ldg(row + q); f[4 * q] = v.x; f[4 * q + 1] = v.y; f[4 * q + 2] = v.z; f[4 * q + 3] = v.w; }
        store_row_split(sB, b_block_bytes, P.m_pad * 16, p, f);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long t_setup = P.timing ? clock64() : 0;

    if (warp < 2) {
        // ===== producers: object rows -> bf16 h/m/l operand blocks (two rows per thread) =====
        uint32_t it = 0;
        long long tw = 0, tk = 0;
        for (uint64_t rb = blockIdx.x; rb < n_rb; rb += gridDim.x, it++) {
            const uint32_t s = it % kStages, ph = (it / kStages) & 1;
            // request this row block's object rows BEFORE waiting for the stage: the DRAM latency hides behind the wait
            float f[2][16];
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const uint64_t row = rb * kRows + threadIdx.x + half * kProducerThreads;
                if (row < P.n) {
                    const float4 *src = reinterpret_cast<const float4 *>(P.fobj + row * 16);
#pragma unroll
                    for (int q = 0; q < 4; q++) { const float4 v = __ldg(src + q); f[half][4 * q] = v.x; f[half][4 * q + 1] = v.y; f[half][4 * q + 2] = v.z; f[half][4 * q + 3] = v.w; }
                } else {
#pragma unroll
                    for (int k = 0; k < 16; k++) f[half][k] = 0.f;
                }
            }
            const long long c0 = P.timing ? clock64() : 0;
            mbar_wait(&a_empty[s], ph ^ 1);
            const long long c1 = P.timing ? clock64() : 0;
#pragma unroll
            for (int half = 0; half < 2; half++)
                store_row_split(sA + s * kAStageBytes, kABlockBytes, kRows * 16, threadIdx.x + half * kProducerThreads, f[half]);
            fence_proxy_async();             // generic-proxy stores -> visible to the tensor core (async proxy)
            mbar_arrive(&a_full[s]);
            if (P.timing) { tw += c1 - c0; tk += clock64() - c1; }
        }
        if (P.timing && threadIdx.x == 0) { P.timing[blockIdx.x * 16 + 0] = tw; P.timing[blockIdx.x * 16 + 1] = tk; }
    } else if (warp == 2) {
        // ===== MMA issuer (one lane) =====
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
            // (A term, B term) in increasing magnitude: hl, lh, mm, hm, mh, hh   (0 = h, 1 = m, 2 = l)
            const int ta[6] = {0, 2, 1, 0, 1, 0}, tb[6] = {2, 0, 1, 1, 0, 0};
            uint32_t it = 0, g = 0;
            long long twa = 0, twt = 0, tis = 0;
            for (uint64_t rb = blockIdx.x; rb < n_rb; rb += gridDim.x, it++) {
                const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                const long long c0 = P.timing ? clock64() : 0;
                mbar_wait(&a_full[s], ph);
                if (P.timing) twa += clock64() - c0;
                tc_fence_after();
                const uint32_t a_addr = smem_u32(sA + s * kAStageBytes);
                for (uint32_t t = 0; t < n_tiles; t++, g++) {
                    const uint32_t buf = g % NBUF, pht = (g / NBUF) & 1;
                    const long long c1 = P.timing ? clock64() : 0;
                    mbar_wait(&t_empty[buf], pht ^ 1);
                    const long long c2 = P.timing ? clock64() : 0;
                    tc_fence_after();
                    const uint32_t d = tmem_base + buf * NT;
#pragma unroll
                    for (int q = 0; q < 6; q++) {
                        const uint64_t da = make_desc(a_addr + ta[q] * kABlockBytes, P.lbo_a, P.sbo_a);
                        const uint64_t db = make_desc(smem_u32(sB) + tb[q] * b_block_bytes + t * NT * 16, P.lbo_b, P.sbo_b);
                        umma_bf16(d, da, db, idesc, q > 0 ? 1u : 0u);
                    }
                    umma_commit(&t_full[buf]);          // accumulator ready (implies fence::before_thread_sync)
                    if (P.timing) { twt += c2 - c1; tis += clock64() - c2; }
                }
                umma_commit(&a_empty[s]);               // operand stage free once every MMA above has read it
            }
            if (P.timing) { P.timing[blockIdx.x * 16 + 2] = twa; P.timing[blockIdx.x * 16 + 3] = twt; P.timing[blockIdx.x * 16 + 4] = tis; }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===== epilogue warps 4..11: TMEM -> registers -> running (best value, best group of 8 columns) =====
        // A warp's tcgen05.ld occupies its own issue slot until the data has landed (measured: load time and reduce time
        // ADD within one warp, profiles/r01_umma_epilogue_experiments.txt), so two warps share every TMEM lane quarter:
        // warps 4-7 ("half A") own the lower NT/2 columns of each tile, warps 8-11 ("half B") the upper NT/2, and while one
        // loads the other reduces on the same SM sub-partition.  Half B hands its (best, group) to half A through shared memory.
        const uint32_t q = warp & 3;                    // TMEM lane quarter this warp may access
        const uint32_t half = (warp - 4) >> 2;          // 0 = columns [0, NT/2), 1 = columns [NT/2, NT)
        const uint32_t lane_base = (q * 32) << 16;
        constexpr int kChunks = NT / 64;                // 32-column chunks per half tile
        constexpr int kBatch = kChunks < LDW ? kChunks : LDW;
        uint32_t it = 0, g = 0;
        long long twf = 0, tld = 0, trs = 0;
        for (uint64_t rb = blockIdx.x; rb < n_rb; rb += gridDim.x, it++) {
            float best = -INFINITY;
            uint32_t bgroup = 0;
            const uint32_t r = q * 32 + lane;           // row inside the row block == TMEM lane
            const uint64_t row = rb * kRows + r;
            for (uint32_t t = 0; t < n_tiles; t++, g++) {
                const uint32_t buf = g % NBUF, pht = (g / NBUF) & 1;
                const long long c0 = P.timing ? clock64() : 0;
                mbar_wait(&t_full[buf], pht);
                const long long c1 = P.timing ? clock64() : 0;
                tc_fence_after();
                const uint32_t tcol = tmem_base + lane_base + buf * NT + half * (NT / 2);
#pragma unroll
                for (int c0i = 0; c0i < kChunks; c0i += kBatch) {
                    uint32_t v[kBatch][32];
#pragma unroll
                    for (int w = 0; w < kBatch; w++) tmem_ld32(tcol + (c0i + w) * 32, v[w]);
                    tmem_wait_ld();
#pragma unroll
                    for (int w = 0; w < kBatch; w++) {
                        uint32_t (&x)[32] = v[w];
                        const uint32_t col0 = t * NT + half * (NT / 2) + (c0i + w) * 32;
                        if (col0 + 32 > P.n_live) {          // warp-uniform: only the padded tail of the last tile
#pragma unroll
                            for (int i = 0; i < 32; i++) if (col0 + i >= P.n_live) x[i] = 0xFF800000u;   // -inf
                        }
#pragma unroll
                        for (int gq = 0; gq < 4; gq++) {
                            const float *f = reinterpret_cast<const float *>(&x[gq * 8]);
                            const float gm = fmaxf(max3f(max3f(f[0], f[1], f[2]), f[3], f[4]), max3f(f[5], f[6], f[7]));
                            if (gm > best) { best = gm; bgroup = (col0 >> 3) + gq; }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&t_empty[buf]);
                if (P.timing) { twf += c1 - c0; tld += clock64() - c1; }
            }
            const long long c2 = P.timing ? clock64() : 0;
            // merge the two column halves: within a tile half A's columns precede half B's, and both walk the tiles in
            // order, so on equal values the smaller group index is the earlier column
            float2 *slot = sMerge + (it & 1) * kRows;
            if (half == 1) slot[r] = make_float2(best, __uint_as_float(bgroup));
            asm volatile("bar.sync 1, 256;" ::: "memory");     // the 8 epilogue warps only
            if (half == 0) {
                const float2 o = slot[r];
                const uint32_t og = __float_as_uint(o.y);
                if (o.x > best || (o.x == best && og < bgroup)) bgroup = og;
                // The candidate re-evaluation is NOT done here: while the tensor core streams its K=16 operands out of shared
                // memory the L1/shared datapath is saturated and every LDS/LDG of an epilogue warp takes hundreds of cycles
                // (profiles/r01_umma_role_cycles_*.txt).  Only the winning group of 8 columns leaves this kernel;
                // k_affinity_resolve turns it into (node index, exact fp32 cost) in a second pass.
                if (row < P.n) P.out_idx[row] = bgroup;
            }
            if (P.timing) trs += clock64() - c2;
        }
        if (P.timing && threadIdx.x == 128) { P.timing[blockIdx.x * 16 + 5] = twf; P.timing[blockIdx.x * 16 + 6] = tld; P.timing[blockIdx.x * 16 + 7] = trs; }
    }

    // ---- teardown -------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (P.timing && threadIdx.x == 0) { P.timing[blockIdx.x * 16 + 8] = t_setup - t_begin; P.timing[blockIdx.x * 16 + 9] = clock64() - t_begin; P.timing[blockIdx.x * 16 + 10] = t_begin; }
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(NBUF * NT)) : "memory");
    }
}

// Second pass: the 8 candidates of each object's winning group, re-evaluated in fp32 with the summation order of
// k_assign_affinity (fmaf over k = 0..15).  EIGHT lanes per object, four objects per warp trip: lane 8q + r scores
// candidate r of object q.  `fnode_g` is the node features regrouped on the host as [group][16-byte piece k][candidate r],
// so the k-th load of the eight lanes of an object is one 128-byte line (4 lines per object, like the contiguous rows),
// and the object's own row is a broadcast inside the 8 lanes.  Three xor-shuffle steps pick the smallest
// (cost, position).  The earlier one-warp-per-object version issued 50 warp instructions per object and was
// issue-bound at 0.79 ms for 10 M objects (profiles/r01_launches_final.csv); this one issues ~16.
// in/out: idx[row] holds the group on entry, the interned node index on exit.
__global__ void __launch_bounds__(256)
k_affinity_resolve(const float *__restrict__ fobj, uint64_t n, const float *__restrict__ fnode_g, const uint32_t *__restrict__ nidx_map, uint32_t n_live,
                   uint32_t *__restrict__ idx, float *__restrict__ out_cost, uint32_t *__restrict__ counters, uint32_t hist_bins) {
    extern __shared__ uint32_t shist[];
    for (uint32_t j = threadIdx.x; j < hist_bins; j += blockDim.x) shist[j] = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, q = lane >> 3, r = lane & 7;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const float4 *fobj4 = reinterpret_cast<const float4 *>(fobj);
    const float4 *fnode4 = reinterpret_cast<const float4 *>(fnode_g) + r;     // piece k of candidate r of group g: + (4*g + k) * 8
    // Software pipeline, one trip ahead: the group index and the object's row of the NEXT trip are in flight (HBM latency)
    // while the current trip reads its candidates (L1/L2 hits) and reduces; without it a warp has one dependent chain
    // idx -> node rows -> dot product in flight and the kernel is latency bound (profiles/r01_ncu_affinity_resolve_v2.json).
    uint64_t base = warp0 * 4;
    bool mine = base + q < n;
    uint64_t row = mine ? base + q : n - 1;                                     // clamp: the tail recomputes the last object
    uint32_t g = 0;
    float4 o[4] = {};
    if (base < n) {
        g = __ldg(idx + row);
#pragma unroll
        for (int k = 0; k < 4; k++) o[k] = __ldg(fobj4 + row * 4 + k);
    }
    for (; base < n; base += nwarps * 4) {
        const uint64_t nbase = base + nwarps * 4;
        const bool nmine = nbase + q < n;
        const uint64_t nrow = nmine ? nbase + q : n - 1;
        uint32_t ng = 0;
        float4 no[4] = {};
        if (nbase < n) {                                                        // warp-uniform
            ng = __ldg(idx + nrow);     // safe to read ahead: idx[nrow] is rewritten only by the trip that owns nrow (this warp, later)
#pragma unroll
            for (int k = 0; k < 4; k++) no[k] = __ldg(fobj4 + nrow * 4 + k);
        }
        const float4 *fn = fnode4 + (size_t)g * 32;
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float4 x = __ldg(fn + k * 8);                                 // rows beyond n_live are zero padding
            a = fmaf(o[k].x, x.x, a); a = fmaf(o[k].y, x.y, a); a = fmaf(o[k].z, x.z, a); a = fmaf(o[k].w, x.w, a);
        }
        // order-preserving integer image of cost = -dot (+inf for padding); minimum of (key, position) over the 8 lanes
        uint32_t p = g * 8 + r;
        const uint32_t bits = p < n_live ? __float_as_uint(-a) : 0x7F800000u;
        int key = (int)(bits ^ ((uint32_t)((int)bits >> 31) & 0x7FFFFFFFu));
        if (p >= n_live) p = kNone;
#pragma unroll
        for (int d = 4; d >= 1; d >>= 1) {
            const int ok = __shfl_xor_sync(0xFFFFFFFFu, key, d);
            const uint32_t op = __shfl_xor_sync(0xFFFFFFFFu, p, d);
            if (ok < key || (ok == key && op < p)) { key = ok; p = op; }
        }
        if (r == 0 && mine) {
            const uint32_t nid = p == kNone ? kNone : __ldg(nidx_map + p);
            idx[row] = nid;
            if (out_cost) {
                const uint32_t kb = (uint32_t)key;
                out_cost[row] = p == kNone ? 0.f : __uint_as_float(kb ^ ((uint32_t)((int)kb >> 31) & 0x7FFFFFFFu));   // the image is an involution
            }
            if (nid != kNone) {
                if (hist_bins) atomicAdd(&shist[nid], 1u);
                else if (counters) atomicAdd(&counters[nid], 1u);
            }
        }
        mine = nmine; row = nrow; g = ng;
#pragma unroll
        for (int k = 0; k < 4; k++) o[k] = no[k];
    }
    if (hist_bins) {
        __syncthreads();
        if (counters)
            for (uint32_t j = threadIdx.x; j < hist_bins; j += blockDim.x) { const uint32_t v = shist[j]; if (v) atomicAdd(&counters[j], v); }
    }
}

}  // namespace

#define RIO_COUNT_LAUNCH(L) do { if ((L).launch_counter) ++*(L).launch_counter; } while (0)

// development hook (tools/umma_timing.py): device buffer of 16 u64 per CTA that receives per-role cycle counters
static unsigned long long *g_umma_timing = nullptr;
void affinity_umma_set_timing_buffer(unsigned long long *d) { g_umma_timing = d; }

// Largest padded live-node count whose operands fit in shared memory beside the A stages.
uint32_t affinity_umma_max_nodes() { return ((227u * 1024u - kBarBytes - kStages * kAStageBytes) / 96u) / 256u * 256u; }

bool launch_assign_affinity_umma(const Launch &L, const float *d_fobj, uint64_t n, const float *d_fnode_c, const float *d_fnode_g, const uint32_t *d_nidx_map, uint32_t n_live,
                                 uint32_t m_pad, uint32_t n_total, uint32_t *d_out_idx, float *d_out_cost, uint32_t *d_counters) {
    if (!n || !n_live) return false;
    const bool small = m_pad <= 64;
    if ((!small && (m_pad % 256)) || m_pad > affinity_umma_max_nodes()) return false;
    // K-major interleaved operands: LBO = distance between the two 16-byte K chunks, SBO = distance between 8-row groups
    // (confirmed on hardware: profiles/r01_umma_first_light.txt)
    const size_t smem = kBarBytes + (size_t)3 * m_pad * 32 + (size_t)kStages * kAStageBytes;
    UmmaParams P{d_fobj, n, d_fnode_c, d_nidx_map, n_live, m_pad, d_out_idx, d_out_cost, d_counters, kRows * 16u, 128u, m_pad * 16u, 128u, g_umma_timing};
    const uint64_t n_rb = (n + kRows - 1) / kRows;
    const int grid = (int)(n_rb < (uint64_t)L.sm_count ? n_rb : (uint64_t)L.sm_count);
    cudaError_t attr = cudaSuccess;
    if (small) {
        attr = cudaFuncSetAttribute(k_affinity_umma<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (attr == cudaSuccess) k_affinity_umma<64, 1><<<grid, kUmmaThreads, smem, L.stream>>>(P);
#ifdef RIO_ASSIGN_TUNING   // A/B points of the epilogue (RIO_BUILD_TUNING=1): not in the shipped library
    } else if (getenv("RIO_UMMA_LDW") && atoi(getenv("RIO_UMMA_LDW")) == 1) {   // one x32 TMEM load per double-buffer stage
        attr = cudaFuncSetAttribute(k_affinity_umma<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (attr == cudaSuccess) k_affinity_umma<256, 1><<<grid, kUmmaThreads, smem, L.stream>>>(P);
    } else if (getenv("RIO_UMMA_NT") && atoi(getenv("RIO_UMMA_NT")) == 128) {   // four 128-column accumulators
        attr = cudaFuncSetAttribute(k_affinity_umma<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (attr == cudaSuccess) k_affinity_umma<128, 2><<<grid, kUmmaThreads, smem, L.stream>>>(P);
#endif
    } else {
        attr = cudaFuncSetAttribute(k_affinity_umma<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (attr == cudaSuccess) k_affinity_umma<256, 2><<<grid, kUmmaThreads, smem, L.stream>>>(P);
    }
    // a launch that did not happen (shared memory over the limit, a device without tcgen05) must not leave the caller with stale
    // indices: report it, the engine then runs the CUDA-core kernel instead
    if (attr != cudaSuccess || cudaPeekAtLastError() != cudaSuccess) { cudaGetLastError(); return false; }
    RIO_COUNT_LAUNCH(L);
    {
        const uint32_t bins = (d_counters && n_total <= 8192) ? n_total : 0;
        const uint64_t blocks = (n + 31) / 32, cap = (uint64_t)L.sm_count * 8;   // 4 objects per warp trip, 8 warps per CTA
        k_affinity_resolve<<<(int)(blocks < cap ? blocks : cap), 256, (size_t)bins * 4, L.stream>>>(d_fobj, n, d_fnode_g, d_nidx_map, n_live, d_out_idx, d_out_cost, d_counters,
                                                                                          bins);
        RIO_COUNT_LAUNCH(L);
    }
    return true;
}

}  // namespace rio
