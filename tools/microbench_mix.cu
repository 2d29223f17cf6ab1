// microbench_mix.cu -- what bounds the rendezvous inner loop (2 IMAD + 1/2 VIMNMX3 per pair)?  Development tool.
// Variants of the same register-resident loop: node constants from registers or from broadcast LDS.128, 2- or
// 3-input max, 4 or 8 objects per thread, and the resident warp count limited through dynamic shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_mix tools/microbench_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// SRC: 0 = node constants advance in registers (2 IMAD per 2 nodes), 1 = 2 x LDS.128 per 2 nodes (broadcast)
// MAXOP: 0 = none (xor-accumulate via LOP3), 3 = VIMNMX3 per 2 pairs, 2 = VIMNMX per pair
template <int OPT, int SRC, int MAXOP>
__global__ void __launch_bounds__(256) mix(uint32_t iters, uint32_t nodes, const uint32_t *in, uint32_t *out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4 *srec = reinterpret_cast<uint4 *>(smem_raw);
    for (uint32_t j = threadIdx.x; j < nodes; j += blockDim.x) srec[j] = make_uint4(in[j & 1023], j, in[(j + 7) & 1023] | 1u, in[(j + 13) & 1023]);
    __syncthreads();
    uint32_t b[OPT], ab[OPT], gm[OPT];
#pragma unroll
    for (int k = 0; k < OPT; k++) { b[k] = in[(threadIdx.x * 8 + k) & 1023] | 1u; ab[k] = in[(threadIdx.x * 8 + 4 + k + 300) & 1023]; gm[k] = 0; }
    uint32_t s0a = in[(threadIdx.x + 64) & 1023] + blockIdx.x, s0b = s0a ^ 0x7F4A7C15u;
    const uint32_t ma = in[(threadIdx.x + 1) & 1023] | 1u, mb = in[(threadIdx.x + 2) & 1023] | 1u;
    const uint32_t s2a = in[(threadIdx.x + 3) & 1023], s2b = in[(threadIdx.x + 4) & 1023];
    for (uint32_t it = 0; it < iters; it++) {
        const uint4 *s = srec + ((it * 16) & (nodes - 1));     // nodes is a power of two >= 16
#pragma unroll
        for (int r = 0; r < 8; r++) {
            uint4 r0, r1;
            if (SRC == 1) { r0 = s[2 * r]; r1 = s[2 * r + 1]; }
            else { r0 = make_uint4(s0a, 0, ma, s2a); r1 = make_uint4(s0b, 0, mb, s2b); }
#pragma unroll
            for (int k = 0; k < OPT; k++) {
                const uint32_t u0 = (r0.x * b[k] + ab[k]) * r0.z + r0.w, u1 = (r1.x * b[k] + ab[k]) * r1.z + r1.w;
                if (MAXOP == 3) gm[k] = __vimax3_u32(gm[k], u0, u1);
                if (MAXOP == 2) gm[k] = max(max(gm[k], u0), u1);
                if (MAXOP == 0) gm[k] ^= u0 ^ u1;
            }
            if (SRC == 0) { s0a = s0a * 747796405u + 2891336453u; s0b = s0b * 1664525u + 1013904223u; }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < OPT; k++) acc ^= gm[k];
    if (acc == 0x12345678u) out[0] = acc;
}

template <int OPT, int SRC, int MAXOP>
int run(const char *name, int ctas_per_sm, const uint32_t *d_in, uint32_t *d_out, int sms) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    // limit residency with dynamic shared memory: 227 KB / ctas_per_sm each (the table itself needs 16 KB)
    size_t smem = (size_t)(227 * 1024 / ctas_per_sm - 1024) & ~(size_t)1023;
    if (smem < 16384) smem = 16384;
    CK(cudaFuncSetAttribute(mix<OPT, SRC, MAXOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mix<OPT, SRC, MAXOP>, 256, smem));
    const int grid = sms * occ; const uint32_t iters = 20000u * 4 / OPT;
    mix<OPT, SRC, MAXOP><<<grid, 256, smem>>>(100, 1024, d_in, d_out);
    CK(cudaEventRecord(a));
    mix<OPT, SRC, MAXOP><<<grid, 256, smem>>>(iters, 1024, d_in, d_out);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    const double pairs = (double)grid * 256 * iters * 16 * OPT;
    printf("%-58s %2d warps/SM  %6.2f Tpairs/s  %5.2f pairs/clk/SM\n", name, occ * 8, pairs / (ms * 1e-3) / 1e12, pairs / (ms * 1e-3) / sms / 1.965e9);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("%s, %d SMs; ideal at 2 IMAD/pair on a 64 lane/clk pipe = 32 pairs/clk/SM\n", p.name, p.multiProcessorCount);
    uint32_t h[1024]; for (int i = 0; i < 1024; i++) h[i] = 2654435761u * (i + 1) ^ (i << 7);
    uint32_t *d_in, *d_out; CK(cudaMalloc(&d_in, 4096)); CK(cudaMalloc(&d_out, 64)); CK(cudaMemcpy(d_in, h, 4096, cudaMemcpyHostToDevice));
    const int s = p.multiProcessorCount;
    for (int c : {8, 4, 3, 2}) {
        run<4, 0, 3>("4 obj, regs,  VIMNMX3", c, d_in, d_out, s);
        run<4, 1, 3>("4 obj, LDS,   VIMNMX3", c, d_in, d_out, s);
    }
    run<4, 1, 2>("4 obj, LDS,   2 x VIMNMX", 3, d_in, d_out, s);
    run<4, 1, 0>("4 obj, LDS,   xor (LOP3) instead of max", 3, d_in, d_out, s);
    run<4, 0, 0>("4 obj, regs,  xor (LOP3) instead of max", 3, d_in, d_out, s);
    run<8, 1, 3>("8 obj, LDS,   VIMNMX3", 3, d_in, d_out, s);
    run<8, 1, 3>("8 obj, LDS,   VIMNMX3", 2, d_in, d_out, s);
    run<8, 0, 3>("8 obj, regs,  VIMNMX3", 2, d_in, d_out, s);
    run<2, 1, 3>("2 obj, LDS,   VIMNMX3", 4, d_in, d_out, s);
    run<2, 1, 3>("2 obj, LDS,   VIMNMX3", 8, d_in, d_out, s);
    return 0;
}
