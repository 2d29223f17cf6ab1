// microbench_int.cu -- integer-pipe throughput probes for sm_100a (development tool, not product).
// Each kernel runs ITER iterations of an unrolled body on 8 independent chains per thread, full occupancy,
// and reports giga lane-ops/s per instruction class plus candidate pair-hash mixes (pairs/s).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int CH = 8;

template <int MODE>
__global__ void __launch_bounds__(256, 2) probe(uint32_t iters, const uint32_t *in, uint32_t *out) {
    uint32_t x[CH], y[CH];
    uint64_t t[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) { x[k] = in[(threadIdx.x * CH + k) & 1023]; y[k] = in[(threadIdx.x * CH + k + 512) & 1023] | 1u; t[k] = ((uint64_t)x[k] << 32) | y[k]; }
    uint32_t c = in[1000 + (blockIdx.x & 7)] | 1u, d = in[900 + (blockIdx.x & 7)];
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int k = 0; k < CH; k++) {
                if (MODE == 0) x[k] = x[k] * y[k] + d;                                    // IMAD
                if (MODE == 1) t[k] = (uint64_t)(uint32_t)t[k] * c + t[k];                // IMAD.WIDE.U32 (64-bit addend)
                if (MODE == 2) x[k] = __umulhi(x[k], y[k]) + d;                           // IMAD.HI
                if (MODE == 3) x[k] = x[k] ^ y[k] ^ d, y[k] = y[k] ^ c ^ x[k];            // LOP3 x2
                if (MODE == 4) x[k] = x[k] ^ (x[k] >> 15);                                // SHF + LOP3
                if (MODE == 5) x[k] = __vimax3_u32(x[k], y[k] + r, d), y[k] ^= x[k];      // VIMNMX3 + IADD + LOP3
                if (MODE == 6) { uint32_t p = x[k] * y[k] + d; uint64_t w = (uint64_t)p * 0x9E3779B1u + t[k]; x[k] = (uint32_t)w ^ (uint32_t)(w >> 32); }   // spec v1 pair hash (dependent chain through x)
                if (MODE == 7) { uint32_t p = x[k] * y[k] + d; p ^= p >> 15; x[k] = p * 0x9E3779B1u + c; }                                                   // 2xIMAD + xorshift
                if (MODE == 8) { uint64_t w = (uint64_t)x[k] * y[k] + t[k]; x[k] = (uint32_t)w ^ (uint32_t)(w >> 32) ^ d; }                                  // single wide + fold
                if (MODE == 9) x[k] = x[k] + y[k] + d;                                    // IADD3
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < CH; k++) acc ^= x[k] ^ y[k] ^ (uint32_t)t[k] ^ (uint32_t)(t[k] >> 32);
    if (acc == 0x12345678u) out[0] = acc;
}

template <int MODE>
int run(const char *name, double ops_per_body, const uint32_t *d_in, uint32_t *d_out, int sms) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int grid = sms * 2 * 4; const uint32_t iters = 4000;
    probe<MODE><<<grid, 256>>>(100, d_in, d_out);
    CK(cudaEventRecord(a));
    probe<MODE><<<grid, 256>>>(iters, d_in, d_out);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    double bodies = (double)grid * 256 * iters * 8 * CH;
    double per_clk_sm = bodies / (ms * 1e-3) / sms / 1.965e9;
    printf("%-34s %8.1f G bodies/s  %6.2f bodies/clk/SM  (%.2f lane-ops/clk/SM at %.1f ops/body)\n", name, bodies / (ms * 1e-3) / 1e9, per_clk_sm, per_clk_sm * ops_per_body, ops_per_body);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    uint32_t h[1024]; for (int i = 0; i < 1024; i++) h[i] = 2654435761u * (i + 1) ^ (i << 7);
    uint32_t *d_in, *d_out; CK(cudaMalloc(&d_in, 4096)); CK(cudaMalloc(&d_out, 64)); CK(cudaMemcpy(d_in, h, 4096, cudaMemcpyHostToDevice));
    int s = p.multiProcessorCount;
    run<0>("IMAD (x*y+d)", 1, d_in, d_out, s);
    run<1>("IMAD.WIDE.U32 (+64b addend)", 1, d_in, d_out, s);
    run<2>("IMAD.HI", 1, d_in, d_out, s);
    run<3>("LOP3 x2", 2, d_in, d_out, s);
    run<4>("SHF+LOP3 (xorshift)", 2, d_in, d_out, s);
    run<5>("VIMNMX3+IADD+LOP3", 3, d_in, d_out, s);
    run<9>("IADD3", 1, d_in, d_out, s);
    run<6>("hash v1: IMAD,WIDE,LOP3", 3, d_in, d_out, s);
    run<7>("hash B: IMAD,SHF,LOP3,IMAD", 4, d_in, d_out, s);
    run<8>("hash C: WIDE,LOP3(3in)", 2, d_in, d_out, s);
    return 0;
}
