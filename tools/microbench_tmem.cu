// microbench_tmem.cu -- TMEM -> register read bandwidth of tcgen05.ld on sm_100a (development tool, not product).
// One CTA per SM allocates 512 TMEM columns; W warps per lane quarter stream the whole allocation REPS times with
// tcgen05.ld.32x32b.x{32,64,128}; reports bytes/clk/SM.  (TMEM contents are whatever they are: only bandwidth matters.)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X> __device__ __forceinline__ uint32_t ld(uint32_t taddr);
template <> __device__ __forceinline__ uint32_t ld<32>(uint32_t taddr) {
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
                   "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t a = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) a ^= v[i];
    return a;
}
// four x32 loads in flight before one wait (what the product epilogue does per double-buffer stage pair)
__device__ __forceinline__ uint32_t ld4(uint32_t taddr) {
    uint32_t v[4][32];
#pragma unroll
    for (int w = 0; w < 4; w++)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[w][0]),"=r"(v[w][1]),"=r"(v[w][2]),"=r"(v[w][3]),"=r"(v[w][4]),"=r"(v[w][5]),"=r"(v[w][6]),"=r"(v[w][7]),"=r"(v[w][8]),"=r"(v[w][9]),"=r"(v[w][10]),"=r"(v[w][11]),"=r"(v[w][12]),"=r"(v[w][13]),"=r"(v[w][14]),"=r"(v[w][15]),
                       "=r"(v[w][16]),"=r"(v[w][17]),"=r"(v[w][18]),"=r"(v[w][19]),"=r"(v[w][20]),"=r"(v[w][21]),"=r"(v[w][22]),"=r"(v[w][23]),"=r"(v[w][24]),"=r"(v[w][25]),"=r"(v[w][26]),"=r"(v[w][27]),"=r"(v[w][28]),"=r"(v[w][29]),"=r"(v[w][30]),"=r"(v[w][31])
                     : "r"(taddr + w * 32) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t a = 0;
#pragma unroll
    for (int w = 0; w < 4; w++)
#pragma unroll
        for (int i = 0; i < 32; i++) a ^= v[w][i];
    return a;
}

template <int MODE>   // 0: one x32 per wait, 1: four x32 per wait
__global__ void __launch_bounds__(256, 1) k(uint32_t reps, uint32_t warps, uint32_t *out) {
    __shared__ uint32_t slot;
    const uint32_t warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    uint32_t acc = 0;
    if (warp < warps) {
        const uint32_t lane_base = ((warp & 3) * 32) << 16;
        for (uint32_t r = 0; r < reps; r++) {
            if (MODE == 0) { for (uint32_t c = 0; c < 512; c += 32) acc ^= ld<32>(base + lane_base + c); }
            else           { for (uint32_t c = 0; c < 512; c += 128) acc ^= ld4(base + lane_base + c); }
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u) : "memory");
}

template <int MODE> int run(const char *name, uint32_t warps, int sms, uint32_t *d_out) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const uint32_t reps = 4000;
    k<MODE><<<sms, 256>>>(10, warps, d_out);
    CK(cudaEventRecord(a));
    k<MODE><<<sms, 256>>>(reps, warps, d_out);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    // every warp reads its 32-lane quarter of all 512 columns: 32 lanes x 512 cols x 4 B = 64 KB per rep per warp
    const double bytes = (double)sms * warps * reps * 65536.0;
    printf("%-28s warps=%u: %8.1f GB/s/SM  %6.1f B/clk/SM (at 1.965 GHz)  total %.2f TB/s\n", name, warps, bytes / (ms * 1e-3) / sms / 1e9,
           bytes / (ms * 1e-3) / sms / 1.965e9, bytes / (ms * 1e-3) / 1e12);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    uint32_t *d_out; CK(cudaMalloc(&d_out, 64));
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    for (uint32_t w : {1u, 4u, 8u}) run<0>("ld.32x32b.x32, wait each", w, p.multiProcessorCount, d_out);
    for (uint32_t w : {1u, 4u, 8u}) run<1>("4 x ld.x32 per wait", w, p.multiProcessorCount, d_out);
    return 0;
}
