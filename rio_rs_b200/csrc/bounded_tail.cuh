// bounded_tail.cuh -- the counter exchange and the bounded-load capacity check of one assignment pass as a
// block-level device routine.  It runs either as its own single-CTA kernel (k_exchange_check, flat rendezvous passes) or in the
// LAST CTA of the HRW2 walk kernel, so that a whole pass -- walk, histogram, exchange over NVLink peer memory, capacity check,
// two words to the host -- is ONE launch.
#pragma once
#include "kernels.cuh"
#include "spec.cuh"

namespace rio {

// Exchange (world > 1): every rank owns a window {slots[2][world][max_nodes] u32, flags[world] u32} that all peers have mapped
// through CUDA IPC.  push: my M counters go into slot [epoch&1][rank] of every peer's window (plain P2P stores over NVLink);
// signal: a release store of the epoch into flags[rank] of every peer; wait: spin (acquire loads, system scope) until my own
// window carries this epoch from every rank; sum.  Slots are double buffered by epoch parity: nobody can be two exchanges
// ahead, because every exchange needs everybody's flag.
// Check: over = live && count > cap, thr = floor(2^32 (count - cap) / count), closed (tagged with the call's epoch) |= over;
// {any over, live nodes still open} go to mapped pinned host memory.  next_zero (nullable): the counter buffer of the NEXT
// pass is cleared here, so the pass needs no memset launch.
// Must be called by every thread of the block.  `local` is read through L2 (other CTAs produced it with atomics).
__device__ __forceinline__ void exchange_and_check_block(const BoundedTail &b, const uint32_t *local) {
    __shared__ uint32_t s_any, s_open;
    if (threadIdx.x == 0) { s_any = 0; s_open = 0; }
    const uint32_t M = b.M, world = b.world, rank = b.rank, epoch = b.xchg_epoch;
    if (world > 1) {
        const size_t slot_words = (size_t)2 * world * b.max_nodes;
        const size_t par = (size_t)(epoch & 1u) * world * b.max_nodes;
        for (uint32_t p = 0; p < world; p++) {
            uint32_t *dst = b.peers.win[p] + par + (size_t)rank * b.max_nodes;
            for (uint32_t j = threadIdx.x; j < M; j += blockDim.x) dst[j] = __ldcg(local + j);
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < world) {
            uint32_t *flag = b.peers.win[threadIdx.x] + slot_words + rank;
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
            const uint32_t *mine = b.peers.win[rank] + slot_words + threadIdx.x;
            uint32_t v;
            // polite spin: a pipelined check shares its SM with five walk CTAs of the next set and may wait a whole kernel for the
            // slowest rank; sleeping between polls leaves that SM's issue slots to the walk (a tight acquire loop cost ~7 % at N = 4)
            for (;;) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
                if ((int32_t)(v - epoch) >= 0) break;
                __nanosleep(256);
            }
        }
        __syncthreads();
        const volatile uint32_t *src = b.peers.win[rank] + par;
        for (uint32_t j = threadIdx.x; j < M; j += blockDim.x) {
            uint32_t sum = 0;
            for (uint32_t r = 0; r < world; r++) sum += src[(size_t)r * b.max_nodes + j];
            b.glob[j] = sum;
        }
    } else {
        __syncthreads();
    }
    // the check: up to 4 nodes per thread per trip, every load of the trip issued before the first use (the block is alone on the
    // machine by now: latency, not bandwidth, is what this loop costs)
    const bool from_local = world <= 1;
    uint32_t my_any = 0, my_open = 0;
    for (uint32_t j0 = threadIdx.x; j0 < M; j0 += 4 * blockDim.x) {
        uint32_t c[4], cp[4], ce[4];
        uint8_t st[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const uint32_t j = j0 + e * blockDim.x;
            if (j < M) { c[e] = from_local ? __ldcg(local + j) : b.glob[j]; cp[e] = b.cap[j]; ce[e] = b.closed_epoch[j]; st[e] = b.state[j]; }
        }
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const uint32_t j = j0 + e * blockDim.x;
            if (j >= M) continue;
            const bool live = st[e] & kNodeLive;
            const bool ov = live && c[e] > cp[e];
            if (from_local && local != b.glob) b.glob[j] = c[e];
            b.over[j] = ov;
            b.thr[j] = ov ? (uint32_t)((((unsigned long long)(c[e] - cp[e])) << 32) / c[e]) : 0u;
            if (ov) b.closed_epoch[j] = b.call_epoch;
            my_any |= ov;
            my_open += live && !ov && ce[e] != b.call_epoch;
            if (b.next_zero) b.next_zero[j] = 0;
        }
    }
    if (my_any) atomicOr(&s_any, 1u);
    if (my_open) atomicAdd(&s_open, my_open);
    __syncthreads();
    if (threadIdx.x == 0) {
        // ONE 16-byte store carries {any over, open nodes, sequence number}: it crosses PCIe as a single write, so the host, which
        // polls the sequence word, never sees a torn report and the device needs no system-wide fence between the words
        *reinterpret_cast<uint4 *>(const_cast<uint32_t *>(b.host_flags)) = make_uint4(s_any, s_open, b.flag_seq, 0u);
        __threadfence_system();
    }
}

}  // namespace rio
