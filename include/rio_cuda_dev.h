/*
 * rio_cuda_dev.h -- development, measurement and test hooks of librio_cuda.so.  NOT part of the provider ABI: a
 * rio-cuda-sys crate binds include/rio_cuda.h only.  bench.py, tools/ and the parity tests use these to time kernels on the
 * engine's own stream, to generate the synthetic key stream in HBM, and to force inputs the public API cannot produce
 * (duplicated node seeds: exact score ties).
 */
#ifndef RIO_CUDA_DEV_H
#define RIO_CUDA_DEV_H

#include "rio_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* key[i] = mix64(GOLDEN*(first+i+1) ^ seed): the synthetic stream of SURVEY 8d, generated in HBM */
rio_status  rio_cuda_set_synth_keys(rio_objset *s, uint64_t first, uint64_t n, uint64_t seed);
/* write a scratch buffer larger than L2 (bench hygiene between steps) */
rio_status  rio_cuda_flush_l2(rio_placement *h);

/* ---- timing on the handle's stream (CUDA events; torch.cuda.Event cannot see this stream) --------------- */
#define RIO_MAX_EVENTS 64
rio_status  rio_cuda_event_record(rio_placement *h, uint32_t slot);
rio_status  rio_cuda_event_elapsed_ms(rio_placement *h, uint32_t slot_start, uint32_t slot_end, float *out_ms);
/* Integer-ALU roofline probe: a register-only replay of the flat rendezvous inner loop (same instruction mix, no memory
 * traffic); reports (object,node) pair hashes per second.  Event slots RIO_MAX_EVENTS-2/-1 are used internally. */
rio_status  rio_cuda_bench_mix_rate(rio_placement *h, uint32_t iters, double *out_pairs_per_s);
/* number of kernels this handle has launched since creation (bench.py's gpu_launches) */
rio_status  rio_cuda_launch_count(rio_placement *h, uint64_t *out);

/* ---- test hooks: inputs the public API cannot produce ------------------------------------------------------ */
/* Override the seed of an interned node (normally mix64(fnv1a64(address))).  Two nodes with the same seed hash every
 * object alike: exact ties in u, and with equal weights in the 64-bit score (DESIGN.md 3.4 tie rules). */
rio_status  rio_dev_set_node_seed(rio_placement *h, uint32_t idx, uint64_t seed);
/* bit 0: build the class-sorted node table with ONE CLASS PER NODE (equal weights no longer merge into a class), so that
 * equal scores meet on the between-class path of the kernels instead of inside a class. */
#define RIO_DEV_SPLIT_CLASSES 1u
rio_status  rio_dev_set_table_options(rio_placement *h, uint32_t flags);
/* per-role cycle counters of the tcgen05 affinity kernel */
rio_status  rio_dev_umma_timing(rio_placement *h, unsigned long long *d_buf);

#ifdef __cplusplus
}
#endif
#endif /* RIO_CUDA_DEV_H */
