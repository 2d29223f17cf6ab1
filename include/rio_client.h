/* rio_client.h -- C ABI of librio_client.so: the CLIENT-side deterministic first hop.
 *
 * Replaces, on the client, the uniform-random pick of
 *     Client::get_service_object_address            rio-rs/src/client/mod.rs:235-267  (random choice :254-263)
 * by the same weighted rendezvous hash the servers' placement solver uses, so that the first
 * request for an object the cluster placed with RIO_PLACE_HRW goes to its owner and no Redirect round trip
 * (rio-rs/src/service.rs:261-298, protocol.rs ResponseError::Redirect) is needed.  With the reference's random pick the
 * expected redirect rate is (M-1)/M.
 *
 * This is NOT a CPU path of the server-side product: clients have no GPU, resolve one id at a time, and the servers
 * (librio_cuda.so, include/rio_cuda.h) never call into this library.  It shares the spec functions
 * (rio_rs_b200/csrc/spec.cuh) with the CUDA code and is checked bit for bit against the oracle in
 * tests/test_client_first_hop.py.
 *
 * The `ring` is the client's view of MembershipStorage::active_members (cluster/storage/mod.rs:95-99) -- what
 * Client::fetch_active_servers already keeps in `active_servers` (client/mod.rs:139-160).  Ties on (score, u)
 * (probability ~2^-32 per node pair) go to the lower position in `addresses`; pass the addresses in the servers'
 * interning order (rio_cuda_set_nodes order) if even those must coincide.
 */
#ifndef RIO_CLIENT_H
#define RIO_CLIENT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RIO_CLIENT_OK 0
#define RIO_CLIENT_ERR (-2)              /* bad argument (maps to ClientError::Unknown) */
#define RIO_CLIENT_NONE 0xFFFFFFFFu      /* no live server: ClientError::NoServersAvailable (client/mod.rs:260-261) */

typedef struct rio_client_ring rio_client_ring;

/* weights == NULL: all 1 (the reference has no weights); weight 0 = not live.  Strings are copied. */
int32_t rio_client_ring_create(const char *const *addresses, const size_t *address_lens, const uint32_t *weights, uint32_t n,
                               rio_client_ring **out);
void rio_client_ring_destroy(rio_client_ring *ring);
uint32_t rio_client_ring_size(const rio_client_ring *ring);
/* copies at most cap bytes, *out_len = full length */
int32_t rio_client_ring_address(const rio_client_ring *ring, uint32_t index, char *buf, size_t cap, size_t *out_len);

/* Which of the servers' solver policies the first hop mirrors (include/rio_cuda.h RIO_SOLVER_*): the flat weighted rendezvous
 * (default) or HRW2, the hierarchical one (trie_bits must equal the servers', 0 = 12).  Under HRW2 the pick
 * does not depend on the order of `addresses` at all (positions and chains are ordered by a hash of the address). */
#define RIO_CLIENT_POLICY_HRW  1u
#define RIO_CLIENT_POLICY_HRW2 2u
int32_t rio_client_ring_set_policy(rio_client_ring *ring, uint32_t policy, uint32_t trie_bits);

/* == rio_cuda_object_key: hash of the bytes of format!("{}.{}", type, id)  (object_placement/local.rs:26-29) */
uint64_t rio_client_object_key(const char *type, size_t type_len, const char *id, size_t id_len);

/* the owner under the weighted rendezvous hash: position in `addresses`, or RIO_CLIENT_NONE */
int32_t rio_client_first_hop(const rio_client_ring *ring, const char *type, size_t type_len, const char *id, size_t id_len, uint32_t *out_index);
int32_t rio_client_first_hop_key(const rio_client_ring *ring, uint64_t key, uint32_t *out_index);
int32_t rio_client_first_hop_batch(const rio_client_ring *ring, const uint64_t *keys, size_t n, uint32_t *out_index);

#ifdef __cplusplus
}
#endif
#endif
