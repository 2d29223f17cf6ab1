// resolver.cu -- micro-batching front end for the per-request call sites.
//
// Besides get_or_create_placement, the trait's own per-id calls (lookup / update / remove: object_placement/mod.rs:46-55) go
// through the same queue: a micro-batch may hold all three kinds; the worker applies the updates first (array order, like the
// engine's batched upsert), then the lookups, then the resolves -- one batched engine call per kind that is present, so N
// concurrent callers share one GPU round trip instead of queueing N of them behind the handle's mutex.
//
// Service::get_or_create_placement (rio-rs/src/service.rs:193-254) runs once per request, on one tokio task per
// connection (rio-rs/src/server.rs:303).  A kernel launch per id would lose to the HashMap, so concurrent per-id calls
// are coalesced here: callers enqueue (key, slot) and block; one worker thread drains the queue into a single
// rio_cuda_place_batch (the same decisions, batched) as soon as either `max_batch` requests are pending or the oldest one
// has waited `max_wait_us`.  Built only on the public C ABI, so the Rust provider gets the same thing through FFI.
#include "../../include/rio_cuda.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

// Waiting is spin-then-block on both sides: a GPU round trip is ~20 us, a futex sleep/wake pair costs about as much again, so
// callers poll the batch's `done` flag for a while before they sleep, and the worker polls for new work for a while after a batch
// before it sleeps.  Under load nobody sleeps; an idle front end costs nothing.
inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
}
constexpr int kSpinUs = 60;

struct Batch {
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<int> done_flag{0};
    std::atomic<int> sleepers{0};
    bool done = false;
    rio_status status = RIO_OK;
    std::string error;
    std::vector<uint64_t> keys;        // get_or_create_placement requests
    std::vector<uint32_t> out;
    std::vector<uint64_t> lk_keys;     // lookup requests
    std::vector<uint32_t> lk_out;
    std::vector<uint64_t> up_keys;     // update / remove requests (idx RIO_NONE = update(None) = remove)
    std::vector<uint32_t> up_idx;
    size_t size() const { return keys.size() + lk_keys.size() + up_keys.size(); }
};

}  // namespace

struct rio_resolver {
    rio_placement *h = nullptr;
    uint32_t policy = RIO_PLACE_HRW, self_idx = 0, max_batch = 4096, max_wait_us = 50;
    std::mutex mu;
    std::condition_variable cv_work;
    std::atomic<uint32_t> pending{0};     // requests in the open batch (lock-free view for the polling worker)
    std::atomic<int> worker_sleeping{0};
    std::shared_ptr<Batch> open;          // batch currently collecting requests
    std::chrono::steady_clock::time_point open_since;
    bool stop = false;
    uint64_t calls = 0, batches = 0, max_seen = 0;
    std::thread worker;

    void run() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            if (!(open && open->size()) && !stop) {   // poll for a while before sleeping: the next request is usually microseconds away
                lk.unlock();
                const auto t_end = std::chrono::steady_clock::now() + std::chrono::microseconds(kSpinUs);
                while (pending.load(std::memory_order_acquire) == 0 && std::chrono::steady_clock::now() < t_end) cpu_relax();
                lk.lock();
            }
            if (!(open && open->size()) && !stop) {
                worker_sleeping.store(1);
                cv_work.wait(lk, [&] { return stop || (open && open->size()); });
                worker_sleeping.store(0);
            }
            if (stop && !(open && open->size())) return;
            // let the batch fill: until max_batch requests, until the oldest has waited max_wait_us, or until arrivals pause (a
            // lone caller must not pay the whole window)
            const auto deadline = open_since + std::chrono::microseconds(max_wait_us);
            lk.unlock();
            uint32_t seen = pending.load(std::memory_order_acquire);
            auto quiet_since = std::chrono::steady_clock::now();
            for (;;) {
                const auto now = std::chrono::steady_clock::now();
                const uint32_t p = pending.load(std::memory_order_acquire);
                if (p >= max_batch || now >= deadline) break;
                if (p != seen) { seen = p; quiet_since = now; }
                else if (now - quiet_since > std::chrono::microseconds(std::max<uint32_t>(3, max_wait_us / 4))) break;   // arrivals paused
                cpu_relax();
            }
            lk.lock();
            std::shared_ptr<Batch> b = std::move(open);
            open.reset();
            pending.store(0, std::memory_order_release);
            batches++;
            if (b->size() > max_seen) max_seen = b->size();
            lk.unlock();
            b->out.assign(b->keys.size(), RIO_NONE);
            b->lk_out.assign(b->lk_keys.size(), RIO_NONE);
            rio_status st = RIO_OK;
            if (!b->up_keys.empty()) st = rio_cuda_upsert_batch(h, b->up_keys.data(), b->up_idx.data(), b->up_keys.size());
            if (st == RIO_OK && !b->lk_keys.empty()) st = rio_cuda_lookup_batch(h, b->lk_keys.data(), b->lk_keys.size(), b->lk_out.data());
            if (st == RIO_OK && !b->keys.empty()) st = rio_cuda_place_batch(h, b->keys.data(), b->keys.size(), policy, self_idx, b->out.data());
            {
                std::lock_guard<std::mutex> g(b->mu);
                b->status = st;
                if (st != RIO_OK) { const char *m = rio_cuda_last_error(h); b->error = m ? m : ""; }
                b->done = true;
                b->done_flag.store(1, std::memory_order_release);
            }
            if (b->sleepers.load(std::memory_order_acquire)) b->cv.notify_all();
            lk.lock();
        }
    }
};

static thread_local std::string t_resolver_error;

extern "C" {


rio_status rio_cuda_resolver_create(rio_placement *h, uint32_t policy, uint32_t self_idx, uint32_t max_batch, uint32_t max_wait_us, rio_resolver **out) {
    if (!h || !out || (policy != RIO_PLACE_SELF && policy != RIO_PLACE_HRW && policy != RIO_PLACE_HRW2)) return RIO_ERR_UNKNOWN;
    rio_resolver *r = new rio_resolver();
    r->h = h; r->policy = policy; r->self_idx = self_idx;
    r->max_batch = max_batch ? max_batch : 4096;
    r->max_wait_us = max_wait_us;
    r->worker = std::thread([r] { r->run(); });
    *out = r;
    return RIO_OK;
}

void rio_cuda_resolver_destroy(rio_resolver *r) {
    if (!r) return;
    { std::lock_guard<std::mutex> g(r->mu); r->stop = true; }
    r->cv_work.notify_all();
    if (r->worker.joinable()) r->worker.join();
    delete r;
}

enum { OP_RESOLVE = 0, OP_LOOKUP = 1, OP_UPDATE = 2 };
static rio_status submit(rio_resolver *r, int op, uint64_t key, uint32_t idx_in, uint32_t *out_idx) {
    std::shared_ptr<Batch> b;
    size_t slot = 0;
    {
        std::lock_guard<std::mutex> g(r->mu);
        if (r->stop) return RIO_ERR_UNKNOWN;
        if (!r->open) { r->open = std::make_shared<Batch>(); r->open_since = std::chrono::steady_clock::now(); }
        b = r->open;
        if (op == OP_RESOLVE) { slot = b->keys.size(); b->keys.push_back(key); }
        else if (op == OP_LOOKUP) { slot = b->lk_keys.size(); b->lk_keys.push_back(key); }
        else { b->up_keys.push_back(key); b->up_idx.push_back(idx_in); }
        r->calls++;
        r->pending.fetch_add(1, std::memory_order_release);
    }
    if (r->worker_sleeping.load(std::memory_order_acquire)) r->cv_work.notify_one();
    {   // spin on the flag first; sleep only when the batch takes unusually long
        const auto t_end = std::chrono::steady_clock::now() + std::chrono::microseconds(4 * kSpinUs);
        while (!b->done_flag.load(std::memory_order_acquire) && std::chrono::steady_clock::now() < t_end) cpu_relax();
    }
    if (!b->done_flag.load(std::memory_order_acquire)) {
        std::unique_lock<std::mutex> lk(b->mu);
        b->sleepers.fetch_add(1);
        b->cv.wait(lk, [&] { return b->done; });
    }
    if (b->status != RIO_OK) { t_resolver_error = b->error; return b->status; }
    if (op == OP_RESOLVE) *out_idx = b->out[slot];
    else if (op == OP_LOOKUP) *out_idx = b->lk_out[slot];
    return RIO_OK;
}

rio_status rio_cuda_resolver_resolve(rio_resolver *r, uint64_t key, uint32_t *out_idx) {
    if (!r || !out_idx) return RIO_ERR_UNKNOWN;
    return submit(r, OP_RESOLVE, key, 0, out_idx);
}

/* ObjectPlacement::lookup / update / remove per id (mod.rs:46-55), coalesced with everybody else's */
rio_status rio_cuda_resolver_lookup(rio_resolver *r, uint64_t key, uint32_t *out_idx) {
    if (!r || !out_idx) return RIO_ERR_UNKNOWN;
    return submit(r, OP_LOOKUP, key, 0, out_idx);
}
rio_status rio_cuda_resolver_update(rio_resolver *r, uint64_t key, uint32_t idx) {
    if (!r) return RIO_ERR_UNKNOWN;
    return submit(r, OP_UPDATE, key, idx, nullptr);
}
rio_status rio_cuda_resolver_lookup_str(rio_resolver *r, const char *type, size_t type_len, const char *id, size_t id_len, char *buf, size_t cap,
                                        size_t *out_len) {
    if (!r || !type || !id || !out_len) return RIO_ERR_UNKNOWN;
    uint32_t idx = RIO_NONE;
    rio_status st = submit(r, OP_LOOKUP, rio_cuda_object_key(type, type_len, id, id_len), 0, &idx);
    if (st != RIO_OK) return st;
    if (idx == RIO_NONE) { *out_len = (size_t)-1; return RIO_OK; }
    return rio_cuda_node_address(r->h, idx, buf, cap, out_len);
}
rio_status rio_cuda_resolver_update_str(rio_resolver *r, const char *type, size_t type_len, const char *id, size_t id_len, const char *address,
                                        size_t address_len) {
    if (!r || !type || !id) return RIO_ERR_UNKNOWN;
    uint32_t idx = RIO_NONE;
    if (address) {   // any address may be recorded, live or not (local.rs:34-36): interned without touching liveness
        const std::string a(address, address_len);
        rio_status st = rio_cuda_node_intern(r->h, a.c_str(), &idx);
        if (st != RIO_OK) { const char *m = rio_cuda_last_error(r->h); t_resolver_error = m ? m : ""; return st; }
    }
    return submit(r, OP_UPDATE, rio_cuda_object_key(type, type_len, id, id_len), idx, nullptr);
}

/* get_or_create_placement(type, id) -> address string, exactly the per-request signature of service.rs:193-197 */
rio_status rio_cuda_resolver_resolve_str(rio_resolver *r, const char *type, size_t type_len, const char *id, size_t id_len, char *buf, size_t cap,
                                         size_t *out_len) {
    if (!r || !type || !id || !out_len) return RIO_ERR_UNKNOWN;
    uint32_t idx = RIO_NONE;
    rio_status st = rio_cuda_resolver_resolve(r, rio_cuda_object_key(type, type_len, id, id_len), &idx);
    if (st != RIO_OK) return st;
    if (idx == RIO_NONE) { *out_len = (size_t)-1; return RIO_OK; }
    return rio_cuda_node_address(r->h, idx, buf, cap, out_len);
}

rio_status rio_cuda_resolver_stats(rio_resolver *r, uint64_t *calls, uint64_t *batches, uint64_t *largest_batch) {
    if (!r) return RIO_ERR_UNKNOWN;
    std::lock_guard<std::mutex> g(r->mu);
    if (calls) *calls = r->calls;
    if (batches) *batches = r->batches;
    if (largest_batch) *largest_batch = r->max_seen;
    return RIO_OK;
}

const char *rio_cuda_resolver_last_error(void) { return t_resolver_error.c_str(); }

}  // extern "C"
