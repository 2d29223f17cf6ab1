// resolver.cu -- micro-batching front end for the per-request call site.
//
// Service::get_or_create_placement (rio-rs/src/service.rs:193-254) runs once per request, on one tokio task per
// connection (rio-rs/src/server.rs:303).  A kernel launch per id would lose to the HashMap, so concurrent per-id calls
// are coalesced here: callers enqueue (key, slot) and block; one worker thread drains the queue into a single
// rio_cuda_place_batch (the same decisions, batched) as soon as either `max_batch` requests are pending or the oldest one
// has waited `max_wait_us`.  Built only on the public C ABI, so the Rust provider gets the same thing through FFI.
#include "../../include/rio_cuda.h"

#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Batch {
    std::mutex mu;
    std::condition_variable cv;
    bool done = false;
    rio_status status = RIO_OK;
    std::string error;
    std::vector<uint64_t> keys;
    std::vector<uint32_t> out;
};

}  // namespace

struct rio_resolver {
    rio_placement *h = nullptr;
    uint32_t policy = RIO_PLACE_HRW, self_idx = 0, max_batch = 4096, max_wait_us = 50;
    std::mutex mu;
    std::condition_variable cv_work;
    std::shared_ptr<Batch> open;          // batch currently collecting requests
    std::chrono::steady_clock::time_point open_since;
    bool stop = false;
    uint64_t calls = 0, batches = 0, max_seen = 0;
    std::thread worker;

    void run() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_work.wait(lk, [&] { return stop || (open && !open->keys.empty()); });
            if (stop && !(open && !open->keys.empty())) return;
            // let the batch fill: until max_batch requests or until the oldest has waited max_wait_us
            const auto deadline = open_since + std::chrono::microseconds(max_wait_us);
            cv_work.wait_until(lk, deadline, [&] { return stop || open->keys.size() >= max_batch; });
            std::shared_ptr<Batch> b = std::move(open);
            open.reset();
            batches++;
            if (b->keys.size() > max_seen) max_seen = b->keys.size();
            lk.unlock();
            b->out.assign(b->keys.size(), RIO_NONE);
            rio_status st = rio_cuda_place_batch(h, b->keys.data(), b->keys.size(), policy, self_idx, b->out.data());
            {
                std::lock_guard<std::mutex> g(b->mu);
                b->status = st;
                if (st != RIO_OK) { const char *m = rio_cuda_last_error(h); b->error = m ? m : ""; }
                b->done = true;
            }
            b->cv.notify_all();
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

rio_status rio_cuda_resolver_resolve(rio_resolver *r, uint64_t key, uint32_t *out_idx) {
    if (!r || !out_idx) return RIO_ERR_UNKNOWN;
    std::shared_ptr<Batch> b;
    size_t slot;
    {
        std::lock_guard<std::mutex> g(r->mu);
        if (r->stop) return RIO_ERR_UNKNOWN;
        if (!r->open) { r->open = std::make_shared<Batch>(); r->open_since = std::chrono::steady_clock::now(); }
        b = r->open;
        slot = b->keys.size();
        b->keys.push_back(key);
        r->calls++;
    }
    r->cv_work.notify_one();
    std::unique_lock<std::mutex> lk(b->mu);
    b->cv.wait(lk, [&] { return b->done; });
    if (b->status != RIO_OK) { t_resolver_error = b->error; return b->status; }
    *out_idx = b->out[slot];
    return RIO_OK;
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
