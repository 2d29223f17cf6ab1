/*
 * directory_model.cpp -- CPU ORACLE for the placement *directory* and the per-request placement
 * policy (test infrastructure, NOT product code; see oracle/README.md for who may load it).
 *
 * This is the part of the path the reference DOES pin.  It restates, statement for statement:
 *   - LocalObjectPlacement                 rio-rs/src/object_placement/local.rs:12-68
 *       map type  Arc<RwLock<HashMap<String,String>>>                 local.rs:12
 *       update    key = format!("{}.{}"); Some(addr) => insert/overwrite, None => remove   local.rs:22-40
 *       lookup    get(key).cloned()                                   local.rs:42-49
 *       clean_server  retain(|_, v| *v != address)                    local.rs:51-58
 *       remove    remove(key)                                         local.rs:60-68
 *   - MembershipStorage::active_members / is_active (default methods) rio-rs/src/cluster/storage/mod.rs:95-110
 *     over LocalStorage::members() == Vec<Member> clone               rio-rs/src/cluster/storage/local.rs:61-63
 *   - Service::get_or_create_placement                                rio-rs/src/service.rs:193-254
 *
 * The reference cannot be compiled here (no cargo/rustc), so the pins are its own
 * known-answer tests, restated in tests/test_oracle_directory.py:
 *   local.rs:75-114, tests/object_placement_backend.rs:11-34, sqlite.rs:149-193.
 *
 * The same object doubles as the timed "port" CPU baseline / `bench.py --impl reference` arm: it keeps the
 * reference's per-request costs on purpose (key formatting, value clone, O(M) member-vector clone per
 * is_active) instead of optimising them away.
 */
#include <cstdint>
#include <cstring>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <atomic>
#include <chrono>

namespace {

struct Member {            // cluster/storage/mod.rs:21-26
    std::string ip, port;
    bool active;
    int64_t last_seen;
    std::string address() const { return ip + ":" + port; }   // mod.rs:56-58
};

struct Membership {        // cluster/storage/local.rs:13-17 (members only; failures are off-path)
    mutable std::shared_mutex mu;
    std::vector<Member> members;

    void push(const std::string &ip, const std::string &port, bool active) {
        std::unique_lock<std::shared_mutex> g(mu);
        members.push_back(Member{ip, port, active, 0});
    }
    void remove(const std::string &ip, const std::string &port) {          // local.rs:26-30
        std::unique_lock<std::shared_mutex> g(mu);
        std::vector<Member> keep;
        for (auto &m : members) if (m.ip != ip || m.port != port) keep.push_back(m);
        members.swap(keep);
    }
    void set_is_active(const std::string &ip, const std::string &port, bool a) {   // local.rs:32-42
        std::unique_lock<std::shared_mutex> g(mu);
        for (auto &m : members) if (m.ip == ip && m.port == port) m.active = a;
    }
    std::vector<Member> all() const {                                       // local.rs:61-63 (clone)
        std::shared_lock<std::shared_mutex> g(mu);
        return members;
    }
    std::vector<Member> active_members() const {                            // mod.rs:95-99
        std::vector<Member> v = all();
        std::vector<Member> out;
        for (auto &m : v) if (m.active) out.push_back(m);
        return out;
    }
    bool is_active(const std::string &ip, const std::string &port) const {  // mod.rs:102-110
        std::vector<Member> act = active_members();
        for (auto &m : act) if (m.ip == ip && m.port == port) return true;
        return false;
    }
};

struct Directory {         // local.rs:12-18
    mutable std::shared_mutex mu;
    std::unordered_map<std::string, std::string> map;

    static std::string key(const char *type, const char *id) {              // local.rs:26-29,43,61
        std::string k(type); k += '.'; k += id; return k;
    }
    void update(const char *type, const char *id, const char *addr) {       // local.rs:22-40
        std::string k = key(type, id);
        std::unique_lock<std::shared_mutex> g(mu);
        if (addr) map[k] = std::string(addr); else map.erase(k);
    }
    bool lookup(const char *type, const char *id, std::string &out) const { // local.rs:42-49
        std::string k = key(type, id);
        std::shared_lock<std::shared_mutex> g(mu);
        auto it = map.find(k);
        if (it == map.end()) return false;
        out = it->second;                                                   // .cloned()
        return true;
    }
    void clean_server(const char *addr) {                                   // local.rs:51-58
        std::string a(addr);
        std::unique_lock<std::shared_mutex> g(mu);
        for (auto it = map.begin(); it != map.end();) { if (it->second == a) it = map.erase(it); else ++it; }
    }
    void remove(const char *type, const char *id) {                         // local.rs:60-68
        std::string k = key(type, id);
        std::unique_lock<std::shared_mutex> g(mu);
        map.erase(k);
    }
};

struct Model { Directory dir; Membership mem; };

/* Service::get_or_create_placement, service.rs:193-254 */
std::string get_or_create_placement(Model &m, const std::string &self_address, const char *type, const char *id) {
    std::string addr; bool have = m.dir.lookup(type, id, addr);             // :199-201
    if (have) {
        size_t c = addr.find(':');                                          // splitn(2, ":") :205-207
        std::string ip = c == std::string::npos ? addr : addr.substr(0, c);
        std::string port = c == std::string::npos ? std::string() : addr.substr(c + 1);
        if (ip.empty() || port.empty()) {                                   // :213-222
            m.dir.remove(type, id); have = false;
        } else if (!m.mem.is_active(ip, port)) {                            // :226-238
            m.dir.clean_server(addr.c_str()); have = false;
        }
    }
    if (have) return addr;                                                  // :241-242
    m.dir.update(type, id, self_address.c_str());                           // :244-251
    return self_address;                                                    // :252
}

/* Service::check_address_mismatch, service.rs:261-298.  Returns 0 = Ok(()), 1 = Err(Redirect(server_address)),
 * 2 = clean_server(server_address) applied + Err(DeallocateServiceObject), 3 = Err(Unknown("Malformed address: Missing PORT ..")) */
int check_address_mismatch(Model &m, const std::string &self_address, const std::string &server_address) {
    if (server_address == self_address) return 0;                           // :262-264
    // `split(':')` (not splitn): ip = first piece (always present), port = SECOND piece, further pieces ignored  :266-278
    size_t c = server_address.find(':');
    if (c == std::string::npos) return 3;                                   // "Missing PORT"
    std::string ip = server_address.substr(0, c);
    size_t c2 = server_address.find(':', c + 1);
    std::string port = server_address.substr(c + 1, c2 == std::string::npos ? std::string::npos : c2 - c - 1);
    if (m.mem.is_active(ip, port)) return 1;                                // :280-288
    m.dir.clean_server(server_address.c_str());                             // :291-296
    return 2;                                                               // :297
}

}  // namespace

extern "C" {

int dm_check_address_mismatch(void *h, const char *self_address, const char *server_address) {
    return check_address_mismatch(*(Model *)h, self_address, server_address);
}

void *dm_new() { return new Model(); }
void dm_free(void *h) { delete (Model *)h; }

void dm_update(void *h, const char *type, const char *id, const char *addr_or_null) { ((Model *)h)->dir.update(type, id, addr_or_null); }
/* returns length of the address copied into buf (truncated to cap), or -1 when lookup == None */
int64_t dm_lookup(void *h, const char *type, const char *id, char *buf, size_t cap) {
    std::string out;
    if (!((Model *)h)->dir.lookup(type, id, out)) return -1;
    size_t n = out.size() < cap ? out.size() : cap;
    if (n) memcpy(buf, out.data(), n);
    return (int64_t)out.size();
}
void dm_clean_server(void *h, const char *addr) { ((Model *)h)->dir.clean_server(addr); }
void dm_remove(void *h, const char *type, const char *id) { ((Model *)h)->dir.remove(type, id); }
uint64_t dm_len(void *h) { Model *m = (Model *)h; std::shared_lock<std::shared_mutex> g(m->dir.mu); return m->dir.map.size(); }

void dm_member_push(void *h, const char *ip, const char *port, int active) { ((Model *)h)->mem.push(ip, port, active != 0); }
void dm_member_remove(void *h, const char *ip, const char *port) { ((Model *)h)->mem.remove(ip, port); }
void dm_member_set_active(void *h, const char *ip, const char *port, int active) { ((Model *)h)->mem.set_is_active(ip, port, active != 0); }
int dm_member_is_active(void *h, const char *ip, const char *port) { return ((Model *)h)->mem.is_active(ip, port) ? 1 : 0; }

int64_t dm_get_or_create_placement(void *h, const char *self_address, const char *type, const char *id, char *buf, size_t cap) {
    std::string out = get_or_create_placement(*(Model *)h, self_address, type, id);
    size_t n = out.size() < cap ? out.size() : cap;
    if (n) memcpy(buf, out.data(), n);
    return (int64_t)out.size();
}

/*
 * Timed reference-path sample (bench.py --impl reference / cpu_baseline "port"):
 * ids ("Obj", decimal(first+i)) for i<n are resolved through get_or_create_placement by `threads`
 * host threads (the reference runs one tokio task per connection, server.rs:303); thread t acts as the
 * server at member (t mod M), i.e. requests land on a server the way the client's uniform-random first hop
 * spreads them (client/mod.rs:254-263).  Members are "10.0.(j>>8).(j&255)":"5000", all active.
 * Returns elapsed seconds; *placed = number of ids resolved.
 */
double dm_bench_resolve(uint64_t first, uint64_t n, uint32_t M, int threads, uint64_t *placed) {
    Model m;
    std::vector<std::string> self;
    for (uint32_t j = 0; j < M; j++) {
        std::string ip = "10.0." + std::to_string(j >> 8) + "." + std::to_string(j & 255);
        m.mem.push(ip, "5000", true);
        self.push_back(ip + ":5000");
    }
    if (threads < 1) threads = 1;
    std::atomic<uint64_t> done{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back([&, t]() {
        uint64_t lo = n * t / threads, hi = n * (t + 1) / threads, cnt = 0;
        const std::string &me = self[t % M];
        for (uint64_t i = lo; i < hi; i++) {
            std::string id = std::to_string(first + i);
            std::string a = get_or_create_placement(m, me, "Obj", id.c_str());
            cnt += !a.empty();
        }
        done += cnt;
    });
    for (auto &x : th) x.join();
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (placed) *placed = done.load();
    return s;
}

/* C1 (BASELINE.json configs[0]): LocalObjectPlacement::lookup over n pre-populated ids, M-node cluster.
 * Returns seconds for `reps` passes of n lookups on one thread. */
double dm_bench_lookup(uint64_t n, uint32_t M, uint32_t reps, uint64_t *hits) {
    Model m;
    for (uint64_t i = 0; i < n; i++) {
        std::string addr = "10.0.0." + std::to_string(i % M) + ":5000";
        m.dir.update("Obj", std::to_string(i).c_str(), addr.c_str());
    }
    uint64_t h = 0; std::string out;
    auto t0 = std::chrono::steady_clock::now();
    for (uint32_t r = 0; r < reps; r++)
        for (uint64_t i = 0; i < n; i++) h += m.dir.lookup("Obj", std::to_string(i).c_str(), out);
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (hits) *hits = h;
    return s;
}

}  // extern "C"
