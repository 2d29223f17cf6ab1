// first_hop.hpp -- C++ mirror of the client-side piece that include/rio_client.h replaces:
//   rio::FirstHop::get_service_object_address  <->  Client::get_service_object_address   rio-rs/src/client/mod.rs:235-267
//   rio::FirstHop::set_active_servers          <->  Client::fetch_active_servers         rio-rs/src/client/mod.rs:153-172
//   rio::FirstHop::record_redirect             <->  Redirect arm of the retry loop       rio-rs/src/client/tower_services.rs:158-168
// Header only; link with -lrio_client (plain C++ library, no CUDA).  The Rust equivalent is rust/rio-client-first-hop.
#pragma once
#include <list>
#include <map>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/rio_client.h"

namespace rio {

struct NoServersAvailable : std::runtime_error {   // ClientError::NoServersAvailable (client/mod.rs:260-261)
    NoServersAvailable() : std::runtime_error("no servers available") {}
};

class FirstHop {
  public:
    // LruCache limit, client/mod.rs:137; policy / trie_bits mirror the servers' solver (RIO_CLIENT_POLICY_*)
    explicit FirstHop(size_t cache_size = 1000, uint32_t policy = RIO_CLIENT_POLICY_HRW, uint32_t trie_bits = 0) : cache_size_(cache_size), policy_(policy), trie_bits_(trie_bits) {}
    ~FirstHop() { rio_client_ring_destroy(ring_); }
    FirstHop(const FirstHop &) = delete;
    FirstHop &operator=(const FirstHop &) = delete;

    // replaces the whole view ("not an incremental operation", client/mod.rs:150-152); weights empty = all 1
    void set_active_servers(const std::vector<std::string> &addresses, const std::vector<uint32_t> &weights = {}) {
        if (!weights.empty() && weights.size() != addresses.size()) throw std::invalid_argument("one weight per address");
        std::vector<const char *> p;
        std::vector<size_t> l;
        for (const auto &a : addresses) { p.push_back(a.data()); l.push_back(a.size()); }
        rio_client_ring *r = nullptr;
        if (rio_client_ring_create(p.data(), l.data(), weights.empty() ? nullptr : weights.data(), (uint32_t)addresses.size(), &r) != RIO_CLIENT_OK)
            throw std::runtime_error("rio_client_ring_create failed");
        if (rio_client_ring_set_policy(r, policy_, trie_bits_) != RIO_CLIENT_OK) { rio_client_ring_destroy(r); throw std::runtime_error("rio_client_ring_set_policy failed"); }
        rio_client_ring_destroy(ring_);
        ring_ = r;
        addresses_ = addresses;
    }

    // position of the owner in the address list, RIO_CLIENT_NONE without a live server
    uint32_t first_hop_index(const std::string &type, const std::string &id) const {
        uint32_t j = RIO_CLIENT_NONE;
        if (!ring_ || rio_client_first_hop(ring_, type.data(), type.size(), id.data(), id.size(), &j) != RIO_CLIENT_OK) return RIO_CLIENT_NONE;
        return j;
    }

    // cached address if any (client/mod.rs:251-253), else the rendezvous owner instead of a random server (:254-263)
    std::string get_service_object_address(const std::string &type, const std::string &id) {
        const auto k = std::make_pair(type, id);
        auto it = index_.find(k);
        if (it != index_.end()) {
            lru_.splice(lru_.begin(), lru_, it->second);
            return it->second->second;
        }
        const uint32_t j = first_hop_index(type, id);
        if (j == RIO_CLIENT_NONE) throw NoServersAvailable();
        return addresses_[j];
    }

    // the server answered Redirect(address): remember it, same key order as the lookup (the reference's differ)
    void record_redirect(const std::string &type, const std::string &id, const std::string &address) {
        const auto k = std::make_pair(type, id);
        auto it = index_.find(k);
        if (it != index_.end()) { lru_.erase(it->second); index_.erase(it); }
        lru_.emplace_front(k, address);
        index_[k] = lru_.begin();
        while (lru_.size() > cache_size_) { index_.erase(lru_.back().first); lru_.pop_back(); }
    }

  private:
    using Key = std::pair<std::string, std::string>;
    rio_client_ring *ring_ = nullptr;
    std::vector<std::string> addresses_;
    size_t cache_size_;
    uint32_t policy_ = RIO_CLIENT_POLICY_HRW, trie_bits_ = 0;
    std::list<std::pair<Key, std::string>> lru_;
    std::map<Key, std::list<std::pair<Key, std::string>>::iterator> index_;
};

}  // namespace rio
