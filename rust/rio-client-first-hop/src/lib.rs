//! Deterministic first hop for `rio_rs::client::Client`.  Source only: this image has no cargo.
//!
//! Call site: `Client::get_service_object_address` (rio-rs/src/client/mod.rs:235-267).  On a cache miss the reference
//! picks `servers.choose(&mut rng)` (:254-263) and lets the server answer `Redirect`; with this crate the miss arm becomes
//!
//! ```ignore
//! None => self.first_hop.owner(&service_object_type.to_string(), &service_object_id.to_string())
//!             .ok_or(ClientError::NoServersAvailable)?,
//! ```
//!
//! and `fetch_active_servers` (:153-172) rebuilds the ring whenever it replaces `active_servers`.
#![allow(non_camel_case_types)]
use libc::{c_char, size_t};

#[repr(C)]
pub struct rio_client_ring {
    _private: [u8; 0],
}
pub const RIO_CLIENT_OK: i32 = 0;
pub const RIO_CLIENT_NONE: u32 = 0xFFFF_FFFF;

extern "C" {
    pub fn rio_client_ring_create(addresses: *const *const c_char, address_lens: *const size_t, weights: *const u32, n: u32, out: *mut *mut rio_client_ring) -> i32;
    pub fn rio_client_ring_destroy(ring: *mut rio_client_ring);
    pub fn rio_client_ring_set_policy(ring: *mut rio_client_ring, policy: u32, trie_bits: u32) -> i32;
    pub fn rio_client_ring_size(ring: *const rio_client_ring) -> u32;
    pub fn rio_client_ring_address(ring: *const rio_client_ring, index: u32, buf: *mut c_char, cap: size_t, out_len: *mut size_t) -> i32;
    pub fn rio_client_object_key(ty: *const c_char, ty_len: size_t, id: *const c_char, id_len: size_t) -> u64;
    pub fn rio_client_first_hop(ring: *const rio_client_ring, ty: *const c_char, ty_len: size_t, id: *const c_char, id_len: size_t, out_index: *mut u32) -> i32;
    pub fn rio_client_first_hop_key(ring: *const rio_client_ring, key: u64, out_index: *mut u32) -> i32;
    pub fn rio_client_first_hop_batch(ring: *const rio_client_ring, keys: *const u64, n: size_t, out_index: *mut u32) -> i32;
}

/// The client's view of the active servers plus the rendezvous pick over it.
pub struct FirstHop {
    ring: *mut rio_client_ring,
    addresses: Vec<String>,
}
// the ring is immutable after creation
unsafe impl Send for FirstHop {}
unsafe impl Sync for FirstHop {}

impl FirstHop {
    /// `weights`: `None` = every server weighs 1 (the reference has no weights).
    pub fn new(addresses: Vec<String>, weights: Option<&[u32]>) -> Option<Self> {
        let ptrs: Vec<*const c_char> = addresses.iter().map(|a| a.as_ptr() as *const c_char).collect();
        let lens: Vec<size_t> = addresses.iter().map(|a| a.len()).collect();
        let mut ring = std::ptr::null_mut();
        let w = weights.map_or(std::ptr::null(), |w| w.as_ptr());
        if weights.map_or(false, |w| w.len() != addresses.len()) {
            return None;
        }
        let st = unsafe { rio_client_ring_create(ptrs.as_ptr(), lens.as_ptr(), w, addresses.len() as u32, &mut ring) };
        (st == RIO_CLIENT_OK).then(|| FirstHop { ring, addresses })
    }

    /// Owner of `(type, id)` under the servers' weighted rendezvous hash; `None` = no live server.
    pub fn owner(&self, service_object_type: &str, service_object_id: &str) -> Option<String> {
        let mut idx = RIO_CLIENT_NONE;
        let st = unsafe {
            rio_client_first_hop(
                self.ring,
                service_object_type.as_ptr() as *const c_char,
                service_object_type.len(),
                service_object_id.as_ptr() as *const c_char,
                service_object_id.len(),
                &mut idx,
            )
        };
        (st == RIO_CLIENT_OK && idx != RIO_CLIENT_NONE).then(|| self.addresses[idx as usize].clone())
    }
}

impl Drop for FirstHop {
    fn drop(&mut self) {
        unsafe { rio_client_ring_destroy(self.ring) }
    }
}
