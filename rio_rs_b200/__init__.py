"""rio_rs_b200 -- B200-native object placement behind rio-rs's ObjectPlacement trait.

The product is librio_cuda.so (hand-written sm_100a CUDA behind the C ABI in include/rio_cuda.h); this
package is the thin host-side mirror of the reference interface used by tests and bench.py:

    provider.GpuObjectPlacement   <->  trait ObjectPlacement      (rio-rs/src/object_placement/mod.rs:38-56)
    provider.ObjectId             <->  ObjectId(String, String)   (rio-rs/src/service_object.rs:19-26)
    provider.ObjectPlacementItem  <->  ObjectPlacementItem        (rio-rs/src/object_placement/mod.rs:20-34)
    provider.ObjectPlacementError <->  ObjectPlacementError       (rio-rs/src/errors.rs:136-142)

There is no CPU fallback: importing works anywhere, but creating a provider without a CUDA device raises.
"""
from .provider import (  # noqa: F401
    GpuObjectPlacement,
    ObjectId,
    ObjectPlacementError,
    ObjectPlacementItem,
    ObjectSet,
    Resolver,
    Unknown,
    Upstream,
)
from ._native import NONE, lib, library_path  # noqa: F401
