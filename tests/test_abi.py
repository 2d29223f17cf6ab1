"""CPU tests of the drop-in boundary: librio_cuda.so loads, exports every symbol include/rio_cuda.h declares,
the ctypes binding covers all of them, host-only helpers agree with the oracle, and -- on a box without a GPU --
creating an engine fails loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    from rio_rs_b200 import _native, build

    build.build()
    return _native


def _declared():
    hdr = open(os.path.join(ROOT, "include", "rio_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rio_cuda_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_a_nontrivial_surface():
    names = _declared()
    assert len(names) >= 50
    for must in ("rio_cuda_create", "rio_cuda_lookup_batch", "rio_cuda_upsert_batch", "rio_cuda_clean_node", "rio_cuda_remove_batch",
                 "rio_cuda_assign_batch", "rio_cuda_place_batch", "rio_cuda_rebalance", "rio_cuda_update_str", "rio_cuda_lookup_str"):
        assert must in names


def test_library_exports_every_declared_symbol(native):
    L = C.CDLL(native.library_path())
    for name in _declared():
        assert hasattr(L, name), name


def test_binding_covers_every_declared_symbol(native):
    assert sorted(native.SIGNATURES) == _declared()
    native.lib()
    assert native.lib().rio_cuda_abi_version() == 2


def test_host_key_helpers_match_oracle(native, oracle):
    L = native.lib()
    for t, i in [("obj", "1"), ("Test", "1"), ("MockService", "77"), ("", ""), ("a.b", "c")]:
        tb, ib = t.encode(), i.encode()
        assert L.rio_cuda_object_key(tb, len(tb), ib, len(ib)) == oracle.object_key(t, i)
    for a in ["0.0.0.0:8888", "10.0.3.255:5000", ""]:
        ab = a.encode()
        assert L.rio_cuda_node_seed(ab, len(ab)) == oracle.node_seed(a)


def test_no_gpu_means_loud_failure_not_fallback(native):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rio_rs_b200 import GpuObjectPlacement, Upstream

    with pytest.raises(Upstream) as e:
        GpuObjectPlacement()
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rio_rs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in src and "rio_oracle" not in src and "directory_model" not in src, f


def test_rust_ffi_declarations_cover_the_headers():
    """The Rust crates are source only (no cargo here), so at least keep their extern blocks in step with include/*.h."""
    import re

    for header, crate, prefix in (("rio_cuda.h", "rio-cuda-sys", "rio_cuda_"), ("rio_client.h", "rio-client-first-hop", "rio_client_")):
        declared = set(re.findall(r"\b(%s[a-z0-9_]+)\s*\(" % prefix, open(os.path.join(ROOT, "include", header)).read()))
        rust = set(re.findall(r"fn (%s[a-z0-9_]+)\s*\(" % prefix, open(os.path.join(ROOT, "rust", crate, "src", "lib.rs")).read()))
        assert declared == rust, (header, declared ^ rust)
