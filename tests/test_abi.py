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


# ---- signature-level agreement of the three statements of the ABI: the C headers, the Rust extern blocks, the ctypes table ------
def _c_prototypes(header):
    """{name: (return type, [argument types])} of every prototype in include/<header>, in a canonical spelling:
    'const char *' -> '*const c_char', 'uint32_t' -> 'u32', 'T name[N]' -> pointer to T ..."""
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    out = {}
    scalar = {"uint8_t": "u8", "uint32_t": "u32", "uint64_t": "u64", "int32_t": "i32", "size_t": "usize", "float": "f32", "double": "f64", "char": "c_char",
              "void": "c_void", "rio_status": "i32", "unsigned long long": "u64"}

    def canon(t):
        t = " ".join(t.replace("*", " * ").split())
        parts = t.split(" ")
        # peel pointers from the right: "const char * const *" -> base "const char", ptrs ["* const", "*"]
        base, ptrs, i = [], [], 0
        while i < len(parts) and parts[i] != "*":
            base.append(parts[i]); i += 1
        while i < len(parts):
            assert parts[i] == "*", t
            if i + 1 < len(parts) and parts[i + 1] == "const":
                ptrs.append("const"); i += 2
            else:
                ptrs.append("mut"); i += 1
        const_base = "const" in base
        name = " ".join(p for p in base if p != "const")
        r = scalar.get(name, name)
        # the constness of a pointer level in Rust is the constness of what it points TO
        quals = [const_base] + [p == "const" for p in ptrs[:-1]]
        for q in quals[:len(ptrs)]:
            r = ("*const " if q else "*mut ") + r
        return r

    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ ]*?[ \*]+)\b((?:rio_cuda|rio_client|rio_dev)_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        argt = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                arr = re.match(r"(.*?)\b([A-Za-z_][A-Za-z0-9_]*)\s*\[[^\]]*\]$", a)
                if arr:
                    argt.append(canon(arr.group(1) + " *"))
                else:
                    argt.append(canon(re.sub(r"\b[A-Za-z_][A-Za-z0-9_]*$", "", a)))
        out[name] = ("()" if ret.strip() == "void" else canon(ret), argt)
    return out


def _rust_prototypes(crate):
    src = open(os.path.join(ROOT, "rust", crate, "src", "lib.rs")).read()
    blocks = "".join(re.findall(r'extern "C" \{(.*?)\n\}', src, flags=re.S))
    out = {}
    for m in re.finditer(r"fn ((?:rio_cuda|rio_client)_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->\s*([^;]+))?;", blocks, flags=re.S):
        args = [a.split(":", 1)[1].strip() for a in m.group(2).split(",") if ":" in a]
        norm = lambda t: " ".join(t.replace("size_t", "usize").replace("rio_status", "i32").split())   # noqa: E731
        out[m.group(1)] = (norm(m.group(3)) if m.group(3) else "()", [norm(a) for a in args])
    return out


@pytest.mark.parametrize("header,crate", [("rio_cuda.h", "rio-cuda-sys"), ("rio_client.h", "rio-client-first-hop")])
def test_rust_extern_signatures_equal_the_c_prototypes(header, crate):
    """No cargo here, so the Rust declarations cannot be compiled -- but every one of them can be compared, type by type, with the
    prototype in the header it binds: argument count, order, width, pointer depth and constness."""
    c, r = _c_prototypes(header), _rust_prototypes(crate)
    assert len(c) > 8 and set(c) == set(r), set(c) ^ set(r)
    bad = {n: (c[n], r[n]) for n in c if c[n] != r[n]}
    assert not bad, bad


def test_ctypes_binding_matches_the_c_prototypes(native):
    """Widths of every scalar argument / return value and 'pointer or not' of every position, ctypes table vs header."""
    c = _c_prototypes("rio_cuda.h")
    c.update(_c_prototypes("rio_cuda_dev.h"))
    width = {"u8": 1, "u32": 4, "i32": 4, "u64": 8, "usize": C.sizeof(C.c_size_t), "f32": 4, "f64": 8}

    def shape(t):   # canonical C type -> ('ptr',) or ('int'|'float', bytes)
        if t.startswith("*"):
            return ("ptr",)
        return ("float" if t in ("f32", "f64") else "int", width[t])

    def cshape(t):
        if t is None:
            return None
        if isinstance(t, type) and issubclass(t, (C._Pointer, C.c_char_p, C.c_void_p)) or t in (C.c_char_p, C.c_void_p):
            return ("ptr",)
        return ("float" if t in (C.c_float, C.c_double) else "int", C.sizeof(t))

    for name, (res, args) in native.SIGNATURES.items():
        assert name in c, name
        cret, cargs = c[name]
        assert (None if cret == "()" else shape(cret)) == cshape(res), (name, cret, res)
        assert [shape(a) for a in cargs] == [cshape(a) for a in args], (name, cargs, args)


def test_headers_are_plain_c99_and_the_cpp_harnesses_compile(tmp_path):
    """The boundary is a C ABI: the headers must go through a C compiler (what bindgen / cgo would see) with -pedantic, and the C++
    conformance harnesses (run on the GPU box by tests/test_gpu_cpp.py) must at least compile against them on every box."""
    import shutil
    import subprocess

    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not gcc or not gxx:
        pytest.skip("no host compiler")
    src = tmp_path / "hdr.c"
    src.write_text('#include "rio_cuda.h"\n#include "rio_cuda_dev.h"\n#include "rio_client.h"\nint main(void) { return (int)(RIO_ABI_VERSION * 0); }\n')
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "hdr.o")])
    cpp = [os.path.join(ROOT, "tests", "cpp", f) for f in sorted(os.listdir(os.path.join(ROOT, "tests", "cpp"))) if f.endswith(".cpp")]
    assert len(cpp) >= 3
    subprocess.check_call([gxx, "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), "-I" + ROOT] + cpp)


def test_rust_provider_calls_pass_the_declared_number_of_arguments():
    """Every FFI call in the (uncompiled) Rust provider / client crates passes as many arguments as the prototype declares."""
    protos = {}
    for hdr in ("rio_cuda.h", "rio_client.h"):
        protos.update(_c_prototypes(hdr))
    seen = 0
    for crate in ("gpu_object_placement", "rio-client-first-hop", "rio-cuda-sys"):
        src = open(os.path.join(ROOT, "rust", crate, "src", "lib.rs")).read()
        src = re.sub(r"//[^\n]*", "", src)
        src = re.sub(r'extern "C" \{.*?\n\}', "", src, flags=re.S)        # declarations are compared elsewhere
        for m in re.finditer(r"\b((?:rio_cuda|rio_client)_[a-z0-9_]+)\s*\(", src):
            name, i, depth, nargs, cur = m.group(1), m.end(), 1, 0, ""
            while depth:
                ch = src[i]
                if ch in "([{":
                    depth += 1
                elif ch in ")]}":
                    depth -= 1
                    if depth == 0:
                        break
                if ch == "," and depth == 1:
                    nargs += 1 if cur.strip() else 0
                    cur = ""
                else:
                    cur += ch
                i += 1
            nargs += 1 if cur.strip() else 0
            assert name in protos, (crate, name)
            assert nargs == len(protos[name][1]), (crate, name, nargs, protos[name][1])
            seen += 1
    assert seen >= 15, seen
