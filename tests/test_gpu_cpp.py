"""Builds the C++ conformance harness against the C++ mirror of the trait and runs it (GPU), and checks on CPU that it
compiles and links against librio_cuda.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "backend_conformance")


def _build():
    from rio_rs_b200 import build

    build.build()
    src = os.path.join(ROOT, "tests", "cpp", "backend_conformance.cpp")
    libdir = os.path.join(ROOT, "rio_rs_b200")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(libdir, "librio_cuda.so"))):
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-o", EXE, src, "-L" + libdir, "-lrio_cuda", "-Wl,-rpath," + libdir])
    return EXE


def test_cpp_mirror_compiles_and_links():
    assert os.path.exists(_build())


@pytest.mark.gpu
def test_cpp_backend_conformance_on_gpu():
    r = subprocess.run([_build()], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all passed" in r.stdout
