"""Builds librio_cuda.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "librio_cuda.so")
SOURCES = ["k_assign.cu", "k_trie.cu", "k_affinity_umma.cu", "k_directory.cu", "engine.cu", "resolver.cu", "durable.cu"]
HEADERS = ["kernels.cuh", "spec.cuh", "bounded_tail.cuh", "trie_table.hpp", os.path.join("..", "..", "include", "rio_cuda.h"), os.path.join("..", "..", "include", "rio_cuda_dev.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_fresh():
    if not os.path.exists(SO):
        return False
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source of the product into rio_rs_b200/librio_cuda.so.  RIO_BUILD_TUNING=1 also compiles the A/B
    tuning points of the flat rendezvous kernel (tools/tune_assign.py); the shipped library carries the default only."""
    if not force and is_fresh():
        return SO
    cmd = [_nvcc()] + NVCC_FLAGS + (["-DRIO_ASSIGN_TUNING"] if os.environ.get("RIO_BUILD_TUNING") else []) + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports CC=/opt/gcc/bin/gcc; nvcc should use the system g++
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return SO


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
