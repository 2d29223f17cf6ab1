"""pytest plugin, loaded only by tests/test_engine_host_sim.py (`-p hostsim_plugin`): points the Python mirror of the provider at
tests/_build/librio_cuda_hostsim.so -- csrc/engine.cu + resolver.cu + durable.cu compiled with g++ against a synchronous stand-in for
the CUDA runtime and plain restatements of the kernel launchers (tests/cpp/hostsim/) -- so that the `-m gpu` test bodies exercise the
engine's HOST logic on a box without a GPU.  Test infrastructure: the product's loader (rio_rs_b200/_native.py) has no such switch."""
import os


def pytest_configure(config):
    from rio_rs_b200 import _native

    path = os.environ["RIO_HOSTSIM_LIBRARY"]
    assert os.path.exists(path), path
    _native.library_path = lambda: path
    _native._lib = None
