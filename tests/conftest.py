"""pytest configuration: registers the `gpu` marker and makes sure the in-tree artefacts exist.

CPU suite (`-m "not gpu"`): oracle vs the reference's golden outputs, host-side table generators, the
C-ABI export list, the sharding logic under gloo. GPU suite (`-m gpu`): parity of the CUDA path with the
oracle through the C ABI. Nothing here reads /root/reference at run time.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    lib = os.path.join(ROOT, "fft-implementation-in-c_b200", "lib", "libfft_b200.so")
    ora = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(lib) or not os.path.exists(ora):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def F():
    import fftb200_loader
    return fftb200_loader.load()


@pytest.fixture(scope="session")
def O():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="session")
def port(O):
    return O.port()


@pytest.fixture(scope="session")
def gpu(F):
    if F.lib.fft_gpu_available() != 1:
        pytest.fail("no CUDA device visible: GPU tests must run on the B200 box (there is no CPU fallback)")
    F.require_gpu()
    return F
