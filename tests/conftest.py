"""pytest configuration: registers the `gpu` marker and makes sure the native pieces exist.

`-m "not gpu"` runs on a CPU-only container: oracle vs golden vectors, oracle vs the unmodified
reference (when oracle/_ref/libazref.so exists), the shared device headers compiled for the host
(RNG, math, bitboards) and the host-emulation build of the engine logic vs the oracle, plus the
C-ABI symbol check of the CUDA library. `-m gpu` runs the CUDA library through the C ABI on a B200.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "alphazero-pybind11_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running parity sweep")


@pytest.fixture(scope="session", autouse=True)
def native_build():
    """Build what is missing (no-op when the prebuilt files travelled with the snapshot)."""
    import __graft_entry__ as ge

    ge.build_host_tests()
    ge.build_oracle()
    if not os.path.exists(ge.LIB):
        ge.build_cuda()
    ge.build_pybind()
    return ge


def has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def ref_available():
    import refdriver

    return refdriver.available()


needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libazref.so")),
                               reason="oracle/_ref/libazref.so not built (needs /root/reference)")
