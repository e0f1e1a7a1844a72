import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_test_libs():
    """Build the CPU-side checkers once: the oracle (both flavours) and the host
    emulator of the kernels, and the product library if it is missing or stale (the same
    nvcc command as __graft_entry__.build())."""
    import shutil

    from gelato_b200 import engine
    from oracle import leaves

    if shutil.which("nvcc"):  # the product library (cross-compiles without a GPU); a no-op when up to date
        engine.build_library()
    leaves.build()
    import emu_binding

    emu_binding.build()
