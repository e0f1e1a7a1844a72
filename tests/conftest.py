import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_devices():
    """CUDA devices on this machine.  Asked of the product library (built first if it is missing and nvcc is
    here); if that fails, of the driver through torch -- so that a GPU box with a broken or missing library
    FAILS the gpu tier loudly instead of skipping it."""
    try:
        import shutil

        from gelato_b200 import engine

        if not os.path.exists(engine.LIB_PATH) and shutil.which("nvcc"):
            engine.build_library()
        return int(engine.load_library().gelato_device_count())
    except Exception:
        pass
    try:
        import torch

        return int(torch.cuda.device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a machine without a CUDA device skips the gpu tier (and the gpu
    parameters of mixed tests) instead of failing it; `-m gpu` on such a machine reports skips too."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu") is not None]
    if not gpu_items or _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (the gpu tier runs on the B200 box: pytest -m gpu)")
    for it in gpu_items:
        it.add_marker(skip)


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
