"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"`: oracle vs the reference's known-answer vectors, the product's host-side layout algebra vs the
oracle, C-ABI symbol/loader checks, the shard planner under gloo (world_size 2).  No compute calls on a GPU.
`-m gpu`: parity tests proper -- every call goes through the C ABI of librstsr_cuda.so and is compared with the
oracle on the same seeded inputs.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # The C-ABI library is the product: build it (nvcc cross-compiles without a GPU) if a fresh checkout has none.
    lib = os.path.join(ROOT, "rstsr_b200", "lib", "librstsr_cuda.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()


def _cuda_device_count() -> int:
    try:
        import rstsr_b200 as rt
        return rt.DeviceCuda.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the gpu-marked tests are SKIPPED (they cannot run: the product has no CPU fallback)
    instead of erroring one by one; on the B200 box nothing is skipped."""
    if not any("gpu" in it.keywords for it in items):
        return
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (librstsr_cuda.so has no CPU fallback)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def dev():
    import rstsr_b200 as rt
    d = rt.DeviceCuda(0, rt.ROW_MAJOR)
    yield d
    d.close()


@pytest.fixture(scope="session")
def dev_col():
    import rstsr_b200 as rt
    d = rt.DeviceCuda(0, rt.COL_MAJOR)
    yield d
    d.close()
