import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "latent-diffusion-segmentation_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library is built in-tree by __graft_entry__.build(); build it if a fresh checkout
    has not done so yet (nvcc cross-compiles without a GPU)."""
    lib = os.path.join(PKG, "lib", "libldmseg_b200.so")
    if not os.path.exists(lib):
        sys.path.insert(0, PKG)
        import build_native
        build_native.build(verbose=False)
    return lib
