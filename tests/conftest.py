import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a fresh checkout has no libemei_b200.so (built files are git-ignored): build it once (nvcc cross-compiles
    # without a GPU) so that the suite tests the library instead of failing at import
    lib = os.path.join(ROOT, "emei_b200", "csrc", "libemei_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__

        __graft_entry__.build()


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))

    return load
