import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def gpu_ctx():
    """One libzpaqgpu context for the whole session.  Fails loudly when the CUDA library is not
    built or no device is present -- there is no fallback to skip to."""
    import zpaq_v_b200 as z
    ctx = z.Context()
    yield ctx
    ctx.close()
