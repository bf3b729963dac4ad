import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Native libraries (oracle, host, CUDA) built in-tree."""
    from frackyfrac_b200 import build
    from oracle import oracle

    oracle.build()
    build.build_all()
    return True


@pytest.fixture(scope="session")
def gpu_ctx(built):
    from frackyfrac_b200 import engine

    ctx = engine.Context(0)
    yield ctx
    ctx.close()
