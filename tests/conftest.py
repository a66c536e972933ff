import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


# every step's tables are bounds-checked on upload while the tests run (pa_step_validate)
os.environ.setdefault("PA_VALIDATE_STEP", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the checkers and the product library are built (idempotent, seconds)."""
    import __graft_entry__ as ge
    ge.build(quiet=True)
    yield
