import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the libraries are build artefacts (git-ignored): build them when a fresh checkout runs the tests
    if not os.path.exists(os.path.join(ROOT, "cfd_proxy_b200", "libcfdp_b200.so")):
        subprocess.run(["make", "-C", ROOT, "lib"], check=True, capture_output=True)
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.run(["make", "-C", ROOT, "oracle"], check=True, capture_output=True)


@pytest.fixture()
def session_factory():
    """Yields a function creating Sessions; closes the active one at teardown (engine is a singleton)."""
    from cfd_proxy_b200.driver import Session
    made = []

    def make(*a, **kw):
        if made:
            made[-1].close()
        s = Session(*a, **kw)
        made.append(s)
        return s

    yield make
    for s in made:
        s.close()
    for k in ("CFDP_TILE_POINTS", "CFDP_TILE_ORDER"):
        os.environ.pop(k, None)
