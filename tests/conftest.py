import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g
    return g.build()


@pytest.fixture(scope="session")
def small_dbs(built):
    """config -> (.mxy bytes, 1 MiB of that config's log) at 1 % scale."""
    from matchy_b200 import synth
    out = {}
    for cfg in (1, 2, 3, 4, 5):
        out[cfg] = (synth.build_db(cfg, 0.01), synth.gen_log(cfg, 1 << 20, 0.01).tobytes())
    return out
