import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    import tnad_b200 as T
    c = T.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np
    with open(os.path.join(ROOT, "tests", "golden", "published.json")) as f:
        pub = json.load(f)
    vec = np.load(os.path.join(ROOT, "tests", "golden", "vectors.npz"))
    return pub, vec


@pytest.fixture
def opt(ctx):
    """Set A/B options of the shared context for one test (tnad_set_option; the library reads the environment only
    once, at tnad_create) and remove them afterwards."""
    changed = []

    def setter(name, value):
        ctx.set_option(name, value)
        changed.append(name)
    yield setter
    for name in changed:
        ctx.set_option(name, None)


@pytest.fixture(scope="session")
def large():
    """Oracle outputs at the benchmarked sizes (tests/golden/make_golden_large.py)."""
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "large.npz"))
