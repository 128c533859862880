import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    class Merged(dict):
        files = property(lambda self: list(self))

    def load(name):
        g = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        extra = os.path.join(GOLDEN, name + "_zoo.npz")   # Car / MarBot (oracle/make_golden_zoo.py)
        if os.path.exists(extra):
            z = np.load(extra, allow_pickle=False)
            return Merged({**{k: g[k] for k in g.files}, **{k: z[k] for k in z.files}})
        return g
    return load
