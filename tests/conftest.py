import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture
def vm_engines(monkeypatch):
    """route the package's engines to the numpy interpreter of the op stream (CPU tier only)."""
    from np_vm import NumpyEngine
    from kagomeperiodicbp_b200 import belief_propagation as bp
    engines = {}

    def fake(key="default", device=0):
        if key not in engines:
            engines[key] = NumpyEngine()
        return engines[key]

    monkeypatch.setattr(bp, "get_engine", fake)
    import kagomeperiodicbp_b200.bubblecon as bc
    monkeypatch.setattr(bc, "get_engine", fake)
    return engines
