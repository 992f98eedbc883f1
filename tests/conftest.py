import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long CPU test (still part of the default CPU suite unless deselected)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


_ORACLES = {}


@pytest.fixture(scope="session")
def get_oracle(oracle_mod):
    """Session cache of oracle contexts keyed by (config name, extra overrides)."""
    from upcgen_b200.config import named_config

    def _get(name, extra=""):
        key = (name, extra)
        if key not in _ORACLES:
            P = named_config(name, extra)
            _ORACLES[key] = (P, oracle_mod.Oracle(P))
        return _ORACLES[key]

    return _get
