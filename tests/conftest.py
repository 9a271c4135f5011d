import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: minutes-scale CPU test")


@pytest.fixture(scope="session")
def oracle_built():
    import oracle_lib
    oracle_lib.build_oracle()
    return True


_fx_cache = {}


@pytest.fixture(scope="session")
def get_fixture():
    import fixtures_def

    def _get(name):
        if name not in _fx_cache:
            _fx_cache[name] = fixtures_def.FIXTURES[name]()
        return _fx_cache[name]
    return _get
