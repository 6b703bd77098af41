import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_PY = os.path.join(ROOT, "gr-mimo-ofdm-jrc_b200", "python")
for p in (ROOT, PKG_PY):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as _orc
    _orc.lib()
    return _orc


@pytest.fixture(scope="session")
def jrc():
    import mimo_ofdm_jrc
    return mimo_ofdm_jrc
