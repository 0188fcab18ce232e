import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle
    oracle.lib()
    return oracle


_KEYS = {}


@pytest.fixture(scope="session")
def keyset(O):
    """keyset(name) -> (P, sk, ck), generated once per parameter set with fixed seeds."""
    def get(name, with_ksk=True):
        k = (name, with_ksk)
        if k not in _KEYS:
            P = O.get_params(name)
            sk = O.SecretKey(P, 0xC0FFEE + len(name))
            ck = O.CloudKey(sk, 0xBEEF, with_ksk=with_ksk)
            _KEYS[k] = (P, sk, ck)
        return _KEYS[k]
    return get
