import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as _orc
    _orc.lib()
    return _orc


_KEYS = {}


@pytest.fixture(scope="session")
def oracle_keys(orc):
    """Seeded oracle key material per preset, generated once per session."""
    def get(preset, seed=1):
        k = (preset, seed)
        if k not in _KEYS:
            _KEYS[k] = orc.Keys(orc.preset(preset), seed)
        return _KEYS[k]
    return get


_CTX = {}


@pytest.fixture(scope="session")
def gpu_ctx(oracle_keys):
    """GPU context with the oracle's keys uploaded (same seeded keygen on both sides)."""
    import fhe_sign_b200 as fsb

    def get(preset, acc_bits=64, seed=1):
        k = (preset, acc_bits, seed)
        if k not in _CTX:
            keys = oracle_keys(preset, seed)
            ctx = fsb.Context(fsb.Params.preset(preset, acc_bits=acc_bits))
            ctx.upload_keys(keys.bsk, keys.ksk)
            _CTX[k] = ctx
        return _CTX[k]
    yield get
    for c in _CTX.values():
        c.close()
    _CTX.clear()


@pytest.fixture()
def rng():
    return np.random.default_rng(0xB200)
