import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: PARAM_OPT end-to-end cases")


def load_pkg():
    """Import the hyphen-named package directory tfhe-aes_b200/ as module `tfhe_aes_b200`."""
    if "tfhe_aes_b200" in sys.modules:
        return sys.modules["tfhe_aes_b200"]
    d = os.path.join(ROOT, "tfhe-aes_b200")
    spec = importlib.util.spec_from_file_location("tfhe_aes_b200", os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    m = importlib.util.module_from_spec(spec)
    sys.modules["tfhe_aes_b200"] = m
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def orc():
    import orc as _orc
    return _orc


@pytest.fixture(scope="session")
def oracle_test(orc):
    return orc.Oracle(orc.param_test(), seed=1)


@pytest.fixture(scope="session")
def oracle_test2(orc):
    return orc.Oracle(orc.param_test2(), seed=5)


@pytest.fixture(scope="session")
def oracle_opt(orc):
    return orc.Oracle(orc.param_opt(), seed=2)


def _engine_with_oracle_keys(pkg, params, oracle):
    eng = pkg.Engine(params)
    eng.load_keys(oracle.bsk(), oracle.ksk(), oracle.pfpksk())
    return eng


@pytest.fixture(scope="session")
def engine_test(pkg, oracle_test):
    return _engine_with_oracle_keys(pkg, pkg.param_test(), oracle_test)


@pytest.fixture(scope="session")
def engine_test2(pkg, oracle_test2):
    return _engine_with_oracle_keys(pkg, pkg.param_test2(), oracle_test2)


@pytest.fixture(scope="session")
def engine_opt(pkg, oracle_opt):
    return _engine_with_oracle_keys(pkg, pkg.param_opt(), oracle_opt)


def torus_absdiff(a, b):
    """max |a - b| with a, b read as elements of Z / 2^64 (signed distance)."""
    with np.errstate(over="ignore"):
        d = (np.asarray(a, dtype=np.uint64) - np.asarray(b, dtype=np.uint64)).astype(np.int64)
    return int(np.abs(d.astype(np.float64)).max()) if d.size else 0
