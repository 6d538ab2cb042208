import importlib
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_pkg():
    """The package directory is `slam-sdvl_b200/` (not an identifier); it is imported as `slam_sdvl_b200`."""
    name = "slam_sdvl_b200"
    if name in sys.modules:
        return sys.modules[name]
    d = os.path.join(ROOT, "slam-sdvl_b200")
    spec = importlib.util.spec_from_file_location(name, os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def sw(pkg):
    return importlib.import_module("slam_sdvl_b200.synthworld")


@pytest.fixture(scope="session")
def scenes(pkg):
    return importlib.import_module("slam_sdvl_b200.scenes")


@pytest.fixture(scope="session")
def abi(pkg):
    return importlib.import_module("slam_sdvl_b200.abi")


@pytest.fixture(scope="session")
def binding(pkg):
    return importlib.import_module("slam_sdvl_b200.binding")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle_py
    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def seq_c2(sw):
    return sw.sequence("C2", 0, 6)


@pytest.fixture(scope="session")
def seq_c1(sw):
    return sw.sequence("C1", 3, 4)
