import importlib
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

PKG_NAME = "myers-furcht-et-al_gab1-shp2-pde-model_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="session")
def ofe(oracle_mod):
    """The Julia-surface frontend bound to the CPU oracle (the checker)."""
    return oracle_mod.frontend()


@pytest.fixture(scope="session")
def ensemble(pkg):
    return pkg.params.load_parameter_ensemble()
