import json
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

DATA = os.path.join(REPO, "surface_17_iqt_b200", "data")
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_json(path):
    with open(path) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def error_models():
    return load_json(os.path.join(DATA, "error_models.json"))


@pytest.fixture(scope="session")
def correction_table():
    return load_json(os.path.join(DATA, "correction_table_depolarising.json"))


@pytest.fixture(scope="session")
def bitflip_table():
    return load_json(os.path.join(DATA, "correction_table_classical_bitflip.json"))


@pytest.fixture(scope="session")
def golden():
    def _load(name):
        return load_json(os.path.join(GOLDEN, name))
    return _load


@pytest.fixture(scope="session")
def oracle(error_models, correction_table, bitflip_table):
    from oracle.s17_oracle import ShotOracle
    return ShotOracle(error_models, correction_table, bitflip_table)
