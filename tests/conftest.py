import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_bytes(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def oracle():
    from oracle.harness import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def model():
    from tools.model_harness import Model
    return Model()


@pytest.fixture(scope="session")
def golden_table():
    import json
    with open(os.path.join(GOLDEN, "golden_sizes.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def encoder():
    import rust_brotli_b200 as rb
    enc = rb.DeviceEncoder(0)
    yield enc
    enc.close()
