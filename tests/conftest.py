import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "flux_lib_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def fcmod():
    """the product package; importing it loads libfluxcalc_b200.so (raises if it is not built)"""
    import components.flux_calculator_b200 as m
    return m
