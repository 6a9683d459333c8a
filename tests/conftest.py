import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


def unhex(x):
    """Inverse of float.hex() applied through nested lists."""
    if isinstance(x, str):
        return float.fromhex(x)
    if isinstance(x, list):
        return [unhex(v) for v in x]
    if isinstance(x, dict):
        return {k: unhex(v) for k, v in x.items()}
    return x


@pytest.fixture(scope="session")
def cscore_cases():
    cases = load_golden("cscore_cases.json")
    for c in cases:
        c["pwms"] = unhex(c["pwms"])
        c["cutoffs"] = unhex(c["cutoffs"])
        c["score"] = unhex(c["score"])
        c["scan"] = [[[s[0], s[1], float.fromhex(s[2]), s[3]] for s in per] for per in c["scan"]]
    return cases


@pytest.fixture(scope="session")
def scanner_toy():
    return load_golden("scanner_toy.json")
