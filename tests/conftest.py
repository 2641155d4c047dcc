"""pytest configuration: the `gpu` marker (tests that need a B200) and shared fixtures."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))

from __graft_entry__ import load_oracle, load_package  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: multi-GiB cases (deselect with -m 'gpu and not slow' for a quick pass)")


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def zo():
    return load_oracle()


@pytest.fixture(scope="session")
def golden():
    def load(name):
        with open(os.path.join(ROOT, "tests", "golden", name)) as f:
            return json.load(f)
    return load


@pytest.fixture(scope="session")
def ctx(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    c = pkg.Context(0)
    yield c
    c.close()
