"""pytest configuration: registers the ``gpu`` marker and puts the repo root (for ``oracle``)
and ``deep-turbulence_b200`` (for the importable ``tmglow_b200`` package) on sys.path."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "deep-turbulence_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=True)


@pytest.fixture(scope="session")
def golden():
    return load_golden
