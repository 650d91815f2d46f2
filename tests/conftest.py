"""pytest configuration: registers the `gpu` marker and puts the product package and the repo root
on sys.path. Product code lives in ``numpy-nn-model_b200/`` (import name ``neunet``); the oracle
(``oracle/``) is test infrastructure and is only ever imported from tests / smoke / bench."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "numpy-nn-model_b200")
EXAMPLES = os.path.join(ROOT, "examples")
for p in (PKG, ROOT, EXAMPLES):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture
def golden():
    return load_golden
