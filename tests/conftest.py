"""pytest configuration: markers and import paths.

The product package directory is ``solar-coronal-inversion_b200/`` (a path, not an importable name: it
contains a hyphen).  Exactly like the reference, which is used with its checkout root on ``sys.path``
(``import CLEDB_PROC.CLEDB_PROC as procinv``), the product is used with that directory on ``sys.path``.
"""
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, 'solar-coronal-inversion_b200')
for p in (REPO, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(REPO, 'tests', 'golden')
