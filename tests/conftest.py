import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def rg_fixtures():
    return json.loads((ROOT / "tests" / "golden" / "rappe_goddard.json").read_text())


@pytest.fixture(scope="session")
def sto_truth():
    return json.loads((ROOT / "tests" / "golden" / "sto_truth.json").read_text())


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def gpu_ctx():
    """One engine context on cuda:0.  Fails (not skips) if the CUDA extension cannot run: GPU tests must not
    pass on a fallback."""
    import cheq_b200
    return cheq_b200.default_context(0)


def rg_cases(fixtures):
    for g in fixtures["groups"]:
        for c in g["cases"]:
            yield g, c
