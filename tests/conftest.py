import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a minute on the GPU box")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device here")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def O():
    """The oracle (checker).  tests/ is one of the few places allowed to import it."""
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def J():
    import jmmonedmc_b200
    jmmonedmc_b200.build()
    return jmmonedmc_b200


def golden(name):
    d = GOLDEN / name
    out = {"dir": d, "deck_text": (d / "INPUT").read_text(), "summary": json.loads((d / "summary.json").read_text())}
    if (d / "rng.u32").exists():
        out["rng"] = np.fromfile(d / "rng.u32", dtype=np.uint32)
    return out


@pytest.fixture(scope="session")
def gold():
    return golden
