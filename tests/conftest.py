import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def load_golden(name):
    import torch
    return torch.load(GOLDEN / f"{name}.pt", map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def golden_names():
    return sorted(p.stem for p in GOLDEN.glob("*.pt"))
