import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def match_golden():
    return load_golden("match_golden.json")["cases"]


@pytest.fixture(scope="session")
def stream_golden():
    return load_golden("stream_golden.json")["cases"]


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu but no CUDA device is visible")
    import tvidz_b200._lib as L
    L.lib()     # raises loudly if the CUDA library is missing: there is no fallback
    return torch.device("cuda:0")
