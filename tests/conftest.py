import base64
import gzip
import json
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A hung kernel must fail one test, not hold the whole run: every test gets a 15-minute ceiling when the
    pytest-timeout plugin is installed (it is in this image)."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(900))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def edge_cases():
    cases = json.load(open(GOLDEN / "edge_cases.json"))
    for case in cases:
        case["text"] = base64.b64decode(case["text_b64"])
    return cases


@pytest.fixture(scope="session")
def golden_configs():
    return json.load(open(GOLDEN / "configs.json"))


@pytest.fixture(scope="session")
def reference_results():
    return json.load(open(GOLDEN / "reference_results.json"))


def read_maybe_gz(path) -> bytes:
    path = Path(path)
    return gzip.open(path, "rb").read() if path.suffix == ".gz" else path.read_bytes()
