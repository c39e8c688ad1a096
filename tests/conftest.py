import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_engine import OracleEngine
    return OracleEngine("map-ont")


@pytest.fixture(scope="session")
def golden():
    import gzip
    import json
    with gzip.open(ROOT / "tests" / "golden" / "minimap2_fixtures.json.gz", "rt") as f:
        return json.load(f)
