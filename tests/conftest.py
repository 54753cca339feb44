"""pytest configuration: GPU marker, import paths, shared fixtures."""
import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ga-dpamsa_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
os.environ.setdefault("GADPAMSA_PROJECT_ROOT", os.path.join("/tmp", "gadpamsa_project_root"))
for p in (PKG, os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_json_gz(name):
    with gzip.open(os.path.join(GOLDEN, name), "rt") as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def report_vectors():
    return load_json_gz("report_vectors.json.gz")


@pytest.fixture(scope="session")
def primitives():
    return load_json_gz("primitives.json.gz")


@pytest.fixture(scope="session")
def ga_traces():
    return load_json_gz("ga_traces.json.gz")


@pytest.fixture(scope="session")
def datasets_json():
    with open(os.path.join(GOLDEN, "datasets.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def native_lib():
    import __graft_entry__ as entry
    entry.build()
    import _native
    return _native
