"""pytest configuration: registers the `gpu` marker and exposes the golden fixtures."""

import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


def _load(name):
    return np.load(GOLDEN / name, allow_pickle=False)


@pytest.fixture(scope="session")
def ref_fields():
    return _load("ref_fields.npz"), json.load(open(GOLDEN / "ref_fields_meta.json"))


@pytest.fixture(scope="session")
def ref_rs():
    return _load("ref_rs.npz")


@pytest.fixture(scope="session")
def live_elementwise():
    return _load("live_elementwise.npz"), json.load(open(GOLDEN / "live_elementwise_meta.json"))


@pytest.fixture(scope="session")
def live_gf128():
    return _load("live_gf2_128.npz")


@pytest.fixture(scope="session")
def live_ntt():
    return _load("live_ntt.npz"), json.load(open(GOLDEN / "live_ntt_meta.json"))


@pytest.fixture(scope="session")
def live_linalg():
    return _load("live_linalg.npz")


@pytest.fixture(scope="session")
def live_rs():
    return _load("live_rs.npz"), json.load(open(GOLDEN / "live_rs_meta.json"))


@pytest.fixture(scope="session")
def ref_bch():
    return _load("ref_bch.npz")


@pytest.fixture(scope="session")
def live_bch():
    return _load("live_bch.npz"), json.load(open(GOLDEN / "live_bch_meta.json"))


@pytest.fixture(scope="session")
def ref_linalg():
    return _load("ref_linalg.npz")


@pytest.fixture(scope="session")
def ref_roots():
    return _load("ref_roots.npz")


@pytest.fixture(scope="session")
def live_next():
    return _load("live_next.npz"), json.load(open(GOLDEN / "live_next_meta.json"))
