import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def synth():
    import mbpol_b200
    from mbpol_b200 import synth as s
    return s


@pytest.fixture(scope="session")
def box64(synth):
    return synth.make_water_box(64)


@pytest.fixture(scope="session")
def golden_box64():
    return np.load(os.path.join(GOLDEN, "box64.npz"))


@pytest.fixture(scope="session")
def golden_clusters():
    return np.load(os.path.join(GOLDEN, "clusters.npz"))


def sort_rows(x):
    x = np.asarray(x)
    return x[np.lexsort(x.T[::-1])] if len(x) else x


def rel_rms(a, b):
    return float(np.sqrt(np.mean((np.asarray(a) - np.asarray(b)) ** 2) / np.mean(np.asarray(b) ** 2)))
