import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden_dense():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'dense_graph.npz'))


@pytest.fixture(scope='session')
def golden_bipartite():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'bipartite_graph.npz'))


@pytest.fixture(scope='session')
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


def quantized_symmetric_W(N, seed, dtype=np.float64):
    """sqaodpy/tests/example_problems.py:16-22 with a seeded generator: symmetric U(-0.5,0.5) on the 2^-14 grid."""
    rng = np.random.default_rng(seed)
    A = rng.random((N, N)) - 0.5
    W = np.triu(A) + np.triu(A, 1).T
    return np.asarray(np.rint(W * 16384) / 16384., dtype)


def quantized_bipartite(N0, N1, seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    q = lambda a: np.asarray(np.rint(a * 16384) / 16384., dtype)
    return q(rng.random(N0) - 0.5), q(rng.random(N1) - 0.5), q(rng.random((N1, N0)) - 0.5)
