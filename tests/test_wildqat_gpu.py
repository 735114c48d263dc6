"""wildqat-style adapter (sqaod_b200.wildqat.opt, counterpart of sqaodpy/sqaod/wildqat/opt.py:5-67): SA and SQA over a small
upper-triangular QUBO reach the brute-force minimum and fill the energy history the way the reference adapter does."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_wildqat_adapter_reaches_the_minimum():
    import sqaod_b200 as sq
    import sqaod_b200.wildqat as wq
    rng = np.random.default_rng(3)
    N = 12
    Q = np.triu(np.rint((rng.random((N, N)) - 0.5) * 64) / 8.)          # upper triangular, as wildqat users write it
    Wsym = (Q + Q.T) / 2
    bf = sq.dense_graph_bf_searcher(Wsym, sq.minimize, np.float32)
    bf.search()
    Emin = float(bf.get_E()[0])
    a = wq.opt()
    a.qubo = Q.tolist()
    a.ite = 2000
    best = np.inf
    for _ in range(4):
        x = np.asarray(a.sa(), np.float64)
        assert x.shape == (N,) and set(np.unique(x)) <= {0., 1.}
        best = min(best, float(x @ Q @ x))
        assert len(a.E) > 10
    assert abs(best - Emin) < 1e-5
    a.tro = 6
    xs = a.sqa()
    assert len(xs) == 6 and len(a.E) > 10
    assert min(float(np.asarray(x, np.float64) @ Q @ np.asarray(x, np.float64)) for x in xs) <= Emin + 1e-5 + 0.5 * abs(Emin)
