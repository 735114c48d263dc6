"""GPU parity tests of the brute-force searchers (through the C ABI).
Bar: bit-exact minimum energy and solution sets against the reference CPU searcher (restated in oracle/) on
inputs whose sums are exact (integers, 2^-14 grid), plus the reference's golden vectors and known-answer tests."""
import numpy as np
import pytest
from conftest import quantized_symmetric_W, quantized_bipartite

pytestmark = pytest.mark.gpu
DT = [np.float32, np.float64]


@pytest.fixture(scope='module')
def sq():
    import sqaod_b200
    return sqaod_b200


def packed(bits):
    v = 0
    for b in bits:
        v = (v << 1) | int(b)
    return v


@pytest.mark.parametrize('dtype', DT)
def test_dense_bf_golden(sq, golden_dense, dtype):
    g = golden_dense
    for name in ('W8', 'Wr12'):
        for opt, tag in ((sq.minimize, 'min'), (sq.maximize, 'max')):
            s = sq.dense_graph_bf_searcher(g[name], opt, dtype)
            s.search()
            E, x = s.get_E(), np.stack(s.get_x())
            want_x = g['%s_bf_%s_x' % (name, tag)]
            assert np.all(E == g['%s_bf_%s_E' % (name, tag)][0]) and len(E) == len(want_x)
            assert np.array_equal(x, want_x)          # ascending order, like the CPU searcher


@pytest.mark.parametrize('dtype', DT)
def test_dense_bf_known_answers(sq, dtype):
    # sqaodpy/tests/test_dense_graph_bf_searcher.py:58-107
    N = 8
    W = np.ones((N, N))
    s = sq.dense_graph_bf_searcher(W, sq.minimize, dtype); s.search()
    assert s.get_E()[0] == 0 and np.array_equal(s.get_x()[0], np.zeros(N, np.int8))
    s = sq.dense_graph_bf_searcher(W, sq.maximize, dtype); s.search()
    assert s.get_E()[0] == N * N and np.array_equal(s.get_x()[0], np.ones(N, np.int8))
    s = sq.dense_graph_bf_searcher(-W, sq.minimize, dtype); s.search()
    assert s.get_E()[0] == -N * N and np.array_equal(s.get_x()[0], np.ones(N, np.int8))
    with pytest.raises(RuntimeError):
        sq.dense_graph_bf_searcher(np.ones((100, 100)), sq.minimize, dtype)
    sq.dense_graph_bf_searcher(np.ones((63, 63)), sq.minimize, dtype)   # N = 63 is accepted
    p = s.get_preferences()
    assert p['algorithm'] == 'brute_force_search' and p['device'] == 'cuda'


@pytest.mark.parametrize('dtype', DT)
@pytest.mark.parametrize('N', [1, 3, 9, 16, 20, 23])
def test_dense_bf_vs_oracle(sq, oracle, N, dtype):
    W = quantized_symmetric_W(N, 77 + N, dtype)
    for opt in (0, 1):
        E0, xs0 = oracle.dense_graph_bf_search(W, opt, dtype, tile_size=1 << min(N, 16))
        s = sq.dense_graph_bf_searcher(W, sq.maximize if opt else sq.minimize, dtype)
        s.search()
        assert s.get_E()[0] == E0
        got = sorted(packed(x) for x in s.get_x())
        assert got == [int(v) for v in xs0]


def test_dense_bf_tiles_and_ranges(sq, oracle):
    """stepwise searchRange with small tiles (unaligned to the kernel's row size) covers the range exactly once."""
    N = 18
    W = quantized_symmetric_W(N, 5, np.float64)
    E0, xs0 = oracle.dense_graph_bf_search(W, 0, np.float64, tile_size=1 << 16)
    for tile in (256, 768, 10240, 1 << 17):
        s = sq.dense_graph_bf_searcher(W, sq.minimize, np.float64, tile_size=tile)
        s.prepare()
        last = 0
        while True:
            done, cur = s.search_range()
            assert cur > last or done
            last = cur
            if done:
                break
        assert last == 1 << N
        s.make_solution()
        assert s.get_E()[0] == E0 and sorted(packed(x) for x in s.get_x()) == [int(v) for v in xs0]
    # sharded: two half ranges, then merge
    halves = []
    for b, e in ((0, 100000), (100000, 1 << N)):
        s = sq.dense_graph_bf_searcher(W, sq.minimize, np.float64)
        s.set_range(b, e)
        s.prepare()
        while not s.search_range()[0]:
            pass
        halves.append((s.get_Emin(), list(s.get_packed_x())))
    Emin = min(h[0] for h in halves)
    merged = sorted(int(x) for h in halves if h[0] == Emin for x in h[1])
    assert Emin == E0 and merged == [int(v) for v in xs0]


def test_dense_bf_degenerate(sq):
    # 8x8 example: E(k ones) = 4k^2 - 36k -> 126 argmins at E = -80 (example_problems.py:4-14)
    W = np.full((8, 8), 4.0); np.fill_diagonal(W, -32.0)
    s = sq.dense_graph_bf_searcher(W, sq.minimize, np.float32); s.search()
    assert len(s.get_x()) == 126 and np.all(s.get_E() == -80)
    # every state ties: the list is the first `cap` states in ascending order
    s = sq.dense_graph_bf_searcher(np.zeros((12, 12)), sq.minimize, np.float64, tile_size=256); s.search()
    xs = [packed(x) for x in s.get_x()]
    assert xs == list(range(256)) and np.all(s.get_E() == 0)


@pytest.mark.parametrize('dtype', DT)
def test_bipartite_bf_golden(sq, golden_bipartite, dtype):
    g = golden_bipartite
    for opt, tag in ((sq.minimize, 'min'), (sq.maximize, 'max')):
        s = sq.bipartite_graph_bf_searcher(g['b0'], g['b1'], g['W'], opt, dtype)
        s.search()
        assert np.all(s.get_E() == g['bf_%s_E' % tag][0])
        got = sorted((tuple(a), tuple(b)) for a, b in s.get_x())
        want = sorted((tuple(a), tuple(b)) for a, b in zip(g['bf_%s_x0' % tag], g['bf_%s_x1' % tag]))
        assert got == want


@pytest.mark.parametrize('dtype', DT)
@pytest.mark.parametrize('N0,N1', [(1, 1), (4, 7), (10, 9), (12, 3)])
def test_bipartite_bf_vs_oracle(sq, oracle, N0, N1, dtype):
    b0, b1, W = quantized_bipartite(N0, N1, 31 + N0, dtype)
    for opt in (0, 1):
        E0, pairs0 = oracle.bipartite_graph_bf_search(b0, b1, W, opt, dtype)
        for t0, t1 in ((1 << 15, 1 << 15), (256, 256)):
            s = sq.bipartite_graph_bf_searcher(b0, b1, W, sq.maximize if opt else sq.minimize, dtype, tile_size_0=t0, tile_size_1=t1)
            s.search()
            assert s.get_E()[0] == E0
            got = sorted((packed(a), packed(b)) for a, b in s.get_x())
            assert got == sorted(pairs0)
    # W = 1: E == N0*N1 + N0 + N1 at x = all ones for maximize (test_bipartite_graph_bf_searcher.py)
    s = sq.bipartite_graph_bf_searcher(np.ones(N0), np.ones(N1), np.ones((N1, N0)), sq.maximize, dtype)
    s.search()
    assert s.get_E()[0] == N0 * N1 + N0 + N1


@pytest.mark.parametrize('dtype', DT)
def test_massively_degenerate_problem_returns_the_lowest_states(sq, dtype):
    """W = 0: every one of the 2^N states is an argmin -- far more ties than the gather buffer holds even inside one kernel span.
    The solution list must be the `cap` LOWEST packed states in ascending order (CPU searcher semantics,
    CPUDenseGraphBFSearcher.cpp:103-131), deterministically."""
    N = 24
    W = np.zeros((N, N), dtype)
    s = sq.dense_graph_bf_searcher(W, sq.minimize, dtype)
    s.search()
    assert float(s.get_E()[0]) == 0.0
    xs = s.get_packed_x()
    assert len(xs) == 65536
    assert np.array_equal(xs, np.arange(65536, dtype=np.uint64))
    s2 = sq.dense_graph_bf_searcher(W, sq.minimize, dtype, tile_size=1 << 12)     # the cap follows the tile size
    s2.search()
    assert np.array_equal(s2.get_packed_x(), np.arange(1 << 12, dtype=np.uint64))
