"""Pins the CPU oracle against golden vectors produced by the reference's sqaod.py package
(tests/golden/make_golden.py) and against the known-answer cases of the reference's own unit tests.  No GPU."""
import numpy as np
import pytest

DT = [np.float32, np.float64]


def tol(dtype):
    return 1e-6 if dtype == np.float32 else 1e-12   # sqaodpy/tests/test_dense_graph_annealer.py:16


@pytest.mark.parametrize('dtype', DT)
@pytest.mark.parametrize('name', ['W8', 'Wr16', 'Wr12'])
def test_dense_formulas(oracle, golden_dense, name, dtype):
    g = golden_dense
    W = g[name]
    h, J, c = oracle.dense_graph_calculate_hamiltonian(W, dtype)
    assert np.allclose(h, g[name + '_h'], atol=tol(dtype) * 10)
    assert np.allclose(J, g[name + '_J'], atol=tol(dtype))
    assert np.allclose(c, g[name + '_c'], atol=tol(dtype) * 100)
    x = g[name + '_x']
    E = oracle.dense_graph_batch_calculate_E(W, x, dtype)
    # quantised W: every partial sum is exact in both precisions -> equality (example_problems.py:16-22)
    assert np.array_equal(E.astype(np.float64), g[name + '_E_x'])
    Eq = oracle.dense_graph_batch_calculate_E_from_spin(g[name + '_h'], g[name + '_J'], g[name + '_c'], 2 * x - 1, dtype)
    assert np.allclose(Eq, g[name + '_E_q'], rtol=tol(dtype) * 10, atol=tol(dtype) * 100)
    # QUBO energy == Ising energy (test_dense_graph_formulas.py:53-81)
    assert np.allclose(Eq, E, rtol=tol(dtype) * 10, atol=tol(dtype) * 100)
    assert E[3] == g[name + '_E_x0']


@pytest.mark.parametrize('dtype', DT)
def test_dense_bf(oracle, golden_dense, dtype):
    g = golden_dense
    for name in ('W8', 'Wr12'):
        W = g[name]; N = W.shape[0]
        for opt, tag in ((0, 'min'), (1, 'max')):
            E, xs = oracle.dense_graph_bf_search(W, opt, dtype, tile_size=1024)
            want_x = g['%s_bf_%s_x' % (name, tag)]
            assert E == g['%s_bf_%s_E' % (name, tag)][0]
            got = np.array([oracle.unpack_bits(x, N) for x in xs], np.int8)
            assert got.shape == want_x.shape and np.array_equal(got, want_x)
    # the 8x8 example has 126 degenerate argmins with E = -80 (E(k ones) = 4k^2 - 36k, k in {4,5})
    E, xs = oracle.dense_graph_bf_search(g['W8'], 0, dtype)
    assert E == -80 and len(xs) == 126
    # solution cap = tile_size (CPUDenseGraphBatchSearch.cpp:41-43): first `tile` argmins in ascending order
    E2, xs2 = oracle.dense_graph_bf_search(g['W8'], 0, dtype, tile_size=64)
    assert E2 == -80 and len(xs2) == 64


def test_dense_bf_known_answers(oracle):
    # sqaodpy/tests/test_dense_graph_bf_searcher.py:64-97
    N = 8
    for dtype in DT:
        W = np.ones((N, N))
        E, xs = oracle.dense_graph_bf_search(W, 0, dtype)
        assert E == 0 and list(xs) == [0]
        E, xs = oracle.dense_graph_bf_search(W, 1, dtype)
        assert E == N * N and list(xs) == [(1 << N) - 1]
        E, xs = oracle.dense_graph_bf_search(-W, 0, dtype)
        assert E == -N * N and list(xs) == [(1 << N) - 1]


@pytest.mark.parametrize('dtype', DT)
def test_dense_system_E(oracle, golden_dense, dtype):
    g = golden_dense
    for opt, tag in ((0, 'min'), (1, 'max')):
        ann = oracle.DenseGraphAnnealer(g['Wr16'], opt, dtype, n_trotters=6)
        ann.prepare()
        ann.set_qset(g['Wr16_sys_%s_q' % tag])
        assert np.allclose(ann.get_E(), g['Wr16_sys_%s_E' % tag], rtol=tol(dtype) * 10, atol=tol(dtype) * 100)
        want = g['Wr16_sys_%s_sysE' % tag]
        got = ann.get_system_E(0.7, 1.0 / 0.03)
        if opt == 0:
            assert abs(got - want) < 2e-5 * abs(want) if dtype == np.float32 else abs(got - want) < 1e-10
        else:
            # the C++ solvers flip the sign once more for maximize (CPUDenseGraphAnnealer.cpp:244-245);
            # sqaod.py does not (py/dense_graph_annealer.py:263-279).
            assert abs(-got - want) < 2e-5 * abs(want) if dtype == np.float32 else abs(-got - want) < 1e-10


@pytest.mark.parametrize('dtype', DT)
def test_bipartite_formulas(oracle, golden_bipartite, dtype):
    g = golden_bipartite
    b0, b1, W = g['b0'], g['b1'], g['W']
    h0, h1, J, c = oracle.bipartite_graph_calculate_hamiltonian(b0, b1, W, dtype)
    for a, b in ((h0, g['h0']), (h1, g['h1']), (J, g['J']), (c, g['c'])):
        assert np.allclose(a, b, atol=tol(dtype) * 10)
    E2d = oracle.bipartite_graph_batch_calculate_E_2d(b0, b1, W, g['x0_2d'], g['x1'], dtype)
    assert np.array_equal(E2d.astype(np.float64), g['E_2d'])
    E = oracle.bipartite_graph_batch_calculate_E(b0, b1, W, g['bx0'], g['bx1'], dtype)
    assert np.array_equal(E.astype(np.float64), g['E_x'])
    Eq = oracle.bipartite_graph_batch_calculate_E_from_spin(g['h0'], g['h1'], g['J'], g['c'],
                                                             2 * g['bx0'] - 1, 2 * g['bx1'] - 1, dtype)
    assert np.allclose(Eq, g['E_q'], rtol=tol(dtype) * 10, atol=tol(dtype) * 100)


@pytest.mark.parametrize('dtype', DT)
def test_bipartite_bf_and_system_E(oracle, golden_bipartite, dtype):
    g = golden_bipartite
    b0, b1, W = g['b0'], g['b1'], g['W']
    N1, N0 = W.shape
    for opt, tag in ((0, 'min'), (1, 'max')):
        E, pairs = oracle.bipartite_graph_bf_search(b0, b1, W, opt, dtype)
        assert E == g['bf_%s_E' % tag][0]
        got = sorted((tuple(oracle.unpack_bits(p[0], N0)), tuple(oracle.unpack_bits(p[1], N1))) for p in pairs)
        want = sorted((tuple(a), tuple(b)) for a, b in zip(g['bf_%s_x0' % tag], g['bf_%s_x1' % tag]))
        assert got == want
        ann = oracle.BipartiteGraphAnnealer(b0, b1, W, opt, dtype, n_trotters=6)
        ann.prepare()
        ann.set_qset(g['sys_%s_q0' % tag], g['sys_%s_q1' % tag])
        assert np.allclose(ann.get_E(), g['sys_%s_E' % tag], rtol=tol(dtype) * 10, atol=tol(dtype) * 100)
        got = ann.get_system_E(0.7, 1.0 / 0.03)
        want = g['sys_%s_sysE' % tag] * (1 if opt == 0 else -1)
        assert abs(got - want) < (2e-5 * abs(want) if dtype == np.float32 else 1e-10)


@pytest.mark.parametrize('dtype', DT)
def test_annealer_known_answers(oracle, dtype):
    # sqaodpy/tests/test_dense_graph_annealer.py:85-96, 154-165: E(q=-1) == 0 and W=1, q=+1 -> E == N^2
    N = 10
    W = np.ones((N, N))
    ann = oracle.DenseGraphAnnealer(W, 0, dtype, n_trotters=4)
    ann.prepare()
    ann.set_q(-np.ones(N, np.int8))
    assert np.allclose(ann.get_E(), 0, atol=tol(dtype) * 100)
    ann.set_q(np.ones(N, np.int8))
    assert np.allclose(ann.get_E(), N * N, atol=tol(dtype) * 1000)
    # bipartite: E == N0*N1 + N0 + N1 (test_bipartite_graph_annealer.py:170-182)
    N0, N1 = 6, 5
    b = oracle.BipartiteGraphAnnealer(np.ones(N0), np.ones(N1), np.ones((N1, N0)), 0, dtype, n_trotters=3)
    b.prepare()
    b.set_qset(np.ones((3, N0), np.int8), np.ones((3, N1), np.int8))
    assert np.allclose(b.get_E(), N0 * N1 + N0 + N1, atol=tol(dtype) * 1000)


@pytest.mark.parametrize('rng', ['mt', 'philox'])
@pytest.mark.parametrize('algo,m', [('coloring', 4), ('naive', 4), ('sa_naive', 1), ('sa_naive', 4)])
def test_dense_anneal_reaches_ground_state(oracle, algo, m, rng):
    # sqaodpy/tests/test_dense_graph_annealer.py:180-260: N=10, W=+-1, 100 steps G 5->0.02, beta=1/0.03
    if algo == 'naive' and rng == 'philox':
        pytest.skip('naive has no counter-based counterpart')
    N = 10
    for sign, opt in ((1.0, 0), (-1.0, 0), (1.0, 1)):
        W = sign * np.ones((N, N))
        ann = oracle.DenseGraphAnnealer(W, opt, np.float64, n_trotters=m, algorithm=algo, rng=rng)
        ann.seed(1)
        ann.prepare()
        ann.randomize_spin()
        Ginit, Gfin, beta, nsteps = (5.0, 0.02, 1. / 0.03, 100) if 'sa' not in algo else (10.0, 0.02, 1.0, 100)
        tau = (Gfin / Ginit) ** (1.0 / nsteps)
        G = Ginit
        for _ in range(nsteps):
            ann.anneal_one_step(G, beta)
            G *= tau
        E = ann.get_E()
        best = E.min() if opt == 0 else E.max()
        want = {(1.0, 0): 0.0, (-1.0, 0): -N * N, (1.0, 1): N * N}[(sign, opt)]
        assert best == want


def test_bipartite_anneal_reaches_ground_state(oracle):
    N0, N1 = 6, 5
    for algo in ('coloring', 'sa_coloring'):
        for rng in ('mt', 'philox'):
            b = oracle.BipartiteGraphAnnealer(-np.ones(N0), -np.ones(N1), -np.ones((N1, N0)), 0, np.float64,
                                              n_trotters=4, algorithm=algo, rng=rng)
            b.seed(3)
            b.prepare()
            b.randomize_spin()
            G = 5.0 if algo == 'coloring' else 10.0
            beta = 1. / 0.03 if algo == 'coloring' else 1.0
            for _ in range(100):
                b.anneal_one_step(G, beta)
                G *= (0.02 / 5.0) ** 0.01
            assert b.get_E().min() == -(N0 * N1 + N0 + N1)


# ---- chains recorded from the reference's own compiled CPU annealers (tests/golden/make_golden_refcpu.py, one worker): the oracle's
# MT19937 mode must walk through exactly the same spins, step after step
def _chain_keys(prefix):
    import os
    f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'refcpu_chains.npz'))
    return sorted({k.split('/')[0] for k in f.files if k.startswith(prefix)}, key=lambda s: int(s[len(prefix):]))


@pytest.fixture(scope='module')
def refcpu_chains():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'refcpu_chains.npz'))


@pytest.mark.parametrize('key', _chain_keys('dense'))
def test_dense_chain_equals_the_compiled_reference(oracle, refcpu_chains, key):
    g = refcpu_chains
    N, m, seed, width = (int(v) for v in g[key + '/meta'])
    dtype = np.float32 if width == 4 else np.float64
    ann = oracle.DenseGraphAnnealer(None, 0, dtype, n_trotters=m, algorithm=str(g[key + '/algo']), n_workers=1, rng='mt')
    ann.set_hamiltonian(g[key + '/h'], g[key + '/J'], dtype(g[key + '/c']))
    ann.seed(seed); ann.prepare(); ann.randomize_spin()
    q = g[key + '/q']
    assert np.array_equal(ann.get_q(), q[0]), 'randomize_spin'
    for k, G in enumerate(g[key + '/G']):
        ann.anneal_one_step(float(G), 1. / 0.02)
        assert np.array_equal(ann.get_q(), q[k + 1]), 'step %d' % k
    assert np.allclose(ann.get_E(), g[key + '/E'], rtol=2e-5 if width == 4 else 1e-12, atol=1e-4 if width == 4 else 1e-10)


@pytest.mark.parametrize('key', _chain_keys('bip'))
def test_bipartite_chain_equals_the_compiled_reference(oracle, refcpu_chains, key):
    g = refcpu_chains
    N0, N1, m, seed, width = (int(v) for v in g[key + '/meta'])
    dtype = np.float32 if width == 4 else np.float64
    ann = oracle.BipartiteGraphAnnealer(g[key + '/b0'], g[key + '/b1'], g[key + '/W'], 0, dtype, n_trotters=m,
                                        algorithm=str(g[key + '/algo']), n_workers=1, rng='mt')
    ann.seed(seed); ann.prepare(); ann.randomize_spin()
    q0, q1 = g[key + '/q0'], g[key + '/q1']
    a, b = ann.get_q()
    assert np.array_equal(a, q0[0]) and np.array_equal(b, q1[0]), 'randomize_spin'
    for k, G in enumerate(g[key + '/G']):
        ann.anneal_one_step(float(G), 1. / 0.02)
        a, b = ann.get_q()
        assert np.array_equal(a, q0[k + 1]) and np.array_equal(b, q1[k + 1]), 'step %d' % k
    assert np.array_equal(ann.get_E(), g[key + '/E'])      # quantised inputs: exact


def test_config_c1_chain_equals_the_compiled_reference(oracle, refcpu_chains):
    """BASELINE.json configs[0]: the reference's tutorial anneal (N = 128, m = 32, fp64, G 5 -> 0.01 with G *= 0.99, 619 steps,
    sqaodpy/example/dense_graph_annealer.py:22-70) as sqaod.cpu ran it -- 2.5 million attempts later the oracle holds the same spins."""
    g = refcpu_chains
    ann = oracle.DenseGraphAnnealer(None, 0, np.float64, n_trotters=32, algorithm='coloring', n_workers=1, rng='mt')
    ann.set_hamiltonian(g['c1/h'], g['c1/J'], np.float64(g['c1/c']))
    ann.seed(13255); ann.prepare(); ann.randomize_spin()
    at = [int(v) for v in g['c1/at']]
    G, k, i = 5.0, 0, 0
    while 0.01 <= G:
        ann.anneal_one_step(G, 1. / 0.02)
        G *= 0.99
        k += 1
        if k % 124 == 0:
            assert at[i] == k and np.array_equal(ann.get_q(), g['c1/q'][i]), 'step %d' % k
            i += 1
    assert k == int(g['c1/steps']) == 619
    assert np.array_equal(ann.get_q(), g['c1/q'][-1])
    assert np.allclose(ann.get_E(), g['c1/E'], rtol=1e-12, atol=1e-10)
