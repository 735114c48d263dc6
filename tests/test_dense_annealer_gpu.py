"""GPU parity tests of the dense-graph annealer, through the C ABI (sqaod_b200 -> libsqaod_b200.so).

Bar (BASELINE.json north_star): energies of identical spin configurations within 1e-5 relative (fp32) / 1e-12 (fp64);
the sweep itself is checked two ways: (1) exact-chain -- the kernel's trajectory equals the CPU oracle's when the oracle
draws from the same Philox stream (bit-exact spins); (2) statistically against the reference's own MT19937 chain
(tests/test_annealer_statistics_gpu.py)."""
import numpy as np
import pytest
from conftest import quantized_symmetric_W

pytestmark = pytest.mark.gpu
DT = [np.float32, np.float64]


def tol(dtype):
    return 1e-5 if dtype == np.float32 else 1e-12


@pytest.fixture(scope='module')
def sq():
    import sqaod_b200
    return sqaod_b200


@pytest.mark.parametrize('dtype', DT)
def test_hamiltonian_and_energy_vs_golden(sq, oracle, golden_dense, dtype):
    g = golden_dense
    for name in ('W8', 'Wr16', 'Wr12'):
        W = g[name]
        ann = sq.dense_graph_annealer(W, sq.minimize, dtype, n_trotters=4)
        h, J, c = ann.get_hamiltonian()
        assert np.allclose(h, g[name + '_h'], atol=tol(dtype) * 10)
        assert np.allclose(J, g[name + '_J'], atol=tol(dtype))
        assert abs(c - g[name + '_c']) <= tol(dtype) * 100
        x = g[name + '_x']
        q = (2 * x - 1).astype(np.int8)
        ann.set_qset(q)                      # all 2^N configurations at once (test_dense_graph_annealer.py:98-120)
        E = ann.get_E()
        want = g[name + '_E_q']
        assert np.allclose(E, want, rtol=tol(dtype), atol=tol(dtype) * 10)
        # formulas object
        Ex = sq.formulas.dense_graph_batch_calculate_E(W, x, dtype)
        assert np.array_equal(Ex.astype(np.float64), g[name + '_E_x'])      # quantised W: exact
        h2, J2, c2 = sq.formulas.dense_graph_calculate_hamiltonian(W, dtype)
        assert np.allclose(h2, g[name + '_h'], atol=tol(dtype) * 10) and np.allclose(J2, g[name + '_J'], atol=tol(dtype))
        Eq = sq.formulas.dense_graph_batch_calculate_E_from_spin(g[name + '_h'], g[name + '_J'], g[name + '_c'], q, dtype)
        assert np.allclose(Eq, want, rtol=tol(dtype), atol=tol(dtype) * 10)


@pytest.mark.parametrize('dtype', DT)
def test_known_answers(sq, dtype):
    N = 10
    W = np.ones((N, N))
    ann = sq.dense_graph_annealer(W, sq.minimize, dtype, n_trotters=4)
    ann.prepare()
    ann.set_q(-np.ones(N, np.int8))
    assert np.allclose(ann.get_E(), 0, atol=1e-5)
    ann.set_q(np.ones(N, np.int8))
    assert np.allclose(ann.get_E(), N * N, atol=1e-4)
    assert all(np.array_equal(x, np.ones(N, np.int8)) for x in ann.get_x())
    with pytest.raises(RuntimeError):        # anneal before prepare / q set (test_dense_graph_annealer.py:390-393)
        a2 = sq.dense_graph_annealer(W, sq.minimize, dtype)
        a2.anneal_one_step(1., 1.)
    p = ann.get_preferences()
    assert p['algorithm'] == 'coloring' and p['n_trotters'] == 4 and p['device'] == 'cuda'
    assert p['precision'] == ('float' if dtype == np.float32 else 'double')
    ann.set_preferences(algorithm='naive')   # unsupported -> default (CUDA table, test_dense_graph_annealer.py:454-482)
    assert ann.get_preferences()['algorithm'] == 'coloring'
    ann.set_preferences(algorithm='sa_default')
    assert ann.get_preferences()['algorithm'] == 'sa_naive'


@pytest.mark.parametrize('dtype', DT)
def test_system_E_vs_golden(sq, golden_dense, dtype):
    g = golden_dense
    for opt, tag in ((sq.minimize, 'min'), (sq.maximize, 'max')):
        ann = sq.dense_graph_annealer(g['Wr16'], opt, dtype, n_trotters=6)
        ann.prepare()
        ann.set_qset(g['Wr16_sys_%s_q' % tag])
        assert np.allclose(ann.get_E(), g['Wr16_sys_%s_E' % tag], rtol=tol(dtype), atol=tol(dtype) * 10)
        want = g['Wr16_sys_%s_sysE' % tag] * (1 if tag == 'min' else -1)   # C++ solvers flip once more for maximize
        got = ann.get_system_E(0.7, 1. / 0.03)
        assert abs(got - want) <= (2e-5 if dtype == np.float32 else 1e-10) * max(1., abs(want))


CHAIN_CASES = [
    # N, m, algorithm, steps
    (10, 4, 'coloring', 6),
    (40, 20, 'coloring', 3),
    (33, 7, 'coloring', 4),        # odd ring: trotter m-1 is its own phase
    (16, 2, 'coloring', 4),        # both neighbours are the same trotter
    (100, 151, 'coloring', 2),     # more trotters than CTAs -> mixed T, odd m
    (300, 296, 'coloring', 1),     # T = 2 everywhere, every CTA has remote neighbours
    (1000, 32, 'coloring', 1),
    (2100, 12, 'coloring', 1),     # more than one 2048-spin super-block, several chunks per row
    (64, 512, 'coloring', 2),      # the C2 trotter layout: 148 CTAs with 4 or 3 trotters each
    (128, 601, 'coloring', 1),     # 5 / 4 trotters per CTA, odd ring
    (4500, 6, 'coloring', 1),      # rows longer than a super-block, partial last chunk
    (72, 1800, 'coloring', 1),     # 13 / 12 trotters per CTA: the wide warp layout (14 streaming warps, one helper warp)
    (40, 4001, 'coloring', 1),     # 28 / 27 trotters per CTA, K shrinks, odd ring, wide layout
    (24, 1, 'sa_naive', 5),
    (50, 9, 'sa_naive', 3),
]


def field_mode_fits(N, m, dtype):
    """mirror of prepare()'s plan: T rows of roundUp(N,128) fields next to ~40 KB of tables in 227 KB of shared memory"""
    T = -(-m // min(148, m))
    return T * (-(-N // 128) * 128) * np.dtype(dtype).itemsize <= 180 * 1024


@pytest.mark.parametrize('dtype', DT)
@pytest.mark.parametrize('mode', ['classic', 'field', 'field_writeback'])
@pytest.mark.parametrize('N,m,algo,steps', CHAIN_CASES)
def test_exact_chain_vs_oracle(sq, oracle, N, m, algo, steps, dtype, mode):
    """The kernel and the oracle (philox mode) must produce identical spins step by step -- in both sweep modes: 'classic'
    (one J row per attempt), 'field' (local fields from the spin GEMM at step start, updated per accepted flip) and
    'field_writeback' (fields carried from step to step in global memory, recomputed only every 1000 steps)."""
    if mode != 'classic' and not field_mode_fits(N, m, dtype):
        pytest.skip('field rows do not fit in shared memory for this shape')
    W = quantized_symmetric_W(N, 1000 + N, dtype)
    for seed in range(5, 25):
        ref = oracle.DenseGraphAnnealer(W, 0, dtype, n_trotters=m, algorithm=algo, rng='philox')
        ref.seed(seed); ref.prepare(); ref.randomize_spin()
        ann = sq.dense_graph_annealer(W, sq.minimize, dtype, n_trotters=m, algorithm=algo)
        ann.set_sweep_mode('classic' if mode == 'classic' else 'field', 1000 if mode == 'field_writeback' else 0)
        ann.seed(seed); ann.prepare(); ann.randomize_spin()
        assert ann.get_sweep_mode() == ('classic' if mode == 'classic' else 'field')
        assert np.array_equal(ann.get_spins(), ref.get_q()), 'randomize_spin stream differs'
        G, beta = (3.0, 1. / 0.3) if algo == 'coloring' else (2.0, 1.0)
        traj_ok = True
        for s in range(steps):
            ref.anneal_one_step(G, beta)
            ann.anneal_one_step(G, beta)
            G *= 0.7
            got, want = ann.get_spins(), ref.get_q()
            if not np.array_equal(got, want):
                # only an accept test that sat on the rounding edge may make the trajectories part: try the next seed
                assert ref.stats()[1] > 0, 'step %d: %d spins differ (seed %d)' % (s, int((got != want).sum()), seed)
                traj_ok = False
                break
        if traj_ok:
            assert ref.stats()[0] > 0          # the chain did move
            assert ann.get_stats()['accepted'] == ref.stats()[0]
            assert np.allclose(ann.get_E(), ref.get_E(), rtol=tol(dtype), atol=tol(dtype) * 10)
            return
    pytest.fail('no seed without a borderline accept test')


@pytest.mark.parametrize('dtype', DT)
@pytest.mark.parametrize('algo,m', [('coloring', 4), ('default', 4), ('sa_naive', 1), ('sa_naive', 4), ('coloring', 1)])
def test_reaches_ground_state(sq, algo, m, dtype):
    # sqaodpy/tests/test_dense_graph_annealer.py:180-260
    N = 10
    for sign, opt, want in ((1.0, sq.minimize, 0.0), (-1.0, sq.minimize, -N * N), (1.0, sq.maximize, N * N)):
        ann = sq.dense_graph_annealer(sign * np.ones((N, N)), opt, dtype, n_trotters=m, algorithm=algo)
        ann.seed(11); ann.prepare(); ann.randomize_spin()
        sa = not sq.algorithm.is_sqa(ann.get_preferences()['algorithm'])
        G, Gfin, beta = (10.0, 0.02, 1.0) if sa else (5.0, 0.02, 1. / 0.03)
        tau = (Gfin / G) ** 0.01
        for _ in range(100):
            ann.anneal_one_step(G, beta)
            G *= tau
        ann.make_solution()
        E = ann.get_E()
        assert (E.min() if opt is sq.minimize else E.max()) == want


def test_set_q_get_q_roundtrip(sq):
    # sqaodc/tests/CUDADenseGraphAnnealerTest.cpp:191-247
    N, m = 40, 20
    rng = np.random.default_rng(0)
    ann = sq.dense_graph_annealer(quantized_symmetric_W(N, 3), sq.minimize, np.float32, n_trotters=m)
    ann.prepare()
    q = (2 * rng.integers(0, 2, N) - 1).astype(np.int8)
    ann.set_q(q)
    assert all(np.array_equal(v, q) for v in ann.get_q()) and len(ann.get_q()) == m
    qs = (2 * rng.integers(0, 2, (m + 3, N)) - 1).astype(np.int8)
    ann.set_qset(qs)                                   # also changes the number of trotters
    assert ann.get_preferences()['n_trotters'] == m + 3
    assert np.array_equal(np.stack(ann.get_q()), qs)
    assert np.array_equal(np.stack(ann.get_x()), (qs + 1) // 2)


@pytest.mark.parametrize('dtype', DT)
@pytest.mark.parametrize('N,m,R,algo', [(64, 8, 5, 'coloring'), (200, 70, 40, 'coloring'), (96, 1, 7, 'sa_naive'), (128, 33, 200, 'coloring')])
def test_replica_batch_equals_separate_solvers(sq, N, m, R, algo, dtype):
    """R replicas annealed side by side in one launch == R separate solvers seeded seed + r (spin for spin)."""
    W = quantized_symmetric_W(N, 31 + N, dtype)
    seed = 100
    batch = sq.dense_graph_annealer(W, sq.minimize, dtype, n_trotters=m, algorithm=algo)
    batch.set_replicas(R)
    batch.seed(seed); batch.prepare(); batch.randomize_spin()
    Gs = [2.0, 1.0, 0.4]
    for G in Gs:
        batch.anneal_one_step(G, 1. / 0.3)
    qb = batch.get_spins().reshape(R, m, N)
    Eb = batch.get_E().reshape(R, m)
    for r in sorted(set([0, 1, R // 2, R - 1])):
        one = sq.dense_graph_annealer(W, sq.minimize, dtype, n_trotters=m, algorithm=algo)
        one.seed(seed + r); one.prepare(); one.randomize_spin()
        for G in Gs:
            one.anneal_one_step(G, 1. / 0.3)
        assert np.array_equal(one.get_spins(), qb[r]), 'replica %d differs' % r
        assert np.allclose(one.get_E(), Eb[r], rtol=tol(dtype), atol=tol(dtype) * 10)
    assert len(batch.get_q()) == R * m


@pytest.mark.parametrize('dtype', DT)
def test_problem_batch_equals_separate_solvers(sq, dtype):
    """set_qubo_batch: R different problems in one launch per step; problem r must follow exactly the trajectory of a solver of
    its own with seed + r."""
    R, N, m, steps = 7, 72, 10, 3
    Ws = np.stack([quantized_symmetric_W(N, 7000 + r, dtype) for r in range(R)])
    batch = sq.dense_graph_annealer(None, sq.minimize, dtype, n_trotters=m)
    batch.set_qubo_batch(Ws)
    batch.set_preferences(n_trotters=m)
    batch.seed(31); batch.prepare(); batch.randomize_spin()
    singles = []
    for r in range(R):
        a = sq.dense_graph_annealer(Ws[r], sq.minimize, dtype, n_trotters=m)
        a.seed(31 + r); a.prepare(); a.randomize_spin()
        singles.append(a)
    q = batch.get_spins()
    assert q.shape == (R * m, N)
    G, beta = 2.0, 3.0
    for s in range(steps):
        batch.anneal_one_step(G, beta)
        for a in singles:
            a.anneal_one_step(G, beta)
        G *= 0.6
        q = batch.get_spins()
        for r, a in enumerate(singles):
            assert np.array_equal(q[r * m:(r + 1) * m], a.get_spins()), 'problem %d, step %d' % (r, s)
    E = batch.get_E()
    for r, a in enumerate(singles):
        assert np.allclose(E[r * m:(r + 1) * m], a.get_E(), rtol=tol(dtype), atol=tol(dtype) * 10)
    assert len(np.unique(np.round(E, 3))) > R          # the problems really differ


FIELD_CASES = [
    # N, m, algorithm, steps, G0   (shapes where rows are long enough that the field update spans many column groups)
    (1500, 9, 'coloring', 3, 3.0),      # 12 column groups over 12 dot warps, odd ring
    (3000, 40, 'coloring', 2, 2.0),     # two super-blocks, several groups per warp
    (8192, 8, 'coloring', 1, 1.0),      # the C2 row length
    (2048, 300, 'coloring', 1, 0.5),    # 3 / 2 trotters per CTA: wide layout next to the chain-critical one
    (700, 5, 'sa_naive', 3, 2.0),
]


@pytest.mark.parametrize('dtype', DT)
@pytest.mark.parametrize('N,m,algo,steps,G0', FIELD_CASES)
def test_field_mode_equals_classic_mode(sq, N, m, algo, steps, G0, dtype):
    """Field mode and classic mode run the same Markov chain: identical spins after every step on a quantised problem
    (every sum exact), with the fields carried across steps (write-back) and with a per-step recomputation."""
    W = quantized_symmetric_W(N, 4000 + N, dtype)
    anns = []
    for mode, refresh in (('classic', 0), ('field', 1), ('field', 1000)):
        a = sq.dense_graph_annealer(W, sq.minimize, dtype, n_trotters=m, algorithm=algo)
        a.set_sweep_mode(mode, refresh)
        a.seed(77); a.prepare(); a.randomize_spin()
        assert a.get_sweep_mode() == mode
        anns.append(a)
    G, beta = G0, 1. / 0.3
    for s in range(steps):
        for a in anns:
            a.anneal_one_step(G, beta)
        G *= 0.6
        q0 = anns[0].get_spins()
        for k, a in enumerate(anns[1:]):
            qa = a.get_spins()
            assert np.array_equal(qa, q0), 'step %d, variant %d: %d spins differ' % (s, k + 1, int((qa != q0).sum()))
    acc = [a.get_stats()['accepted'] for a in anns]
    assert acc[0] > 0 and acc[0] == acc[1] == acc[2]
    # spins written from outside invalidate the carried fields
    a = anns[2]
    rng = np.random.default_rng(5)
    qs = (2 * rng.integers(0, 2, (m, N)) - 1).astype(np.int8)
    for b in (anns[0], a):
        b.set_qset(qs)
        b.anneal_one_step(G, beta)
    assert np.array_equal(anns[0].get_spins(), a.get_spins())


def test_automatic_sweep_mode(sq):
    # field mode whenever the field rows fit in shared memory (prepare()'s measured rule) ...
    a = sq.dense_graph_annealer(quantized_symmetric_W(192, 5), sq.minimize, np.float32, n_trotters=512)
    a.prepare()
    assert a.get_sweep_mode() == 'field'
    a = sq.dense_graph_annealer(quantized_symmetric_W(256, 5), sq.minimize, np.float32, n_trotters=64)
    a.prepare()
    assert a.get_sweep_mode() == 'field'
    a.set_sweep_mode('classic')
    a.prepare()
    assert a.get_sweep_mode() == 'classic'
    # ... and with a replica batch, in a geometry of at most four trotters per CTA
    r = sq.dense_graph_annealer(quantized_symmetric_W(256, 5), sq.minimize, np.float32, n_trotters=64)
    r.set_replicas(20)
    r.prepare()
    assert r.get_sweep_mode() == 'field'
    # 28 trotters per CTA x 2048 doubles do not fit: automatic mode falls back to the classic kernel, forcing field mode is an error
    W = quantized_symmetric_W(2048, 6, np.float64)
    b = sq.dense_graph_annealer(W, sq.minimize, np.float64, n_trotters=4000)
    b.prepare()
    assert b.get_sweep_mode() == 'classic'
    c = sq.dense_graph_annealer(W, sq.minimize, np.float64, n_trotters=4000)
    c.set_sweep_mode('field')
    with pytest.raises(RuntimeError, match='shared memory'):
        c.prepare()


@pytest.mark.parametrize('N,m', [(1000, 296), (8192, 512)])
def test_field_writeback_drift_is_bounded(sq, N, m):
    """Field mode with carried fields on NON-dyadic J: after many steps the fields the sweep wrote back (seeded by the split-bf16
    tensor-core GEMM, then updated in fp32 with one J row per accepted flip) stay within a few fp32 roundings of a fresh float64
    evaluation h + 2 J q of the spins they belong to -- 50 steps without any recomputation (the production default refreshes far
    more often)."""
    rng = np.random.default_rng(N)
    A = rng.random((N, N), dtype=np.float32) - np.float32(0.5)
    W = np.ascontiguousarray(np.triu(A) + np.triu(A, 1).T)
    ann = sq.dense_graph_annealer(W, sq.minimize, np.float32, n_trotters=m)
    ann.set_sweep_mode('field', 1000)
    ann.seed(3); ann.prepare(); ann.randomize_spin()
    assert ann.get_fields() is None                      # nothing carried yet
    G, nsteps = 2.0, (50 if N < 4096 else 25)
    for _ in range(nsteps):
        ann.anneal_one_step(G, 50.0)
        G *= 0.85
    H = ann.get_fields()
    assert H is not None
    h, J, c = ann.get_hamiltonian()
    q = ann.get_spins().astype(np.float64)
    rows = np.arange(0, m, max(1, m // 16))
    fresh = h.astype(np.float64)[None, :] + 2.0 * (q[rows] @ J.astype(np.float64).T)
    err = np.abs(H[rows].astype(np.float64) - fresh).max()
    scale = np.abs(fresh).max()
    flips = ann.get_stats()['accepted'] / float(m)
    print('N=%d m=%d: %d steps, %.0f accepted flips per trotter, max |H - fresh| = %.3g (fields up to %.3g)' % (N, m, nsteps, flips, err, scale))
    assert flips > 100
    assert err <= 2e-4 * max(1.0, scale)


def test_problem_batch_then_single_problem(sq):
    """a problem batch followed by set_hamiltonian / set_qubo on the same solver anneals ONE problem again (the batch state is reset)"""
    N, m, R = 48, 6, 3
    Ws = np.stack([quantized_symmetric_W(N, 70 + r, np.float32) for r in range(R)])
    ann = sq.dense_graph_annealer(None, sq.minimize, np.float32)
    ann.set_qubo_batch(Ws)
    ann.set_preferences(n_trotters=m)
    ann.seed(1); ann.prepare(); ann.randomize_spin()
    ann.anneal_one_step(1.0, 10.0)
    assert ann.get_E().shape[0] == R * m
    one = sq.dense_graph_annealer(Ws[1], sq.minimize, np.float32, n_trotters=m)
    h, J, c = one.get_hamiltonian()
    for setter in ('hamiltonian', 'qubo'):
        if setter == 'hamiltonian':
            ann.set_hamiltonian(h, J, c)
        else:
            ann.set_qubo(Ws[1])
        ann.set_preferences(n_trotters=m)
        ann.seed(9); ann.prepare(); ann.randomize_spin()
        one.seed(9); one.prepare(); one.randomize_spin()
        for G in (2.0, 0.5):
            ann.anneal_one_step(G, 10.0); one.anneal_one_step(G, 10.0)
        assert ann.get_E().shape[0] == m
        assert np.array_equal(ann.get_spins(), one.get_spins())
        assert np.allclose(ann.get_E(), one.get_E(), rtol=1e-5, atol=1e-4)
