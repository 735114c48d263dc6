"""Oracle checks of the BASELINE.json configurations AT THEIR OWN SIZES (through the C ABI):

  C2   dense SQA N=8192, m=512 fp32: one annealOneStep, spins equal to the CPU oracle's Philox-mode chain (restated
       sqaodc/cpu/CPUDenseGraphAnnealer.cpp:250-338), in the classic AND the field sweep; calculate_E vs the oracle.
  C3   bipartite SQA N0=N1=4096, m=512 fp32: one annealOneStep vs the oracle (CPUBipartiteGraphAnnealer.cpp:331-441).
  C4   dense brute force N=40 fp32: sampled x windows vs the oracle's range mode (CPUDenseGraphBatchSearch.cpp:25-50).
  C5b  dense SQA N=32768 (J = 4 GiB: every row offset needs 64-bit arithmetic), a few trotters, one step vs the oracle.
  C5a  N=1024, m=128 replicas: a replica batch equals separately seeded oracle chains.

Quantised W (2^-14 grid) keeps every sum exact in fp32, so the only way the trajectories can part is an accept test that
sits within a rounding error of its threshold; the oracle counts those, and a seed that hits one is replaced by the next."""
import numpy as np
import pytest
from conftest import quantized_symmetric_W, quantized_bipartite

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def sq():
    import sqaod_b200
    return sqaod_b200


def _dense_one_step_vs_oracle(sq, oracle, W, m, mode, seeds, G, beta, steps=1):
    workers = oracle.num_threads()
    for seed in seeds:
        ref = oracle.DenseGraphAnnealer(W, 0, np.float32, n_trotters=m, algorithm='coloring', n_workers=workers, rng='philox')
        ref.seed(seed); ref.prepare(); ref.randomize_spin()
        ann = sq.dense_graph_annealer(W, sq.minimize, np.float32, n_trotters=m)
        ann.set_sweep_mode(mode, 0)
        ann.seed(seed); ann.prepare(); ann.randomize_spin()
        assert ann.get_sweep_mode() == mode
        assert np.array_equal(ann.get_spins(), ref.get_q()), 'randomize_spin stream differs'
        same = True
        for s in range(steps):
            ref.anneal_one_step(G, beta); ann.anneal_one_step(G, beta)
            got, want = ann.get_spins(), ref.get_q()
            if not np.array_equal(got, want):
                assert ref.stats()[1] > 0, 'seed %d step %d: %d spins differ without a borderline accept test' % (seed, s, int((got != want).sum()))
                same = False
                break
        if same:
            assert ref.stats()[0] > 0 and ann.get_stats()['accepted'] == ref.stats()[0]
            E, Eref = ann.get_E().astype(np.float64), ref.get_E().astype(np.float64)
            assert np.allclose(E, Eref, rtol=1e-5, atol=1e-5 * np.abs(Eref).max())
            return
        del ann, ref
    pytest.fail('every seed hit a borderline accept test')


@pytest.mark.parametrize('mode', ['classic', 'field'])
def test_c2_one_step_equals_oracle(sq, oracle, mode):
    W = quantized_symmetric_W(8192, 8192, np.float32)
    _dense_one_step_vs_oracle(sq, oracle, W, 512, mode, (3, 4, 5, 6), 0.01, 50.0)


@pytest.mark.parametrize('mode', ['classic', 'field'])
def test_c5b_row_length_32768_equals_oracle(sq, oracle, mode):
    """N = 32768: J is 4 GiB, byte offsets of rows exceed 2^32.  Four trotters (one per CTA), one step."""
    N = 32768
    rng = np.random.default_rng(32768)
    W = rng.random((N, N), dtype=np.float32)
    W -= np.float32(0.5)
    W = np.triu(W)
    W += np.triu(W, 1).T
    np.rint(W * np.float32(16384), out=W)
    W /= np.float32(16384)
    _dense_one_step_vs_oracle(sq, oracle, W, 4, mode, (1, 2, 3), 0.5, 20.0)


def test_c5a_replica_batch_equals_oracle_chains(sq, oracle):
    """N=1024, m=128 (the C5a replica shape): replicas r of one batched launch == oracle chains seeded seed + r."""
    N, m, R, seed = 1024, 128, 6, 40
    W = quantized_symmetric_W(N, 1024, np.float32)
    batch = sq.dense_graph_annealer(W, sq.minimize, np.float32, n_trotters=m)
    batch.set_replicas(R)
    batch.seed(seed); batch.prepare(); batch.randomize_spin()
    Gs = (1.0, 0.2)
    for G in Gs:
        batch.anneal_one_step(G, 50.0)
    qb = batch.get_spins().reshape(R, m, N)
    checked = 0
    for r in range(R):
        ref = oracle.DenseGraphAnnealer(W, 0, np.float32, n_trotters=m, algorithm='coloring', n_workers=oracle.num_threads(), rng='philox')
        ref.seed(seed + r); ref.prepare(); ref.randomize_spin()
        for G in Gs:
            ref.anneal_one_step(G, 50.0)
        if np.array_equal(qb[r], ref.get_q()):
            checked += 1
        else:
            assert ref.stats()[1] > 0, 'replica %d differs without a borderline accept test' % r
    assert checked >= R - 2


def test_c3_one_step_equals_oracle(sq, oracle):
    N0 = N1 = 4096; m = 512
    b0, b1, W = quantized_bipartite(N0, N1, 4096, np.float32)
    for seed in (2, 3, 4, 5):
        ref = oracle.BipartiteGraphAnnealer(b0, b1, W, 0, np.float32, n_trotters=m, algorithm='coloring', n_workers=oracle.num_threads(), rng='philox')
        ref.seed(seed); ref.prepare(); ref.randomize_spin()
        ann = sq.bipartite_graph_annealer(b0, b1, W, sq.minimize, np.float32, n_trotters=m)
        ann.seed(seed); ann.prepare(); ann.randomize_spin()
        ref.anneal_one_step(0.01, 50.0); ann.anneal_one_step(0.01, 50.0)
        q = ann.get_q()
        got0, got1 = np.stack([p[0] for p in q]), np.stack([p[1] for p in q])
        want0, want1 = ref.get_q()
        nbad = int((got0 != want0).sum() + (got1 != want1).sum())
        if nbad == 0:
            assert ref.stats()[0] > 0
            E, Eref = ann.get_E().astype(np.float64), ref.get_E().astype(np.float64)
            assert np.allclose(E, Eref, rtol=1e-5, atol=1e-5 * np.abs(Eref).max())
            return
        assert ref.stats()[1] > 0, 'seed %d: %d spins differ without a borderline accept test' % (seed, nbad)
    pytest.fail('every seed hit a borderline accept test')


def test_c4_n40_windows_equal_oracle(sq, oracle):
    """N = 40: the full range is 2^40 states; windows at the start, across bit 39 and at the very end of the range are searched
    by both the B200 searcher (set_range) and the oracle's range mode: minimum and argmin lists bit-exact."""
    N = 40
    W = quantized_symmetric_W(N, 40, np.float32)
    span = 1 << 19
    windows = [(0, span), ((1 << 39) - span // 2, (1 << 39) + span // 2), ((1 << 40) - span, 1 << 40), (123456789012, 123456789012 + span + 77)]
    for b, e in windows:
        E0, xs0 = oracle.dense_graph_bf_search(W, 0, np.float32, tile_size=1 << 16, x_begin=b, x_end=e)
        s = sq.dense_graph_bf_searcher(W, sq.minimize, np.float32)
        s.set_range(b, e); s.prepare()
        while not s.search_range()[0]:
            pass
        assert s.get_Emin() == float(E0), (b, e)
        assert np.array_equal(np.sort(s.get_packed_x()), xs0), (b, e)
