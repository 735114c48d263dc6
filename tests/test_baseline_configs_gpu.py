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


def _device_problem(sq, N, seed):
    """quantised symmetric W generated on the device (seconds instead of minutes at N = 32768); the host copy feeds the oracle"""
    gen = sq.dense_graph_annealer(None, sq.minimize, np.float32)
    return gen.get_qubo_random(N, seed, True)


def _dense_one_step_vs_oracle(sq, oracle, W, m, mode, seeds, G, beta, steps=1, random_spec=None):
    workers = oracle.num_threads()
    for seed in seeds:
        ref = oracle.DenseGraphAnnealer(W, 0, np.float32, n_trotters=m, algorithm='coloring', n_workers=workers, rng='philox')
        ref.seed(seed); ref.prepare(); ref.randomize_spin()
        if random_spec is None:
            ann = sq.dense_graph_annealer(W, sq.minimize, np.float32, n_trotters=m)
        else:   # the same matrix, generated in place on the device
            ann = sq.dense_graph_annealer(None, sq.minimize, np.float32)
            ann.set_qubo_random(random_spec[0], random_spec[1], True)
            ann.set_preferences(n_trotters=m)      # after the problem: setting a problem resets m to N / 4, as in the reference
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


@pytest.fixture(scope='module')
def w_c2(sq):
    return _device_problem(sq, 8192, 8192)


@pytest.fixture(scope='module')
def w_c5b(sq):
    return _device_problem(sq, 32768, 32768)


def test_device_generated_problem_is_what_it_says(sq, w_c2):
    assert np.array_equal(w_c2, w_c2.T) and np.abs(w_c2).max() <= 0.5
    assert np.array_equal(np.rint(w_c2 * 16384), w_c2 * 16384)                   # on the 2^-14 grid
    assert abs(float(w_c2.mean())) < 1e-3 and abs(float(w_c2.std()) - 12 ** -0.5) < 1e-3


@pytest.mark.parametrize('mode', ['classic', 'field'])
def test_c2_one_step_equals_oracle(sq, oracle, w_c2, mode):
    _dense_one_step_vs_oracle(sq, oracle, w_c2, 512, mode, (3, 4, 5, 6), 0.01, 50.0, random_spec=(8192, 8192))


@pytest.mark.parametrize('mode', ['classic', 'field'])
def test_c5b_row_length_32768_equals_oracle(sq, oracle, w_c5b, mode):
    """N = 32768: J is 4 GiB, byte offsets of rows exceed 2^32.  Four trotters (one per CTA), one step."""
    _dense_one_step_vs_oracle(sq, oracle, w_c5b, 4, mode, (1, 2, 3), 0.5, 20.0, random_spec=(32768, 32768))


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
