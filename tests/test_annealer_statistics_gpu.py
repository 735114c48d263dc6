"""Statistical parity of the annealers with the reference CPU chain (BASELINE.json north_star): the RNG streams differ
(Philox on the GPU, per-thread MT19937 in sqaod.cpu), so over 256 seeds the final-energy distribution and the
ground-state hit rate of the B200 solver must be indistinguishable from the reference CPU algorithm's (restated in
oracle/, MT19937 mode, single worker = annealOneStepColoring, CPUDenseGraphAnnealer.cpp:281-300)."""
import numpy as np
import pytest
from scipy import stats
from conftest import quantized_symmetric_W, quantized_bipartite

pytestmark = pytest.mark.gpu
NSEEDS = 256


def schedule(steps, Ginit=5.0, Gfin=0.01):
    tau = (Gfin / Ginit) ** (1.0 / steps)
    return [Ginit * tau ** k for k in range(steps)]


def compare(e_gpu, e_ref, ground):
    hit_g, hit_r = float((e_gpu == ground).mean()), float((e_ref == ground).mean())
    p = 0.5 * (hit_g + hit_r)
    sigma = max(np.sqrt(2 * p * (1 - p) / NSEEDS), 1e-3)
    assert abs(hit_g - hit_r) < 4 * sigma, 'ground-state hit rate %.3f (B200) vs %.3f (reference CPU)' % (hit_g, hit_r)
    assert stats.ks_2samp(e_gpu, e_ref).pvalue > 1e-3
    se = np.sqrt(e_gpu.var() / NSEEDS + e_ref.var() / NSEEDS) + 1e-9
    assert abs(e_gpu.mean() - e_ref.mean()) < 4 * se
    return hit_g, hit_r


@pytest.mark.parametrize('N,m,steps,algo', [(24, 4, 4, 'coloring'), (64, 16, 20, 'coloring'), (48, 8, 12, 'sa_naive')])
def test_dense_final_energy_distribution(oracle, N, m, steps, algo):
    import sqaod_b200 as sq
    W = quantized_symmetric_W(N, 2024, np.float32)
    beta = 1. / 0.02
    Gs = schedule(steps) if algo == 'coloring' else schedule(steps, 2.0, 0.02)
    ann = sq.dense_graph_annealer(W, sq.minimize, np.float32, n_trotters=m, algorithm=algo)
    e_gpu, e_ref = np.empty(NSEEDS), np.empty(NSEEDS)
    for s in range(NSEEDS):
        ann.seed(s); ann.prepare(); ann.randomize_spin()
        for G in Gs:
            ann.anneal_one_step(G, beta)
        e_gpu[s] = ann.get_E().min()
        ref = oracle.DenseGraphAnnealer(W, 0, np.float32, n_trotters=m, algorithm=algo, n_workers=1, rng='mt')
        ref.seed(s); ref.prepare(); ref.randomize_spin()
        for G in Gs:
            ref.anneal_one_step(G, beta)
        e_ref[s] = ref.get_E().min()
    if N <= 24:
        bf = sq.dense_graph_bf_searcher(W, sq.minimize, np.float32)
        bf.search()
        ground = float(bf.get_E()[0])
        assert min(e_gpu.min(), e_ref.min()) >= ground - 1e-4
    else:
        ground = min(e_gpu.min(), e_ref.min())
    hit_g, hit_r = compare(e_gpu, e_ref, ground)
    print('N=%d m=%d %s: hit rate B200 %.3f, reference CPU %.3f' % (N, m, algo, hit_g, hit_r))


@pytest.mark.parametrize('algo', ['coloring', 'sa_coloring'])
def test_bipartite_final_energy_distribution(oracle, algo):
    import sqaod_b200 as sq
    N0, N1, m, steps = 20, 16, 8, 6
    b0, b1, W = quantized_bipartite(N0, N1, 99, np.float32)
    beta = 1. / 0.02
    Gs = schedule(steps) if algo == 'coloring' else schedule(steps, 2.0, 0.02)
    ann = sq.bipartite_graph_annealer(b0, b1, W, sq.minimize, np.float32, n_trotters=m, algorithm=algo)
    e_gpu, e_ref = np.empty(NSEEDS), np.empty(NSEEDS)
    for s in range(NSEEDS):
        ann.seed(s); ann.prepare(); ann.randomize_spin()
        for G in Gs:
            ann.anneal_one_step(G, beta)
        e_gpu[s] = ann.get_E().min()
        ref = oracle.BipartiteGraphAnnealer(b0, b1, W, 0, np.float32, n_trotters=m, algorithm=algo, n_workers=1, rng='mt')
        ref.seed(s); ref.prepare(); ref.randomize_spin()
        for G in Gs:
            ref.anneal_one_step(G, beta)
        e_ref[s] = ref.get_E().min()
    ground = min(e_gpu.min(), e_ref.min())
    compare(e_gpu, e_ref, ground)
