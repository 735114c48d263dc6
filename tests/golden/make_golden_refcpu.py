"""Generate tests/golden/refcpu_chains.npz: Markov chains of the REFERENCE'S OWN CPU annealers (sqaod.cpu, compiled from the reference
sources by `make -C oracle refcpu`), recorded spin for spin.

Run in the build container only (needs oracle/_ref, i.e. /root/reference at build time):

    python tests/golden/make_golden_refcpu.py

The process pins itself to ONE cpu before the libraries load, so the reference takes its serial forms with the single generator
MT19937(seed) (CPUDenseGraphAnnealer.cpp:281-300, CPUBipartiteGraphAnnealer.cpp:346-373) -- a chain that does not depend on the
machine.  The output is committed; tests/test_oracle_golden.py checks the oracle against it wherever the tests run (the compiled
reference is not needed for that).  Dense problems are given as (h, J, c) so that nothing but the annealing loop is exercised; the
bipartite contraction runs through the matrix library, hence quantised inputs (sums exact in any order)."""
import os
import sys

os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[0]})
os.environ['OMP_NUM_THREADS'] = '1'
import warnings  # noqa: E402
import numpy as np  # noqa: E402

warnings.simplefilter('ignore')
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import refsuite_runner  # noqa: E402

sq = refsuite_runner.assemble('cpu')
BETA = 1. / 0.02


def schedule(n, sa):
    return [2.0 * 0.5 ** k for k in range(n)] if sa else [3.0 * (0.02 / 3.0) ** (k / max(1.0, n - 1.0)) for k in range(n)]


def dense(out, key, N, m, dtype, algo, seed, steps):
    rng = np.random.default_rng(N * 131 + m)
    A = rng.random((N, N)) - 0.5
    W = np.triu(A) + np.triu(A, 1).T                    # non-dyadic on purpose
    J = (-0.25 * W).astype(dtype)
    np.fill_diagonal(J, 0)
    h = (-0.5 * W.sum(axis=0)).astype(dtype)
    c = dtype(0.25 * W.sum())
    ann = sq.cpu.dense_graph_annealer(dtype=dtype, algorithm=algo)
    ann.set_hamiltonian(h, J, c)
    ann.set_preferences(n_trotters=m)
    ann.seed(seed); ann.prepare(); ann.randomize_spin()
    Gs = schedule(steps, algo.startswith('sa'))
    traj = [np.asarray(ann.get_q(), np.int8)]
    for G in Gs:
        ann.anneal_one_step(G, BETA)
        traj.append(np.asarray(ann.get_q(), np.int8))
    out[key + '/h'], out[key + '/J'], out[key + '/c'] = h, J, np.asarray(c)
    out[key + '/G'], out[key + '/q'], out[key + '/E'] = np.asarray(Gs), np.asarray(traj), np.asarray(ann.get_E())
    out[key + '/meta'] = np.asarray([N, m, seed, 4 if dtype == np.float32 else 8])
    out[key + '/algo'] = np.asarray(algo)


def bipartite(out, key, N0, N1, m, dtype, algo, seed, steps):
    rng = np.random.default_rng(N0 * 17 + N1)
    W = (np.rint((rng.random((N1, N0)) - 0.5) * 64) / 64).astype(dtype)
    b0 = (np.rint((rng.random(N0) - 0.5) * 64) / 64).astype(dtype)
    b1 = (np.rint((rng.random(N1) - 0.5) * 64) / 64).astype(dtype)
    ann = sq.cpu.bipartite_graph_annealer(b0, b1, W, sq.minimize, dtype, n_trotters=m, algorithm=algo)
    ann.seed(seed); ann.prepare(); ann.randomize_spin()
    Gs = schedule(steps, algo.startswith('sa'))

    def snap():
        q = ann.get_q()
        return np.asarray([p[0] for p in q], np.int8), np.asarray([p[1] for p in q], np.int8)
    t0, t1 = [], []
    a, b = snap(); t0.append(a); t1.append(b)
    for G in Gs:
        ann.anneal_one_step(G, BETA)
        a, b = snap(); t0.append(a); t1.append(b)
    out[key + '/b0'], out[key + '/b1'], out[key + '/W'] = b0, b1, W
    out[key + '/G'], out[key + '/q0'], out[key + '/q1'], out[key + '/E'] = np.asarray(Gs), np.asarray(t0), np.asarray(t1), np.asarray(ann.get_E())
    out[key + '/meta'] = np.asarray([N0, N1, m, seed, 4 if dtype == np.float32 else 8])
    out[key + '/algo'] = np.asarray(algo)


def config_c1(out):
    """BASELINE.json configs[0] (C1), the reference's tutorial anneal through sqaod.cpu (sqaodpy/example/dense_graph_annealer.py:22-70):
    N = 128, m = 32, fp64, seed 13255, G = 5 -> 0.01 with G *= 0.99 (619 steps), beta = 1/0.02; W ~ U(-0.5, 0.5) symmetric, not
    quantised.  Spins recorded every 124 steps and at the end; QUBO -> Ising goes through set_hamiltonian as above."""
    N, m, beta, dtype = 128, 32, 1. / 0.02, np.float64
    rng = np.random.default_rng(13255)
    A = rng.random((N, N)) - 0.5
    W = np.triu(A) + np.triu(A, 1).T
    J = (-0.25 * W).astype(dtype)
    np.fill_diagonal(J, 0)
    h = (-0.5 * W.sum(axis=0)).astype(dtype)
    c = dtype(0.25 * W.sum())
    ann = sq.cpu.dense_graph_annealer(dtype=dtype, algorithm=sq.algorithm.coloring)
    ann.set_hamiltonian(h, J, c)
    ann.set_preferences(n_trotters=m)
    ann.seed(13255); ann.prepare(); ann.randomize_spin()
    G, k, snaps, at = 5.0, 0, [], []
    while 0.01 <= G:
        ann.anneal_one_step(G, beta)
        G *= 0.99
        k += 1
        if k % 124 == 0:
            snaps.append(np.asarray(ann.get_q(), np.int8)); at.append(k)
    snaps.append(np.asarray(ann.get_q(), np.int8)); at.append(k)
    out['c1/h'], out['c1/J'], out['c1/c'] = h, J, np.asarray(c)
    out['c1/q'], out['c1/at'], out['c1/E'] = np.asarray(snaps), np.asarray(at), np.asarray(ann.get_E())
    out['c1/steps'] = np.asarray(k)


def main():
    A = sq.algorithm
    out = {}
    config_c1(out)
    k = 0
    for dtype in (np.float32, np.float64):
        for (N, m, algo, steps) in ((40, 10, A.coloring, 4), (33, 7, A.coloring, 4), (130, 4, A.coloring, 3), (24, 6, A.naive, 3),
                                    (31, 5, A.sa_naive, 3), (48, 1, A.sa_naive, 3)):
            dense(out, 'dense%d' % k, N, m, dtype, algo, 100 + k, steps); k += 1
    k = 0
    for dtype in (np.float32, np.float64):
        for (N0, N1, m, algo, steps) in ((12, 9, 8, A.coloring, 3), (20, 33, 5, A.coloring, 3), (10, 7, 4, A.naive, 2),
                                         (16, 12, 4, A.sa_coloring, 3), (8, 6, 3, A.sa_naive, 2)):
            bipartite(out, 'bip%d' % k, N0, N1, m, dtype, algo, 200 + k, steps); k += 1
    np.savez_compressed(os.path.join(HERE, 'refcpu_chains.npz'), **out)
    print('wrote refcpu_chains.npz: %d arrays' % len(out))


if __name__ == '__main__':
    main()
