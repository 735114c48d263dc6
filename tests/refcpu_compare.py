#!/usr/bin/env python
"""Pin the CPU oracle (oracle/oracle.cpp) against the REFERENCE'S OWN CPU back end run here.

`make -C oracle refcpu` compiles sqaodc/common + sqaodc/cpu + the reference's cpu_*.cpp CPython glue, every source unmodified and
where it lies, into oracle/_ref/ (Eigen, which this image lacks, is replaced by oracle/eigen_standin -- the Metropolis loops of the
reference do not go through it).  This script drives both through their Python front ends on the same inputs and seeds:

    python tests/refcpu_compare.py WORKERS     prints one line per case, `ok` or `DIFF`, and REFCPU_COMPARE_OK at the end

WORKERS is the number of CPUs the process pins itself to BEFORE the libraries load: the reference sizes its worker pool and its
per-worker MT19937 streams (seed + 17 i) from the affinity mask (common/os_dependent_linux.cpp:5-10), and with more than one worker it
takes the OpenMP form of the colouring sweep (CPUDenseGraphAnnealer.cpp:303-329, CPUBipartiteGraphAnnealer.cpp:377-430).

Run as a subprocess by tests/test_oracle_vs_reference_cpu.py; needs no GPU and nothing of the product library."""
import os
import sys

WORKERS = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cpus = sorted(os.sched_getaffinity(0))
if len(cpus) < WORKERS:
    print('REFCPU_COMPARE_SKIP only %d cpus' % len(cpus))
    sys.exit(0)
os.sched_setaffinity(0, set(cpus[:WORKERS]))
os.environ['OMP_NUM_THREADS'] = str(WORKERS)
os.environ['OMP_DYNAMIC'] = 'false'

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import refsuite_runner  # noqa: E402

if not os.path.exists(os.path.join(refsuite_runner.SUITE, 'glue_cpu', 'cpu_dg_annealer.so')):
    print('REFCPU_COMPARE_SKIP reference CPU build absent (make -C oracle refcpu)')
    sys.exit(0)
sq = refsuite_runner.assemble('cpu')
from oracle import pyoracle as orc  # noqa: E402

FAILED = []


def report(name, ok, detail=''):
    print('%-100s %s %s' % (name, 'ok' if ok else 'DIFF', detail), flush=True)
    if not ok:
        FAILED.append(name)


def sym(rng, N, grid):
    """symmetric W in [-0.5, 0.5); grid = 0: as drawn (non-dyadic), else rounded to multiples of 1/grid (every sum exact)"""
    A = rng.random((N, N)) - 0.5
    W = np.triu(A) + np.triu(A, 1).T
    return np.rint(W * grid) / grid if grid else W


def schedule(n):
    return [3.0 * (0.02 / 3.0) ** (k / max(1.0, n - 1.0)) for k in range(n)]


# ------------------------------------------------------------------ dense annealer
def dense_case(N, m, dtype, algo, seed, via, steps=4):
    rng = np.random.default_rng(1000 * N + m)
    name = 'dense annealer N=%d m=%d %s %s workers=%d via %s' % (N, m, np.dtype(dtype).name, algo, WORKERS, via)
    ref = sq.cpu.dense_graph_annealer(dtype=dtype, n_trotters=m, algorithm=algo)
    mine = orc.DenseGraphAnnealer(None, 0, dtype, n_trotters=m, algorithm=algo, n_workers=WORKERS, rng='mt')
    if via == 'qubo':           # QUBO -> Ising runs through the matrix library: quantised W, every sum exact
        W = sym(rng, N, 64).astype(dtype)
        ref.set_qubo(W, sq.minimize)
        mine.set_qubo(W, 0)
        h, J, c = ref.get_hamiltonian()
        h2, J2, c2 = mine.get_hamiltonian()
        if not (np.array_equal(h, h2) and np.array_equal(J, J2) and c == c2):
            return report(name, False, 'hamiltonian differs')
    else:                       # the Metropolis loop itself on a non-dyadic problem: hand both the same h, J, c
        W = sym(rng, N, 0)
        J = (-0.25 * W).astype(dtype)
        np.fill_diagonal(J, 0)
        h = (-0.5 * W.sum(axis=0)).astype(dtype)
        c = dtype(0.25 * W.sum())
        ref.set_hamiltonian(h, J, c)
        mine.set_hamiltonian(h, J, c)
    ref.set_preferences(n_trotters=m)     # setting a problem resets the trotter count to N / 4 (CPUDenseGraphAnnealer.cpp:53-83)
    ref.seed(seed); mine.seed(seed)
    ref.prepare(); mine.prepare()
    ref.randomize_spin(); mine.randomize_spin()
    if not np.array_equal(np.asarray(ref.get_q()), mine.get_q()):
        return report(name, False, 'randomize_spin differs')
    beta = 1. / 0.02
    Gs = schedule(steps) if not algo.startswith('sa') else [2.0, 1.0, 0.5, 0.25][:steps]
    for k, G in enumerate(Gs):
        ref.anneal_one_step(G, beta); mine.anneal_one_step(G, beta)
        if not np.array_equal(np.asarray(ref.get_q()), mine.get_q()):
            return report(name, False, 'spins differ after step %d' % k)
    tol = 2e-5 if dtype == np.float32 else 1e-12
    Eok = np.allclose(np.asarray(ref.get_E()), mine.get_E(), rtol=tol, atol=tol * N)
    sok = np.isclose(ref.get_system_E(Gs[-1], beta), mine.get_system_E(Gs[-1], beta), rtol=tol, atol=tol * N)
    report(name, Eok and sok, '' if (Eok and sok) else 'energies differ')


# ------------------------------------------------------------------ bipartite annealer (its contraction runs through the matrix library: quantised inputs)
def bipartite_case(N0, N1, m, dtype, algo, seed, steps=3):
    rng = np.random.default_rng(77 * N0 + N1 + m)
    name = 'bipartite annealer N0=%d N1=%d m=%d %s %s workers=%d' % (N0, N1, m, np.dtype(dtype).name, algo, WORKERS)
    W = (np.rint((rng.random((N1, N0)) - 0.5) * 64) / 64).astype(dtype)
    b0 = (np.rint((rng.random(N0) - 0.5) * 64) / 64).astype(dtype)
    b1 = (np.rint((rng.random(N1) - 0.5) * 64) / 64).astype(dtype)
    ref = sq.cpu.bipartite_graph_annealer(b0, b1, W, sq.minimize, dtype, n_trotters=m, algorithm=algo)
    mine = orc.BipartiteGraphAnnealer(b0, b1, W, 0, dtype, n_trotters=m, algorithm=algo, n_workers=WORKERS, rng='mt')
    ref.seed(seed); mine.seed(seed)
    ref.prepare(); mine.prepare()
    ref.randomize_spin(); mine.randomize_spin()

    def same():
        rq = ref.get_q()
        q0, q1 = mine.get_q()
        r0 = np.asarray([p[0] for p in rq]); r1 = np.asarray([p[1] for p in rq])
        return np.array_equal(r0, q0) and np.array_equal(r1, q1)
    if not same():
        return report(name, False, 'randomize_spin differs')
    beta = 1. / 0.02
    Gs = schedule(steps) if not algo.startswith('sa') else [2.0, 1.0, 0.5][:steps]
    for k, G in enumerate(Gs):
        ref.anneal_one_step(G, beta); mine.anneal_one_step(G, beta)
        if not same():
            return report(name, False, 'spins differ after step %d' % k)
    ok = np.array_equal(np.asarray(ref.get_E()), mine.get_E())
    report(name, ok, '' if ok else 'energies differ')


# ------------------------------------------------------------------ brute force
def dense_bf_case(N, dtype, optimize, tile):
    rng = np.random.default_rng(5 * N)
    name = 'dense brute force N=%d %s %s tile=%d' % (N, np.dtype(dtype).name, 'max' if optimize else 'min', tile)
    W = sym(rng, N, 16).astype(dtype)
    s = sq.cpu.dense_graph_bf_searcher(W, sq.maximize if optimize else sq.minimize, dtype, tile_size=tile)
    s.search()
    E, xs = orc.dense_graph_bf_search(W, optimize, dtype, tile_size=tile)
    rx = np.asarray(s.get_x())
    mine = np.asarray([orc.unpack_bits(x, N) for x in xs])
    ok = np.all(np.asarray(s.get_E()) == E) and rx.shape == mine.shape and \
        np.array_equal(np.asarray(sorted(map(tuple, rx))), np.asarray(sorted(map(tuple, mine))))
    report(name, bool(ok))


def degenerate_bf_case(dtype):
    """W = 0: every state is a minimum; the list keeps the lowest tile_size states (CPUDenseGraphBatchSearch.cpp:38-40,
    CPUDenseGraphBFSearcher.cpp:103-131; the solver rounds tile sizes up to multiples of 256, common/Solver.cpp:201-209)."""
    N, tile = 10, 256
    W = np.zeros((N, N), dtype)
    s = sq.cpu.dense_graph_bf_searcher(W, sq.minimize, dtype, tile_size=tile)
    s.search()
    E, xs = orc.dense_graph_bf_search(W, 0, dtype, tile_size=tile)
    rx = np.asarray(s.get_x())
    mine = np.asarray([orc.unpack_bits(x, N) for x in xs])
    report('dense brute force W=0 N=%d %s: %d of %d states kept' % (N, np.dtype(dtype).name, len(rx), 1 << N),
           len(rx) == tile and np.array_equal(rx, mine) and float(E) == float(np.asarray(s.get_E())[0]))


def bipartite_bf_case(N0, N1, dtype, optimize):
    rng = np.random.default_rng(N0 * 31 + N1)
    name = 'bipartite brute force N0=%d N1=%d %s %s' % (N0, N1, np.dtype(dtype).name, 'max' if optimize else 'min')
    W = (np.rint((rng.random((N1, N0)) - 0.5) * 16) / 16).astype(dtype)
    b0 = (np.rint((rng.random(N0) - 0.5) * 16) / 16).astype(dtype)
    b1 = (np.rint((rng.random(N1) - 0.5) * 16) / 16).astype(dtype)
    s = sq.cpu.bipartite_graph_bf_searcher(b0, b1, W, sq.maximize if optimize else sq.minimize, dtype)
    s.search()
    E, pairs = orc.bipartite_graph_bf_search(b0, b1, W, optimize, dtype)
    ref_pairs = sorted((tuple(p[0]), tuple(p[1])) for p in s.get_x())
    mine = sorted((tuple(orc.unpack_bits(a, N0)), tuple(orc.unpack_bits(b, N1))) for a, b in pairs)
    report(name, bool(np.all(np.asarray(s.get_E()) == E)) and ref_pairs == mine)


# ------------------------------------------------------------------ formulas (through the matrix library: exact on quantised inputs, to rounding otherwise)
def formulas_case(dtype, grid):
    rng = np.random.default_rng(3)
    N, B = 37, 11
    W = sym(rng, N, grid).astype(dtype)
    x = rng.integers(0, 2, (B, N)).astype(np.int8)
    q = (2 * x - 1).astype(np.int8)
    tol = 0 if grid else (3e-5 if dtype == np.float32 else 1e-12)
    eq = (lambda a, b: np.array_equal(np.asarray(a), np.asarray(b))) if grid else (lambda a, b: np.allclose(a, b, rtol=tol, atol=tol))
    F = sq.cpu.formulas
    name = 'formulas %s %s' % (np.dtype(dtype).name, 'quantised' if grid else 'non-dyadic')
    ok = eq(F.dense_graph_batch_calculate_E(W, x, dtype), orc.dense_graph_batch_calculate_E(W, x, dtype))
    h, J, c = F.dense_graph_calculate_hamiltonian(W, dtype)
    h2, J2, c2 = orc.dense_graph_calculate_hamiltonian(W, dtype)
    ok = ok and eq(h, h2) and eq(J, J2) and eq(c, c2)
    ok = ok and eq(F.dense_graph_batch_calculate_E_from_spin(h, J, c, q, dtype), orc.dense_graph_batch_calculate_E_from_spin(h2, J2, c2, q, dtype))
    N0, N1 = 9, 14
    Wb = (rng.random((N1, N0)) - 0.5)
    b0 = rng.random(N0) - 0.5
    b1 = rng.random(N1) - 0.5
    if grid:
        Wb, b0, b1 = np.rint(Wb * grid) / grid, np.rint(b0 * grid) / grid, np.rint(b1 * grid) / grid
    Wb, b0, b1 = Wb.astype(dtype), b0.astype(dtype), b1.astype(dtype)
    x0 = rng.integers(0, 2, (B, N0)).astype(np.int8)
    x1 = rng.integers(0, 2, (B, N1)).astype(np.int8)
    ok = ok and eq(F.bipartite_graph_batch_calculate_E(b0, b1, Wb, x0, x1, dtype), orc.bipartite_graph_batch_calculate_E(b0, b1, Wb, x0, x1, dtype))
    # the reference's Python wrapper of the 2-D form refers to undefined names (common/formulas_base.py:115): call its C extension
    E2d = np.empty((5, B), dtype)
    F.cext.bipartite_graph_batch_calculate_E_2d(F._bgobj, E2d, b0, b1, Wb, x0, np.ascontiguousarray(x1[:5]), dtype)
    ok = ok and eq(E2d, orc.bipartite_graph_batch_calculate_E_2d(b0, b1, Wb, x0, x1[:5], dtype))
    h0, h1, Jb, cb = F.bipartite_graph_calculate_hamiltonian(b0, b1, Wb, dtype)
    g0, g1, Jb2, cb2 = orc.bipartite_graph_calculate_hamiltonian(b0, b1, Wb, dtype)
    ok = ok and eq(h0, g0) and eq(h1, g1) and eq(Jb, Jb2) and eq(cb, cb2)
    ok = ok and eq(F.bipartite_graph_batch_calculate_E_from_spin(h0, h1, Jb, cb, 2 * x0 - 1, 2 * x1 - 1, dtype),
                   orc.bipartite_graph_batch_calculate_E_from_spin(g0, g1, Jb2, cb2, (2 * x0 - 1).astype(np.int8), (2 * x1 - 1).astype(np.int8), dtype))
    report(name, bool(ok))


# ------------------------------------------------------------------ the north star's statistical criterion, on the CPU
def statistics_case(N, m, steps, algo, nseeds=256):
    """The B200 sweep walks the oracle's Philox-mode chain bit for bit (GPU exact-chain tests), and the RNG streams of the two solvers
    differ by design -- so annealer output is judged statistically (BASELINE.json): over 256 seeds the final-energy distribution and
    the ground-state hit rate of the Philox chain must be indistinguishable from those of the reference's own sqaod.cpu annealer.
    Same thresholds as tests/test_annealer_statistics_gpu.py."""
    from scipy import stats
    rng = np.random.default_rng(2024 + N)
    W = sym(rng, N, 16384).astype(np.float32)
    beta = 1. / 0.02
    sa = algo.startswith('sa')
    G0, G1 = (2.0, 0.02) if sa else (5.0, 0.01)
    Gs = [G0 * (G1 / G0) ** (k / float(steps)) for k in range(steps)]
    e_ph, e_ref = np.empty(nseeds), np.empty(nseeds)
    ref = sq.cpu.dense_graph_annealer(W, sq.minimize, np.float32, n_trotters=m, algorithm=algo)
    for s in range(nseeds):
        ref.seed(s); ref.prepare(); ref.randomize_spin()
        mine = orc.DenseGraphAnnealer(W, 0, np.float32, n_trotters=m, algorithm=algo, n_workers=1, rng='philox')
        mine.seed(s); mine.prepare(); mine.randomize_spin()
        for G in Gs:
            ref.anneal_one_step(G, beta); mine.anneal_one_step(G, beta)
        e_ref[s] = np.min(ref.get_E()); e_ph[s] = mine.get_E().min()
    ground = min(e_ph.min(), e_ref.min())
    tol = 2e-5 * max(1.0, abs(ground))
    hit_p, hit_r = float((e_ph <= ground + tol).mean()), float((e_ref <= ground + tol).mean())
    p = 0.5 * (hit_p + hit_r)
    sigma = max(np.sqrt(2 * p * (1 - p) / nseeds), 1e-3)
    ks = stats.ks_2samp(np.round(e_ph / tol) * tol, np.round(e_ref / tol) * tol).pvalue
    se = np.sqrt(e_ph.var() / nseeds + e_ref.var() / nseeds) + 1e-9
    ok = abs(hit_p - hit_r) < 4 * sigma and ks > 1e-3 and abs(e_ph.mean() - e_ref.mean()) < 4 * se + tol
    report('statistics over %d seeds, dense %s N=%d m=%d %d steps: Philox chain vs sqaod.cpu' % (nseeds, algo, N, m, steps), bool(ok),
           'hit rate %.3f vs %.3f, KS p = %.3f, mean %.4f vs %.4f' % (hit_p, hit_r, ks, e_ph.mean(), e_ref.mean()))


def main():
    A = sq.algorithm
    for dtype in (np.float32, np.float64):
        # the colouring sweep: serial with one worker, the OpenMP form with more; even and odd rings, row lengths that leave a SIMD tail
        for (N, m) in ((40, 10), (33, 7), (130, 12), (257, 5), (64, 2)):
            for via in ('hamiltonian', 'qubo'):
                dense_case(N, m, dtype, A.coloring, 11 + N, via)
        if WORKERS == 1:
            # single RNG stream forms: naive SQA, SA (the reference shares ONE generator between the OpenMP threads of its SA loop,
            # CPUDenseGraphAnnealer.cpp:358-369, so only the one-worker run is a deterministic chain), m = 1 (SA by default)
            dense_case(24, 6, dtype, A.naive, 5, 'hamiltonian')
            dense_case(31, 9, dtype, A.sa_naive, 6, 'hamiltonian')
            dense_case(48, 1, dtype, A.sa_naive, 7, 'hamiltonian')
        for (N0, N1, m) in ((12, 9, 8), (20, 33, 5), (64, 48, 4)):
            bipartite_case(N0, N1, m, dtype, A.coloring, 21 + N0)
        if WORKERS == 1:
            bipartite_case(10, 7, 4, dtype, A.naive, 8)
            bipartite_case(16, 12, 4, dtype, A.sa_coloring, 9)
            for N in (4, 9, 12):
                for tile in (256, 1024):
                    dense_bf_case(N, dtype, 0, tile)
            dense_bf_case(10, dtype, 1, 256)
            degenerate_bf_case(dtype)
            bipartite_bf_case(5, 6, dtype, 0)
            bipartite_bf_case(7, 4, dtype, 1)
            formulas_case(dtype, 64)
            formulas_case(dtype, 0)
    if WORKERS == 1:
        statistics_case(24, 4, 4, A.coloring)
        statistics_case(64, 16, 20, A.coloring)
        statistics_case(48, 8, 12, A.sa_naive)
        statistics_case(128, 32, 20, A.coloring)      # the shape of BASELINE config C1 (N = 128, m = 32), fp32, shortened schedule
    if FAILED:
        print('REFCPU_COMPARE_FAILED %d: %s' % (len(FAILED), '; '.join(FAILED)))
        return 1
    print('REFCPU_COMPARE_OK workers=%d' % WORKERS)
    return 0


if __name__ == '__main__':
    sys.exit(main())
