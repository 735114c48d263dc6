#!/usr/bin/env python
"""The shapes of the reference's C++ performance program (sqaodc/tests/perf.cpp:87-219), same protocol: wall clock around
prepare + randomize_spin + 200 annealing steps (G geometric between 20 and 0.01, beta = 1/0.02) + make_solution, and around a
full brute-force search.

    dense brute force N = 24                       perf.cpp:118-138
    dense annealer N = 1024, m = 512               perf.cpp:141-165
    bipartite brute force (14, 14)                 perf.cpp:168-191
    bipartite annealer (1024, 512), m = 768        perf.cpp:194-219

(perf.cpp multiplies G by tau = (Ginit/Gfin)^(1/nSteps) > 1, i.e. its G grows from 20; the schedule here descends from 20 to
0.01 as the variable names say.)  `--cpu` times the reference's CPU algorithm (oracle port, all host cores) next to it."""
import argparse
import json
import os
import sys
import time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
N_STEPS, SEED = 200, 1133557


def anneal(an, get_E):
    Ginit, Gfin, beta = 20., 0.01, 1. / 0.02
    tau = (Gfin / Ginit) ** (1. / N_STEPS)
    t0 = time.perf_counter()
    an.prepare()
    an.randomize_spin()
    G = Ginit
    for _ in range(N_STEPS):
        an.anneal_one_step(G, beta)
        G *= tau
    E = get_E(an)
    return time.perf_counter() - t0, float(np.min(E))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cpu', action='store_true')
    ap.add_argument('--dtypes', default='float32,float64')
    args = ap.parse_args()
    import sqaod_b200 as sq
    for dname in args.dtypes.split(','):
        dtype = np.dtype(dname).type
        rng = np.random.default_rng(SEED)
        # dense brute force
        N = 24
        W = sq.generate_random_symmetric_W(N, dtype=dtype)
        s = sq.dense_graph_bf_searcher(W, sq.minimize, dtype)
        t0 = time.perf_counter(); s.search(); dt = time.perf_counter() - t0
        print(json.dumps({'shape': 'dense BF N=24', 'dtype': dname, 'seconds': dt, 'states_per_s': (1 << N) / dt, 'E_min': float(s.get_E()[0])}), flush=True)
        # dense annealer
        N, m = 1024, 512
        W = sq.generate_random_symmetric_W(N, dtype=dtype)
        an = sq.dense_graph_annealer(W, sq.minimize, dtype, n_trotters=m)
        an.seed(SEED)
        dt, E = anneal(an, lambda a: (a.make_solution(), a.get_E())[1])
        row = {'shape': 'dense annealer N=1024 m=512, 200 steps', 'dtype': dname, 'seconds': dt, 'attempts_per_s': N_STEPS * N * m / dt, 'E_min': E}
        if args.cpu:
            from oracle import pyoracle as orc
            ref = orc.DenseGraphAnnealer(W, 0, dtype, n_trotters=m, algorithm='coloring', n_workers=orc.num_threads(), rng='mt')
            ref.seed(SEED)
            dtc, Ec = anneal(ref, lambda a: a.get_E())
            row.update({'cpu_seconds': dtc, 'cpu_cores': orc.num_threads(), 'cpu_E_min': Ec})
        print(json.dumps(row), flush=True)
        # bipartite brute force
        N0 = N1 = 14
        b0, b1, Wb = np.asarray(rng.random(N0) - 0.5, dtype), np.asarray(rng.random(N1) - 0.5, dtype), np.asarray(rng.random((N1, N0)) - 0.5, dtype)
        s = sq.bipartite_graph_bf_searcher(b0, b1, Wb, sq.minimize, dtype)
        t0 = time.perf_counter(); s.search(); dt = time.perf_counter() - t0
        print(json.dumps({'shape': 'bipartite BF (14,14)', 'dtype': dname, 'seconds': dt, 'states_per_s': float(1 << (N0 + N1)) / dt, 'E_min': float(s.get_E()[0])}), flush=True)
        # bipartite annealer
        N0, N1 = 1024, 512
        m = (N0 + N1) // 2
        b0, b1, Wb = np.asarray(rng.random(N0) - 0.5, dtype), np.asarray(rng.random(N1) - 0.5, dtype), np.asarray(rng.random((N1, N0)) - 0.5, dtype)
        an = sq.bipartite_graph_annealer(b0, b1, Wb, sq.minimize, dtype, n_trotters=m)
        an.seed(SEED)
        dt, E = anneal(an, lambda a: (a.make_solution(), a.get_E())[1])
        row = {'shape': 'bipartite annealer (1024,512) m=768, 200 steps', 'dtype': dname, 'seconds': dt, 'attempts_per_s': N_STEPS * (N0 + N1) * m / dt, 'E_min': E}
        if args.cpu:
            from oracle import pyoracle as orc
            ref = orc.BipartiteGraphAnnealer(b0, b1, Wb, 0, dtype, n_trotters=m, algorithm='coloring', n_workers=orc.num_threads(), rng='mt')
            ref.seed(SEED)
            dtc, Ec = anneal(ref, lambda a: a.get_E())
            row.update({'cpu_seconds': dtc, 'cpu_cores': orc.num_threads(), 'cpu_E_min': Ec})
        print(json.dumps(row), flush=True)


if __name__ == '__main__':
    main()
