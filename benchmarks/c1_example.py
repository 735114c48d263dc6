#!/usr/bin/env python
"""BASELINE.json config C1: the reference's tutorial anneal (sqaodpy/example/dense_graph_annealer.py:22-70):
dense SQA, N=128, m=32, fp64, G: 5 -> 0.01 with G *= 0.99 (620 steps), beta = 1/0.02, seed 13255.
Runs the B200 solver and the reference CPU algorithm (oracle port, MT19937, all host cores) and prints one JSON line each."""
import json
import os
import sys
import time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def schedule():
    G, out = 5.0, []
    while G > 0.01:
        out.append(G)
        G *= 0.99
    return out


def main():
    N, m, beta = 128, 32, 1.0 / 0.02
    rng = np.random.default_rng(13255)
    A = rng.random((N, N)) - 0.5
    W = np.triu(A) + np.triu(A, 1).T
    Gs = schedule()
    import sqaod_b200 as sq
    ann = sq.dense_graph_annealer(W, sq.minimize, np.float64, n_trotters=m)
    ann.seed(13255); ann.prepare(); ann.randomize_spin()
    ann.anneal_one_step(Gs[0], beta); ann._device.synchronize()        # warm-up launch
    ann.seed(13255); ann.prepare(); ann.randomize_spin()
    t0 = time.perf_counter()
    for G in Gs:
        ann.anneal_one_step(G, beta)
    ann.make_solution()
    E = ann.get_E()
    dt = time.perf_counter() - t0
    print(json.dumps({'config': 'C1 dense SQA N=128 m=32 fp64, %d steps' % len(Gs), 'backend': 'sqaod_b200 (1 B200)', 'seconds': dt,
                      'attempts_per_s': len(Gs) * N * m / dt, 'E_min': float(E.min()), 'E_mean': float(E.mean())}), flush=True)
    from oracle import pyoracle as orc
    ref = orc.DenseGraphAnnealer(W, 0, np.float64, n_trotters=m, algorithm='coloring', n_workers=orc.num_threads(), rng='mt')
    ref.seed(13255); ref.prepare(); ref.randomize_spin()
    t0 = time.perf_counter()
    for G in Gs:
        ref.anneal_one_step(G, beta)
    E2 = ref.get_E()
    dt2 = time.perf_counter() - t0
    print(json.dumps({'config': 'C1 dense SQA N=128 m=32 fp64, %d steps' % len(Gs), 'backend': 'reference CPU algorithm (oracle port, %d cores)' % ref.n_workers,
                      'seconds': dt2, 'attempts_per_s': len(Gs) * N * m / dt2, 'E_min': float(E2.min()), 'E_mean': float(E2.mean())}), flush=True)


if __name__ == '__main__':
    main()
