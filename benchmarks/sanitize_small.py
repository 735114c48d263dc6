import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqaod_b200 as sq
rng = np.random.default_rng(3)
for N, m, dt in ((300, 296, np.float32), (72, 1800, np.float32), (2100, 12, np.float64), (40, 4001, np.float64), (130, 1, np.float32)):
    A = rng.random((N, N)) - 0.5
    W = (np.triu(A) + np.triu(A, 1).T).astype(dt)
    for mode, refresh in (('classic', 0), ('field', 0), ('field', 1000)):   # both sweep kernels, fields recomputed / carried
        ann = sq.dense_graph_annealer(W, sq.minimize, dt, n_trotters=m)
        ann.set_sweep_mode(mode, refresh)
        ann.seed(7); ann.prepare(); ann.randomize_spin()
        for _ in range(2):
            ann.anneal_one_step(1.0, 2.0)
        print(N, m, dt.__name__, mode, refresh, float(ann.get_E().min()))

# bipartite annealer (tcgen05 contraction for fp32, CUDA-core for fp64), brute-force searchers, formulas
for N0, N1, m, dt in ((70, 45, 9, np.float32), (33, 64, 6, np.float64)):
    b0 = (rng.random(N0) - 0.5).astype(dt); b1 = (rng.random(N1) - 0.5).astype(dt)
    W = (rng.random((N1, N0)) - 0.5).astype(dt)
    ann = sq.bipartite_graph_annealer(b0, b1, W, sq.minimize, dt, n_trotters=m)
    ann.seed(3); ann.prepare(); ann.randomize_spin()
    for _ in range(2):
        ann.anneal_one_step(1.0, 2.0)
    print('bipartite', N0, N1, m, dt.__name__, float(ann.get_E().min()))
A = rng.random((14, 14)) - 0.5
W = (np.triu(A) + np.triu(A, 1).T).astype(np.float32)
bf = sq.dense_graph_bf_searcher(W, sq.minimize, np.float32)
bf.search()
print('dense bf', float(bf.get_E()[0]), len(bf.get_x()))
bbf = sq.bipartite_graph_bf_searcher((rng.random(6) - 0.5), (rng.random(7) - 0.5), (rng.random((7, 6)) - 0.5), sq.minimize, np.float64)
bbf.search()
print('bipartite bf', float(bbf.get_E()[0]), len(bbf.get_x()))
