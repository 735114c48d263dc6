import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqaod_b200 as sq
rng = np.random.default_rng(3)
for N, m, dt in ((300, 296, np.float32), (72, 1800, np.float32), (2100, 12, np.float64), (40, 4001, np.float64), (130, 1, np.float32)):
    A = rng.random((N, N)) - 0.5
    W = (np.triu(A) + np.triu(A, 1).T).astype(dt)
    ann = sq.dense_graph_annealer(W, sq.minimize, dt, n_trotters=m)
    ann.seed(7); ann.prepare(); ann.randomize_spin()
    for _ in range(2):
        ann.anneal_one_step(1.0, 2.0)
    print(N, m, dt.__name__, float(ann.get_E().min()))
