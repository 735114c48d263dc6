import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqaod_b200 as sq
N, m = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(1)
A = rng.random((N, N), dtype=np.float32) - np.float32(0.5)
W = np.triu(A) + np.triu(A, 1).T
ann = sq.dense_graph_annealer(W, sq.minimize, np.float32, n_trotters=m)
ann.seed(1); ann.prepare(); ann.randomize_spin()
for _ in range(6):
    ann.anneal_one_step(0.01, 50.0)
ann._device.synchronize()
