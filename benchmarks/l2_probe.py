"""One sweep launch per J size (run under `ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct -k regex:denseSweep`):
how much of the one-row-per-attempt traffic reaches HBM as J grows past the L2 capacity."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqaod_b200 as sq
m = 512
for N in [int(a) for a in sys.argv[1:]] or [4096, 5120, 5760, 6912, 8192]:
    rng = np.random.default_rng(1)
    A = rng.random((N, N), dtype=np.float32) - np.float32(0.5)
    W = np.triu(A) + np.triu(A, 1).T
    ann = sq.dense_graph_annealer(W, sq.minimize, np.float32, n_trotters=m)
    ann.seed(1); ann.prepare(); ann.randomize_spin()
    for _ in range(2):
        ann.anneal_one_step(0.01, 50.0)
    ann._device.synchronize()
    print('N=%d J=%.0f MiB algorithmic bytes per launch %.3e' % (N, N * N * 4 / 2**20, float(N) * N * m * 4))
    del ann
