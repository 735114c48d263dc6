"""Where the sweep kernel's time goes: busy cycles of dot warp 0, the chain warp, the helper warp (snapshots, masks) and the
prep warp (Philox tables), averaged per CTA, for a few problem sizes.  Profiling aid for DESIGN.md §3.1."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqaod_b200 as sq

CLK = 1.965e6   # cycles per ms at the boost clock the sweeps run at
rng = np.random.default_rng(1)
sizes = [tuple(int(v) for v in a.split('x')) for a in sys.argv[1:]] or \
        [(128, 32), (128, 128), (512, 512), (1024, 128), (1024, 1024), (2048, 256), (8192, 512)]
for N, m in sizes:
    A = rng.random((N, N), dtype=np.float32) - np.float32(0.5)
    W = np.triu(A) + np.triu(A, 1).T
    ann = sq.dense_graph_annealer(W, sq.minimize, np.float32, n_trotters=m)
    ann.seed(1); ann.prepare(); ann.randomize_spin()
    for _ in range(3):
        ann.anneal_one_step(0.01, 50.0)
    s0 = ann.get_stats()
    ann._device.synchronize(); t = time.perf_counter()
    n = 10
    for _ in range(n):
        ann.anneal_one_step(0.01, 50.0)
    ann._device.synchronize(); dt = (time.perf_counter() - t) / n * 1e3
    s1 = ann.get_stats(); G = min(148, m)
    d = {k: (s1[k] - s0[k]) / G / CLK / n for k in s1 if 'cycles' in k}
    print('N=%d m=%d: %.3f ms/step | per CTA busy ms: dot warp %.3f chain warp %.3f helper warp %.3f prep warp %.3f | chain waits: rows %.3f nb %.3f | '
          'flag polls/step %.0f accepted/step %.0f' %
          (N, m, dt, d['barrier_cycles_dot'], d['barrier_cycles_chain'], d['helper_cycles'], d['prep_cycles'], d['chain_wait_rows_cycles'], d['chain_wait_neighbour_cycles'],
           (s1['flag_waits'] - s0['flag_waits']) / n,
           (s1['accepted'] - s0['accepted']) / n))
