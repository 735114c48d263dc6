#!/usr/bin/env python
"""Secondary measurements of the other SURVEY section 8 rows on one B200 (bench.py is the headline):
  energy   : calculate_E at C2 (N=8192, m=512, fp32), tcgen05 split GEMM vs CUDA-core kernel
  bipartite: C3 anneal_one_step (N0=N1=4096, m=512, fp32), tcgen05 vs CUDA-core contraction
  bf       : dense brute force, states/s for N = 28..36 (fp32 solver, double arithmetic inside)
Each line is JSON; times are CUDA events on the launching stream after warm-up."""
import argparse
import json
import os
import sys
import time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--what', default='energy,bipartite,bf')
    ap.add_argument('--bf-n', default='28,32,36')
    args = ap.parse_args()
    import torch
    import sqaod_b200 as sq
    dev = sq.Device(0)
    sq.set_active_device(dev)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dev.set_stream(stream.cuda_stream)

    def timed(fn, reps, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    rng = np.random.default_rng(1133557)
    what = args.what.split(',')
    if 'energy' in what:
        N, m = 8192, 512
        A = rng.random((N, N), dtype=np.float32) - np.float32(0.5)
        W = np.triu(A) + np.triu(A, 1).T
        for tc in (True, False):
            os.environ['SQAOD_B200_NO_TC'] = '0' if tc else '1'
            ann = sq.dense_graph_annealer(W, sq.minimize, np.float32, n_trotters=m)
            ann.seed(1); ann.prepare(); ann.randomize_spin()
            ms = timed(ann.calculate_E, 10)
            flops = 2.0 * m * N * N + 4.0 * m * N
            print(json.dumps({'row': 'a7 calculate_E', 'config': 'C2 N=8192 m=512 fp32', 'path': 'tcgen05 bf16x3' if tc else 'cuda-core fp32',
                              'ms': ms, 'algorithmic_TFLOPs': flops / ms / 1e9, 'E0': float(ann.get_E()[0])}), flush=True)
            del ann
    if 'bipartite' in what:
        N0 = N1 = 4096; m = 512
        b0 = rng.random(N0, dtype=np.float32) - np.float32(0.5)
        b1 = rng.random(N1, dtype=np.float32) - np.float32(0.5)
        W = rng.random((N1, N0), dtype=np.float32) - np.float32(0.5)
        for tc in (True, False):
            os.environ['SQAOD_B200_NO_TC'] = '0' if tc else '1'
            ann = sq.bipartite_graph_annealer(b0, b1, W, sq.minimize, np.float32, n_trotters=m)
            ann.seed(1); ann.prepare(); ann.randomize_spin()
            ms = timed(lambda: ann.anneal_one_step(0.01, 50.0), 20, warm=3)
            attempts = (N0 + N1) * m
            print(json.dumps({'row': 'a6 bipartite annealOneStep', 'config': 'C3 N0=N1=4096 m=512 fp32', 'path': 'tcgen05 bf16x3' if tc else 'cuda-core fp32',
                              'ms_per_step': ms, 'attempts_per_s': attempts / ms * 1e3, 'algorithmic_TFLOPs': 4.0 * m * N0 * N1 / ms / 1e9,
                              'E_min': float(ann.get_E().min())}), flush=True)
            del ann
        os.environ['SQAOD_B200_NO_TC'] = '0'
    if 'bf' in what:
        for N in [int(v) for v in args.bf_n.split(',')]:
            A = np.rint((rng.random((N, N)) - 0.5) * 16384) / 16384.
            W = np.asarray(np.triu(A) + np.triu(A, 1).T, np.float32)
            s = sq.dense_graph_bf_searcher(W, sq.minimize, np.float32)
            t0 = time.perf_counter()
            s.search()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(json.dumps({'row': 'a11 dense brute force', 'config': 'N=%d fp32 quantised W' % N, 'seconds': dt, 'states_per_s': (2.0 ** N) / dt,
                              'E_min': float(s.get_E()[0]), 'n_solutions': len(s.get_x())}), flush=True)


if __name__ == '__main__':
    main()
