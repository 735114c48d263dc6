#!/usr/bin/env python
"""Size sweep in the protocol of the reference's own benchmark (sqaodpy/benchmark/annealer.py:16-55, benchmark.py:7-51):
fp32, n_trotters = N, fixed G = 0.01, beta = 1/0.02; seconds per anneal_one_step for dense and bipartite (N0 = N1 = N/2).
One JSON line per size."""
import argparse
import json
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sizes', default='128,256,512,1024,2048,4096')
    ap.add_argument('--dtype', default='float32')
    args = ap.parse_args()
    import torch
    import sqaod_b200 as sq
    dev = sq.Device(0)
    sq.set_active_device(dev)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dev.set_stream(stream.cuda_stream)
    dtype = np.dtype(args.dtype).type
    rng = np.random.default_rng(7)

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for N in [int(v) for v in args.sizes.split(',')]:
        A = rng.random((N, N)) - 0.5
        W = np.asarray(np.triu(A) + np.triu(A, 1).T, dtype)
        ann = sq.dense_graph_annealer(W, sq.minimize, dtype, n_trotters=N)
        ann.seed(1); ann.prepare(); ann.randomize_spin()
        reps = max(3, min(200, int(2e9 / (N * N * N))))
        ms = timed(lambda: ann.anneal_one_step(0.01, 50.0), reps)
        out = {'solver': 'dense', 'N': N, 'm': N, 'dtype': args.dtype, 'ms_per_step': ms, 'attempts_per_s': N * N / ms * 1e3}
        del ann
        N0 = N1 = N // 2
        b0, b1, Wb = (np.asarray(rng.random(N0) - 0.5, dtype), np.asarray(rng.random(N1) - 0.5, dtype),
                      np.asarray(rng.random((N1, N0)) - 0.5, dtype))
        bg = sq.bipartite_graph_annealer(b0, b1, Wb, sq.minimize, dtype, n_trotters=N)
        bg.seed(1); bg.prepare(); bg.randomize_spin()
        msb = timed(lambda: bg.anneal_one_step(0.01, 50.0), 20)
        out.update({'bipartite_ms_per_step': msb, 'bipartite_attempts_per_s': N * N / msb * 1e3})
        del bg
        print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
