#!/usr/bin/env python
"""Size sweep in the protocol of the reference's own benchmark harness (sqaodpy/benchmark/annealer.py:16-55 and
benchmark.py:7-51): fp32, n_trotters = N, fixed G = 0.01, beta = 1/0.02, dense and bipartite (N0 = N1 = N/2) annealers over
the reference's N list 128 ... 8192; anneal_one_step calls followed by make_solution, wall clock per iteration.  The
reference runs every size for 60 s; `--duration` bounds it (default 1.5 s per size after a short calibration).

Output: one JSON line per size, and with `--csv DIR` the reference's report files (report.py: N, nIters, time per iteration)
`b200_dense_graph.csv` / `b200_bipartite_graph.csv`."""
import argparse
import csv
import json
import os
import sys
import timeit
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_NLIST = [128, 192, 256, 384, 512, 768, 1024, 1280, 1536, 1792, 2048, 2560, 3072, 3574, 4096, 5120, 6144, 7168, 8192]


def anneal(an, duration):
    """benchmark.py:9-51 with the 5 s warm-up / 60 s run scaled down"""
    G, beta = 0.01, 1 / 0.02
    an.prepare()
    an.randomize_spin()
    timer = timeit.default_timer
    n_iters, warm = 3, 0.
    while warm < duration / 6.:
        begin = timer()
        for _ in range(n_iters):
            an.anneal_one_step(G, beta)
        an.make_solution()
        warm = timer() - begin
        if warm < duration / 6.:
            n_iters *= 3
    n_iters = int(duration / warm * n_iters) + 1
    begin = timer()
    for _ in range(n_iters):
        an.anneal_one_step(G, beta)
    an.make_solution()
    elapsed = timer() - begin
    return n_iters, elapsed / n_iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sizes', default=','.join(str(n) for n in REF_NLIST))
    ap.add_argument('--dtype', default='float32')
    ap.add_argument('--duration', type=float, default=1.5)
    ap.add_argument('--csv', default=None, help='directory for the report files')
    ap.add_argument('--skip-bipartite', action='store_true')
    args = ap.parse_args()
    import sqaod_b200 as sq
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    dtype = np.dtype(args.dtype).type
    rng = np.random.default_rng(7)
    dense, bip = [], []
    for N in [int(v) for v in args.sizes.split(',')]:
        W = sq.generate_random_symmetric_W(N, dtype=dtype)
        # the sweep kernels carry at most 32 trotters per CTA: m <= 32 x #SMs (4736 on a B200); beyond that the size runs with that m
        m = min(N, 32 * sms)
        ann = sq.dense_graph_annealer(W, sq.minimize, dtype)
        ann.set_preferences(n_trotters=m)
        n_it, sec = anneal(ann, args.duration)
        out = {'solver': 'dense', 'N': N, 'm': m, 'dtype': args.dtype, 'sweep_mode': ann.get_sweep_mode(), 'n_iters': n_it,
               'ms_per_step': sec * 1e3, 'attempts_per_s': N * m / sec}
        dense.append((N, n_it, sec))
        del ann
        if not args.skip_bipartite:
            N0 = N1 = N // 2
            b0, b1, Wb = (np.asarray(rng.random(N0) - 0.5, dtype), np.asarray(rng.random(N1) - 0.5, dtype),
                          np.asarray(rng.random((N1, N0)) - 0.5, dtype))
            bg = sq.bipartite_graph_annealer(b0, b1, Wb, sq.minimize, dtype)
            bg.set_preferences(n_trotters=N)
            n_it, sec = anneal(bg, args.duration)
            out.update({'bipartite_n_iters': n_it, 'bipartite_ms_per_step': sec * 1e3, 'bipartite_attempts_per_s': N * N / sec})
            bip.append((N, n_it, sec))
            del bg
        print(json.dumps(out), flush=True)
    if args.csv:
        os.makedirs(args.csv, exist_ok=True)
        for name, rows in (('b200_dense_graph.csv', dense), ('b200_bipartite_graph.csv', bip)):
            if rows:
                with open(os.path.join(args.csv, name), 'w', newline='') as f:
                    wr = csv.writer(f)
                    wr.writerow(['N', 'nIters', 'time'])
                    wr.writerows(rows)


if __name__ == '__main__':
    main()
