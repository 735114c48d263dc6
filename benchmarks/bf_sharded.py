#!/usr/bin/env python
"""Dense brute-force search with the x range sharded over the ranks of a torchrun job (BASELINE.json config C4:
N=40 fp32, NCCL min-reduce).  Rank 0 prints one JSON line; with --check the result is also compared with an
unsharded search on rank 0 (only sensible for small N).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29512 \
        benchmarks/bf_sharded.py --N 40
"""
import argparse
import json
import os
import sys
import time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--N', type=int, default=40)
    ap.add_argument('--check', action='store_true')
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import sqaod_b200 as sq
    from sqaod_b200.multigpu import sharded_dense_bf_search
    sq.set_active_device(sq.Device(local))
    N = args.N
    rng = np.random.default_rng(1133557)
    A = np.rint((rng.random((N, N)) - 0.5) * 16384) / 16384.          # 2^-14 grid: sums exact
    W = np.asarray(np.triu(A) + np.triu(A, 1).T, np.float32)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    E, xs = sharded_dense_bf_search(W, sq.minimize, np.float32)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    ok = None
    if args.check and rank == 0:
        s = sq.dense_graph_bf_searcher(W, sq.minimize, np.float32)
        s.search()
        ok = bool(s.get_E()[0] == E and np.array_equal(np.stack(s.get_x()), xs))
    if rank == 0:
        print(json.dumps({'row': 'a11 sharded dense brute force', 'N': N, 'n_gpus': world, 'seconds': dt, 'states_per_s': 2.0 ** N / dt,
                          'E_min': float(E), 'x': [''.join(str(int(b)) for b in x) for x in xs[:4]], 'n_solutions': len(xs),
                          'matches_single_gpu': ok}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
