#!/usr/bin/env python
"""Independent dense-SQA replicas sharded over the GPUs of a torchrun job (BASELINE.json config C5a: N=1024, m=128,
4096 replicas on 8 B200 -> 512 per GPU).  J (4 MiB) is L2 resident, so this is not an HBM-bound workload.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29514 \
        benchmarks/replicas.py --replicas-per-gpu 512 --steps 8
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
    ap.add_argument('--N', type=int, default=1024)
    ap.add_argument('--m', type=int, default=128)
    ap.add_argument('--replicas-per-gpu', type=int, default=512)
    ap.add_argument('--steps', type=int, default=8)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import sqaod_b200 as sq
    from sqaod_b200.multigpu import anneal_replicas
    sq.set_active_device(sq.Device(local))
    rng = np.random.default_rng(1133557)
    A = rng.random((args.N, args.N), dtype=np.float32) - np.float32(0.5)
    W = np.triu(A) + np.triu(A, 1).T
    R = args.replicas_per_gpu * world
    Gs = [5.0 * (0.01 / 5.0) ** (k / max(args.steps - 1, 1)) for k in range(args.steps)]
    anneal_replicas(W, world, Gs[:2], 50.0, np.float32, n_trotters=args.m)       # warm-up (one replica per rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    best, local_best, best_id, _ = anneal_replicas(W, R, Gs, 50.0, np.float32, n_trotters=args.m)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        attempts = float(R) * args.steps * args.N * args.m
        print(json.dumps({'row': 'replica batch (C5a)', 'N': args.N, 'm': args.m, 'replicas': R, 'n_gpus': world, 'steps_per_replica': args.steps,
                          'seconds': dt, 'attempts_per_s': attempts / dt, 'best_E': best,
                          'note': 'wall clock incl. seed/prepare/randomize/get_E per replica; J resident per GPU'}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
