#!/usr/bin/env python
"""One dense SQA instance with its trotter ring sharded over the GPUs of a torchrun job (BASELINE.json config C5b:
N=32768, m=2048 on 8 B200 -> 256 trotters per GPU; J replicated).  Rank 0 prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29515 \
        benchmarks/ring_shard.py --N 32768 --trotters-per-gpu 256 --steps 3
"""
import argparse
import json
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--N', type=int, default=16384)
    ap.add_argument('--trotters-per-gpu', type=int, default=256)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=1)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import sqaod_b200 as sq
    from sqaod_b200.multigpu import RingShardedDenseAnnealer
    dev = sq.Device(local)
    sq.set_active_device(dev)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dev.set_stream(stream.cuda_stream)
    N, m = args.N, args.trotters_per_gpu * world
    rng = np.random.default_rng(1133557)
    W = rng.random((N, N), dtype=np.float32)
    W -= np.float32(0.5)
    iu = np.triu_indices(N, 1)
    W.T[iu] = W[iu]                      # mirror the upper triangle in place (no second 4 GiB temporary)
    os.environ['SQAOD_B200_NO_TC'] = '1' if N > 16384 else os.environ.get('SQAOD_B200_NO_TC', '0')
    ring = RingShardedDenseAnnealer(W, 0, np.float32, n_trotters=m)
    del W
    ring.seed(4242); ring.prepare(); ring.randomize_spin()
    G, beta = 0.01, 50.0
    for _ in range(args.warmup):
        ring.anneal_one_step(G, beta)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        ring.anneal_one_step(G, beta)
    e1.record(stream)
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    best = ring.best_energy()
    st = ring.ann.get_stats()
    if rank == 0:
        attempts = float(N) * m
        print(json.dumps({'row': 'ring-sharded dense SQA (C5b)', 'N': N, 'm': m, 'n_gpus': world, 'trotters_per_gpu': args.trotters_per_gpu,
                          'ms_per_step': ms, 'attempts_per_s': attempts / ms * 1e3,
                          'algorithmic_GBps_per_gpu': attempts / world * N * 4 / ms / 1e6, 'best_E': best,
                          'flag_wait_polls_rank0': st['flag_waits']}), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
