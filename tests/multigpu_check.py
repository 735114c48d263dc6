#!/usr/bin/env python
"""Multi-GPU checks of the sharded workloads of SURVEY.md 8e over NCCL (run under torchrun; driven by
tests/test_multigpu_gpu.py): (1) dense brute force N = 40 with the x range sharded over the ranks and the NCCL
all_reduce(MIN) + all_gather merge -- minimum and argmin list identical to an unsharded search and independent of the number of
ranks; (2) replica batches sharded over the ranks (C5a shape) -- replica r equals a single solver seeded seed + r, the global best
is the minimum over the ranks."""
import hashlib
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch
    import torch.distributed as dist
    from conftest import quantized_symmetric_W
    rank = int(os.environ['RANK']); local = int(os.environ['LOCAL_RANK']); world = int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(local)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import sqaod_b200 as sq
    from sqaod_b200 import multigpu
    sq.set_active_device(sq.Device(local))
    ok = True

    # ---- (1) brute force N = 40 (C4), plus a degenerate problem whose argmin list crosses the rank boundaries
    for name, W in (('N=40 quantised', quantized_symmetric_W(40, 40, np.float32)),
                    # x^T W x = 4 k^2 - 28 k for k ones: minimum -48 at k = 3 and k = 4, 1140 + 4845 argmins spread over the whole range
                    ('N=20 degenerate', np.asarray(np.full((20, 20), 4.0) - 28.0 * np.eye(20), np.float32))):
        E, xs = multigpu.sharded_dense_bf_search(W, 0, np.float32)
        one = sq.dense_graph_bf_searcher(W, sq.minimize, np.float32)
        one.search()
        E1, xs1 = one.get_E()[0], np.stack(one.get_x())
        same = (float(E) == float(E1)) and np.array_equal(np.stack(xs), xs1)
        h = hashlib.sha256(np.float64(E).tobytes() + np.asarray(xs, np.int8).tobytes()).hexdigest()[:16]
        hs = [None] * world
        dist.all_gather_object(hs, h)
        same = same and len(set(hs)) == 1
        if rank == 0:
            print('bf %s: E=%r, %d argmins, sha %s, world %d: %s' % (name, float(E), len(xs), h, world, 'ok' if same else 'FAILED'), flush=True)
        ok = ok and same

    # ---- (2) replica batches over the ranks (C5a: N = 1024, m = 128)
    N, m, per = 1024, 128, 24
    W = quantized_symmetric_W(N, 1024, np.float32)
    Gs = [2.0, 0.5, 0.1]
    best, local_best, best_id, best_q = multigpu.anneal_replicas(W, per * world, Gs, 50.0, np.float32, n_trotters=m, base_seed=300)
    begin = rank * per
    r = begin + (per // 2)
    one = sq.dense_graph_annealer(W, sq.minimize, np.float32, n_trotters=m)
    one.seed(300 + r); one.prepare(); one.randomize_spin()
    for G in Gs:
        one.anneal_one_step(G, 50.0)
    same = abs(float(one.get_E().min()) - float(local_best[r - begin])) < 1e-3
    bests = [None] * world
    dist.all_gather_object(bests, float(np.min(local_best)))
    same = same and abs(min(bests) - best) < 1e-9
    oks = [None] * world
    dist.all_gather_object(oks, bool(same))
    if rank == 0:
        print('replicas: %d over %d ranks, best %.4f: %s' % (per * world, world, best, 'ok' if all(oks) else 'FAILED'), flush=True)
    ok = ok and all(oks)
    dist.barrier()
    if rank == 0:
        print('MULTIGPU_OK' if ok else 'MULTIGPU_FAILED', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
