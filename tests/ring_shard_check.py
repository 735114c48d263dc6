#!/usr/bin/env python
"""Exact-chain check of the ring-sharded dense annealer on G >= 2 GPUs (run under torchrun; driven by
tests/test_ring_shard_gpu.py).  Every rank anneals m/G trotters of one ring; the gathered spins must equal the CPU
oracle's trajectory for the same Philox seed -- i.e. sharding does not change the chain."""
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
    from sqaod_b200.multigpu import RingShardedDenseAnnealer
    from oracle import pyoracle as orc
    sq.set_active_device(sq.Device(local))
    cases = [(40, 2 * world, 4), (100, 6 * world, 3), (300, 20 * world, 2), (1000, 150 * world, 1), (2100, 8 * world, 1),
             (64, 450 * world, 2), (36, 1001 * world, 1)]   # 3-4 and 6-7 trotters per CTA: the wide warp layout
    ok_all = True
    for N, m, steps in cases:
        for dtype in (np.float32, np.float64):
            W = quantized_symmetric_W(N, 4000 + N, dtype)
            done = False
            for seed in range(3, 20):
                ring = RingShardedDenseAnnealer(W, 0, dtype, n_trotters=m)
                ring.seed(seed); ring.prepare(); ring.randomize_spin()
                ref = orc.DenseGraphAnnealer(W, 0, dtype, n_trotters=m, algorithm='coloring', rng='philox')
                ref.seed(seed); ref.prepare(); ref.randomize_spin()
                assert np.array_equal(ring.gather_spins(), ref.get_q()), 'randomize differs'
                G, beta = 3.0, 1. / 0.3
                border = False
                for s in range(steps):
                    ring.anneal_one_step(G, beta); ref.anneal_one_step(G, beta)
                    G *= 0.7
                    got, want = ring.gather_spins(), ref.get_q()
                    if not np.array_equal(got, want):
                        if ref.stats()[1] > 0:      # an accept test on the rounding edge may legitimately differ: next seed
                            border = True
                            break
                        ok_all = False
                        if rank == 0:
                            bad = np.argwhere(got != want)
                            print('MISMATCH N=%d m=%d dtype=%s seed=%d step=%d: %d spins, trotters %s' % (
                                N, m, np.dtype(dtype).name, seed, s, len(bad), sorted(set(bad[:, 0].tolist()))[:12]), flush=True)
                        break
                del ring
                if not border:
                    done = True
                    break
            if rank == 0:
                print('case N=%d m=%d %s: %s' % (N, m, np.dtype(dtype).name, 'ok' if (done and ok_all) else 'FAILED'), flush=True)
            if not done:
                ok_all = False
    # ---- the C5b row length: N = 32768 (J = 4 GiB per GPU, generated in place on every GPU), a few trotters per GPU, one step
    # against the oracle (rank 0 runs it; the host copy of W comes from the same device generator)
    N, m = 32768, 2 * world
    gen = sq.dense_graph_annealer(None, sq.minimize, np.float32)
    done = False
    for seed in (1, 2, 3):
        ring = RingShardedDenseAnnealer(('random', N, 32768, True), 0, np.float32, n_trotters=m)
        ring.seed(seed); ring.prepare(); ring.randomize_spin()
        ring.anneal_one_step(0.5, 20.0)
        got = ring.gather_spins()
        verdict = np.zeros(1, np.int64)
        if rank == 0:
            W = gen.get_qubo_random(N, 32768, True)
            ref = orc.DenseGraphAnnealer(W, 0, np.float32, n_trotters=m, algorithm='coloring', n_workers=orc.num_threads(), rng='philox')
            ref.seed(seed); ref.prepare(); ref.randomize_spin()
            ref.anneal_one_step(0.5, 20.0)
            same = np.array_equal(got, ref.get_q())
            verdict[0] = 1 if same else (2 if ref.stats()[1] > 0 else 0)   # 1 equal, 2 borderline accept test: next seed, 0 mismatch
            del W, ref
        v = torch.from_numpy(verdict).cuda()
        dist.broadcast(v, 0)
        del ring
        if int(v.item()) == 1:
            done = True
            break
        if int(v.item()) == 0:
            break
    if rank == 0:
        print('case N=%d m=%d float32 (C5b row length): %s' % (N, m, 'ok' if done else 'FAILED'), flush=True)
    if not done:
        ok_all = False
    dist.barrier()
    if rank == 0:
        print('RING_SHARD_OK' if ok_all else 'RING_SHARD_FAILED', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
