"""world_size-2 (and 3) gloo tests of the multi-GPU host logic: range sharding + the one-step min/gather merge of the
sharded brute-force search.  The per-rank search is injected (CPU oracle), so no GPU is needed."""
import os
import numpy as np
import pytest
import torch.multiprocessing as mp
from conftest import quantized_symmetric_W


def test_shard_range_covers_exactly_once():
    from sqaod_b200.multigpu import shard_range
    for xmax in (1, 7, 256, (1 << 40), (1 << 40) + 5):
        for world in (1, 2, 3, 8):
            slabs = [shard_range(xmax, r, world) for r in range(world)]
            assert slabs[0][0] == 0 and slabs[-1][1] == xmax
            assert all(slabs[i][1] == slabs[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in slabs]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, N, opt, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import pyoracle as orc
    from sqaod_b200.multigpu import sharded_dense_bf_search, best_energy_over_ranks

    def local_search(W, optimize, dtype, b, e):
        E, xs = orc.dense_graph_bf_search(W, int(optimize), dtype, tile_size=1 << 16, x_begin=b, x_end=e)
        return float(E), xs

    W = quantized_symmetric_W(N, 42) if N != 8 else (np.full((8, 8), 4.0) - 36.0 * np.eye(8))
    E, xs = sharded_dense_bf_search(W, opt, np.float64, cap=1 << 16, local_search=local_search)
    best = best_energy_over_ranks(float(rank), minimize=True)
    out.put((rank, float(E), [int(''.join(str(int(b)) for b in x), 2) for x in xs], best))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
@pytest.mark.parametrize('N,opt', [(12, 0), (12, 1), (8, 0)])
def test_sharded_bf_is_invariant_to_world_size(oracle, world, N, opt):
    W = quantized_symmetric_W(N, 42) if N != 8 else (np.full((8, 8), 4.0) - 36.0 * np.eye(8))
    E0, xs0 = oracle.dense_graph_bf_search(W, opt, np.float64, tile_size=1 << 16)
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = 29600 + world * 10 + N + opt
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, opt, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, E, xs, best in res:
        assert E == float(E0)
        assert xs == [int(v) for v in xs0]
        assert best == 0.0
