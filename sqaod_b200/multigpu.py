"""Multi-GPU drivers for the parts of the path that shard.  One process per GPU, torch.distributed for
the plumbing (NCCL on the GPUs; the host-side logic also runs under gloo for the CPU tests).

  * brute-force search: the x range is cut into `world` contiguous slabs, every rank searches its slab with the
    single-GPU searcher, then ONE exchange step: all_reduce(MIN) on the minimum (exact: min is order independent),
    all_gather of the argmin lists of the ranks that hold the global minimum, merge ascending, cap.  The result does
    not depend on the number of ranks.
  * annealing replicas: independent seeds per rank, no data-path collective; an optional final MIN-reduce of the best
    energy.
"""
import numpy as np


def shard_range(x_max, rank, world):
    """contiguous slab [begin, end) of [0, x_max) for `rank`; slabs differ by at most one element."""
    base, rem = divmod(int(x_max), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist


def _device_for(group_backend):
    import torch
    return torch.device('cuda', torch.cuda.current_device()) if group_backend == 'nccl' else torch.device('cpu')


def merge_bf_results(local_Emin, local_xs, cap, minimize=True, group=None):
    """The exchange step of the sharded search.  local_Emin: this rank's best energy (user sign), local_xs: its packed
    argmins (uint64).  Returns (Emin, merged ascending packed list capped at `cap`), identical on every rank."""
    import torch
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()):
        xs = np.sort(np.asarray(local_xs, np.uint64))[:cap]
        return float(local_Emin), xs
    world = dist.get_world_size(group)
    dev = _device_for(dist.get_backend(group))
    key = float(local_Emin) if minimize else -float(local_Emin)
    e = torch.tensor([key], dtype=torch.float64, device=dev)
    dist.all_reduce(e, op=dist.ReduceOp.MIN, group=group)
    gmin = float(e.item())
    mine = np.sort(np.asarray(local_xs, np.uint64))[:cap] if key == gmin else np.empty(0, np.uint64)
    n = torch.tensor([len(mine)], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    width = max(max(counts), 1)
    buf = torch.zeros(width, dtype=torch.int64, device=dev)
    if len(mine):
        buf[:len(mine)] = torch.from_numpy(mine.view(np.int64)).to(dev)
    gathered = [torch.zeros(width, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(gathered, buf, group=group)
    parts = [g[:c].cpu().numpy().view(np.uint64) for g, c in zip(gathered, counts) if c]
    xs = np.sort(np.concatenate(parts)) if parts else np.empty(0, np.uint64)
    return (gmin if minimize else -gmin), xs[:cap]


def sharded_dense_bf_search(W, optimize, dtype, cap=1 << 16, group=None, local_search=None, **prefs):
    """Exhaustive search over x in [0, 2^N) split across the ranks of `group`.

    local_search(W, optimize, dtype, x_begin, x_end) -> (Emin, packed xs) may be injected (the CPU tests inject the
    oracle); by default the B200 searcher is used.  Returns (Emin, list of int8 bit vectors, ascending)."""
    dist = _dist()
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    N = W.shape[0]
    begin, end = shard_range(1 << N, rank, world)
    minimize = int(optimize) == 0
    if local_search is None:
        from . import solvers

        def local_search(W_, opt_, dtype_, b, e):
            s = solvers.dense_graph_bf_searcher(W_, opt_, dtype_, **prefs)
            s.set_range(b, e)
            s.prepare()
            while not s.search_range()[0]:
                pass
            return s.get_Emin(), s.get_packed_x()
    if end > begin:
        Emin, xs = local_search(W, optimize, dtype, begin, end)
    else:
        Emin, xs = (np.inf if minimize else -np.inf), np.empty(0, np.uint64)
    Emin, xs = merge_bf_results(Emin, xs, cap, minimize, group)
    from .common import create_bitset_sequence
    return np.dtype(dtype).type(Emin), create_bitset_sequence([int(v) for v in xs], N)


def replica_seed(base_seed, rank, replica):
    """seed of replica `replica` on rank `rank` (replica r of the whole job has seed base + r)."""
    return int(base_seed) + int(replica)


def best_energy_over_ranks(local_best, minimize=True, group=None):
    import torch
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()):
        return float(local_best)
    dev = _device_for(dist.get_backend(group))
    t = torch.tensor([float(local_best)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN if minimize else dist.ReduceOp.MAX, group=group)
    return float(t.item())


def anneal_replicas(W, n_replicas, schedule, beta, dtype=np.float32, n_trotters=None, base_seed=0, optimize=0, group=None,
                    algorithm='coloring'):
    """Independent annealing replicas of ONE problem, sharded over the ranks.

    Replica r of the whole job has seed base_seed + r and runs on rank r // ceil(n_replicas / world); no data-path
    collective.  `schedule` is the list of G (kT for SA) values, one anneal_one_step each.  Returns
    (best energy over all ranks, this rank's per-replica best energies, this rank's best replica id, its best spins)."""
    from . import solvers
    dist = _dist()
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    begin, end = shard_range(n_replicas, rank, world)
    minimize = int(optimize) == 0
    opt = solvers.minimize if minimize else solvers.maximize
    prefs = {'algorithm': algorithm}
    if n_trotters is not None:
        prefs['n_trotters'] = n_trotters
    n_local = end - begin
    best = np.full(max(n_local, 0), np.inf if minimize else -np.inf)
    best_q, best_id = None, -1
    if n_local > 0:
        # one solver per rank: J uploaded once, the rank's replicas annealed side by side (seed base + r for replica r)
        ann = solvers.dense_graph_annealer(W, opt, dtype, **prefs)
        ann.set_replicas(n_local)
        ann.seed(replica_seed(base_seed, rank, begin))
        ann.prepare()
        ann.randomize_spin()
        for G in schedule:
            ann.anneal_one_step(G, beta)
        m = ann.get_preferences()['n_trotters']
        E = ann.get_E().reshape(n_local, m)
        idx = np.argmin(E, axis=1) if minimize else np.argmax(E, axis=1)
        best = E[np.arange(n_local), idx].astype(np.float64)
        k = int(np.argmin(best) if minimize else np.argmax(best))
        best_id = begin + k
        best_q = ann.get_spins().reshape(n_local, m, -1)[k, idx[k]].copy()
    local_best = (best.min() if minimize else best.max()) if len(best) else (np.inf if minimize else -np.inf)
    return best_energy_over_ranks(local_best, minimize, group), best, best_id, best_q


class RingShardedDenseAnnealer(object):
    """ONE dense SQA instance whose trotter ring is split over the ranks of `group`.

    Rank g anneals trotters [g*m/G, (g+1)*m/G); J and h are replicated.  The sweep kernel's inter-CTA hand-off (accept
    flags + per-window snapshots of the edge trotters) is extended across GPUs by mirroring the two edge trotters'
    publications into the neighbouring GPU's memory over NVLink (CUDA IPC peer mappings, system-scope release/acquire),
    and every sweep ends with a push of the edge trotters' spins to the neighbours.  The chain is the same as on one
    GPU: spins are identical to an unsharded run with the same seed (tests/ring_shard_check.py)."""

    def __init__(self, W, optimize=0, dtype=np.float32, n_trotters=None, group=None):
        from . import solvers
        dist = _dist()
        assert dist.is_available() and dist.is_initialized(), 'RingShardedDenseAnnealer needs torch.distributed'
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        opt = solvers.minimize if int(optimize) == 0 else solvers.maximize
        if isinstance(W, tuple) and W[0] == 'random':
            # ('random', N, seed[, quantize]): the same synthetic W generated on every GPU (no host matrix, no upload: 4 GiB at N = 32768)
            n = int(W[1])
            self.ann = solvers.dense_graph_annealer(None, opt, dtype)
            self.ann.set_qubo_random(n, int(W[2]), bool(W[3]) if len(W) > 3 else False, opt)
        else:
            n = W.shape[0]
            self.ann = solvers.dense_graph_annealer(W, opt, dtype)
        self.m = int(n_trotters if n_trotters is not None else n // 4)
        self.ann.ring_configure(self.rank, self.world, self.m)
        self.m_local = self.m // self.world
        self._attached = False

    def seed(self, seed):
        self.ann.seed(seed)          # same seed on every rank: Philox is keyed by the GLOBAL trotter index

    def prepare(self):
        dist = _dist()
        self.ann.prepare()
        handles = [None] * self.world
        dist.all_gather_object(handles, self.ann.ring_export(), group=self.group)
        left, right = handles[(self.rank - 1) % self.world], handles[(self.rank + 1) % self.world]
        self.ann.ring_attach(left, right)
        dist.barrier(group=self.group)
        self._attached = True

    def randomize_spin(self):
        self.ann.randomize_spin()
        self._sync_halos()

    def set_qset(self, q_all):
        """q_all: the whole m x N spin matrix (every rank passes the same array); each rank keeps its rows."""
        q_all = np.asarray(q_all, np.int8)
        lo = self.rank * self.m_local
        n_before = self.ann.get_preferences()['n_trotters']
        self.ann.set_qset(q_all[lo:lo + self.m_local])
        assert self.ann.get_preferences()['n_trotters'] == n_before
        self._sync_halos()

    def _sync_halos(self):
        dist = _dist()
        self.ann._device.synchronize()
        dist.barrier(group=self.group)   # every rank has (re)initialised its hand-off block before anyone pushes into it
        self.ann.ring_push_halos()

    def anneal_one_step(self, G, beta):
        self.ann.anneal_one_step(G, beta)

    def get_local_spins(self):
        return self.ann.get_spins()

    def get_local_E(self):
        return self.ann.get_E()

    def gather_spins(self):
        """the whole m x N spin matrix on every rank"""
        import torch
        dist = _dist()
        dev = _device_for(dist.get_backend(self.group))
        mine = torch.from_numpy(np.ascontiguousarray(self.ann.get_spins())).to(dev)
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        return np.concatenate([p.cpu().numpy() for p in parts], axis=0)

    def best_energy(self, minimize=True):
        E = self.ann.get_E()
        return best_energy_over_ranks(E.min() if minimize else E.max(), minimize, self.group)
