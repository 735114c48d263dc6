"""Python solver classes of the B200 back end.

Same public surface as the reference's `sqaod.cuda` package (sqaodpy/sqaod/cuda/*.py + sqaodpy/sqaod/common/*_base.py):
factories dense_graph_annealer / bipartite_graph_annealer / dense_graph_bf_searcher / bipartite_graph_bf_searcher and
the methods of their *Base classes, each forwarding to one C-ABI call (include/sqaod_b200.h) where the reference
forwards to `self._cext.<fn>(self._cobj, ..., self.dtype)`.
"""
import ctypes as C
import numpy as np
from . import _lib
from . import common
from . import device as _device
from .common import minimize, maximize, algorithm  # noqa: F401

L = _lib.lib
ptr = _lib.ptr


def _prefs_from_string(s):
    out = {}
    for item in s.split(';'):
        if not item:
            continue
        k, v = item.split('=', 1)
        out[k] = int(v) if k in ('n_trotters', 'tile_size', 'tile_size_0', 'tile_size_1', 'experiment') else v
    return out


class _SolverBase(object):
    _prefix = None

    def _fn(self, name):
        return getattr(L, 'sqb_%s_%s' % (self._prefix, name))

    def _init_native(self, dtype, dev):
        self.dtype = np.dtype(dtype).type
        self._dt = _lib.dtype_code(dtype)
        self._cobj = C.c_void_p()
        _lib.check(self._fn('new')(C.byref(self._cobj), self._dt))
        self._device = dev if dev is not None else _device.active_device()
        _lib.check(self._fn('assign_device')(self._cobj, self._device._cobj, self._dt))

    def __del__(self):
        try:
            if getattr(self, '_cobj', None):
                self._fn('delete')(self._cobj, self._dt)
                self._cobj = None
        except Exception:
            pass

    def set_preferences(self, prefdict=None, **prefs):
        merged = {}
        if prefdict is not None:
            merged.update(prefdict)
        merged.update(prefs)
        for k, v in merged.items():
            if isinstance(v, str):
                _lib.check(self._fn('set_preference')(self._cobj, k.encode(), v.encode(), C.c_long(0), self._dt))
            else:
                _lib.check(self._fn('set_preference')(self._cobj, k.encode(), None, C.c_long(int(v)), self._dt))

    def get_preferences(self):
        buf = C.create_string_buffer(512)
        _lib.check(self._fn('get_preferences')(self._cobj, buf, 512, self._dt))
        return _prefs_from_string(buf.value.decode())

    def get_optimize_dir(self):
        return self._optimize

    def prepare(self):
        _lib.check(self._fn('prepare')(self._cobj, self._dt))

    def calculate_E(self):
        _lib.check(self._fn('calculate_E')(self._cobj, self._dt))

    def make_solution(self):
        _lib.check(self._fn('make_solution')(self._cobj, self._dt))


# ------------------------------------------------------------------------------------------------ dense annealer
class DenseGraphAnnealer(_SolverBase):
    """sqaod.cuda.DenseGraphAnnealer (common/dense_graph_annealer_base.py:7-101)."""
    _prefix = 'dg_annealer'

    def __init__(self, W=None, optimize=minimize, dtype=np.float64, prefdict=None, device=None):
        self._init_native(dtype, device)
        self._optimize = optimize
        if W is not None:
            self.set_qubo(W, optimize)
        self.set_preferences(prefdict)

    def seed(self, seed):
        _lib.check(L.sqb_dg_annealer_seed(self._cobj, C.c_ulonglong(seed), self._dt))

    def set_qubo(self, W, optimize=minimize):
        common.check_dense_qubo(W)
        W = common.symmetrize(common.fix_type(W, self.dtype))
        W = np.ascontiguousarray(W)
        _lib.check(L.sqb_dg_annealer_set_qubo(self._cobj, ptr(W), W.shape[0], W.strides[0] // W.itemsize, int(optimize), self._dt))
        self._optimize = optimize

    def set_qubo_random(self, N, seed, quantize=False, optimize=minimize):
        """synthetic symmetric W ~ U(-0.5, 0.5), generated on the device (no host matrix, no upload); see get_qubo_random."""
        _lib.check(L.sqb_dg_annealer_set_qubo_random(self._cobj, int(N), C.c_ulonglong(int(seed)), int(bool(quantize)), int(optimize), self._dt))
        self._optimize = optimize

    def get_qubo_random(self, N, seed, quantize=False):
        """the matrix set_qubo_random(N, seed, quantize) anneals, as a host array"""
        W = np.empty((N, N), self.dtype)
        _lib.check(L.sqb_dg_annealer_get_qubo_random(self._cobj, ptr(W), int(N), int(N), C.c_ulonglong(int(seed)), int(bool(quantize)), self._dt))
        return W

    def set_qubo_batch(self, Ws, optimize=minimize):
        """a batch of DIFFERENT problems of one size, annealed side by side in one launch per step (problem r uses seed + r).
        Ws: array (n_problems, N, N) of symmetric matrices.  get_E / get_spins / get_q then return n_problems * n_trotters rows,
        problem-major."""
        Ws = np.ascontiguousarray(np.asarray(Ws, self.dtype))
        if Ws.ndim != 3 or Ws.shape[1] != Ws.shape[2]:
            raise ValueError('Ws must have shape (n_problems, N, N)')
        for W in Ws:
            common.check_dense_qubo(W)
        _lib.check(L.sqb_dg_annealer_set_qubo_batch(self._cobj, ptr(Ws), Ws.shape[0], Ws.shape[1], Ws.shape[2], int(optimize), self._dt))
        self._optimize = optimize

    def set_hamiltonian(self, h, J, c):
        common.check_dense_hJc(h, J, c)
        h, J = common.fix_type([h, J], self.dtype)
        J = np.ascontiguousarray(common.symmetrize(J))
        _lib.check(L.sqb_dg_annealer_set_hamiltonian(self._cobj, ptr(h), ptr(J), J.shape[0], J.strides[0] // J.itemsize,
                                                     C.c_double(float(self.dtype(c))), self._dt))
        self._optimize = minimize

    def get_problem_size(self):
        n = C.c_int(0)
        _lib.check(L.sqb_dg_annealer_get_problem_size(self._cobj, C.byref(n), self._dt))
        return n.value

    def _m(self):
        """rows of the spin / energy buffers: n_trotters (x n_replicas for a replica batch)"""
        m = C.c_int(0); r = C.c_int(1)
        _lib.check(L.sqb_dg_annealer_get_num_trotters(self._cobj, C.byref(m), self._dt))
        _lib.check(L.sqb_dg_annealer_get_num_replicas(self._cobj, C.byref(r), self._dt))
        return m.value * r.value

    def set_replicas(self, n_replicas):
        """anneal n_replicas independent replicas (seed, seed+1, ...) side by side; call before prepare()."""
        _lib.check(L.sqb_dg_annealer_set_num_replicas(self._cobj, int(n_replicas), self._dt))

    def get_hamiltonian(self):
        N = self.get_problem_size()
        h = np.empty(N, self.dtype); J = np.empty((N, N), self.dtype); c = np.empty(1, self.dtype)
        _lib.check(L.sqb_dg_annealer_get_hamiltonian(self._cobj, ptr(h), ptr(J), N, ptr(c), self._dt))
        return h, J, c[0]

    def get_E(self):
        m = self._m()
        E = np.empty(m, self.dtype)
        _lib.check(L.sqb_dg_annealer_get_E(self._cobj, ptr(E), m, self._dt))
        return E

    def _bits(self, fn):
        m, N = self._m(), self.get_problem_size()
        out = np.empty((m, N), np.int8)
        _lib.check(fn(self._cobj, ptr(out), self._dt))
        return [out[i] for i in range(m)]     # the reference returns a list of m int8 arrays (annealer.inc:350-364)

    def get_x(self):
        return self._bits(L.sqb_dg_annealer_get_x)

    def get_q(self):
        return self._bits(L.sqb_dg_annealer_get_q)

    def get_spins(self, out=None):
        """m x N int8 matrix straight from the device (no solution-list bookkeeping); `out` may be a caller-owned
        (e.g. pinned) C-contiguous int8 array of that shape, which is then filled in place."""
        shape = (self._m(), self.get_problem_size())
        if out is None:
            out = np.empty(shape, np.int8)
        elif out.shape != shape or out.dtype != np.int8 or not out.flags.c_contiguous:
            raise ValueError('out must be a C-contiguous int8 array of shape %s' % (shape,))
        _lib.check(L.sqb_dg_annealer_get_spins(self._cobj, ptr(out), self._dt))
        return out

    def set_q(self, q):
        q = common.fix_type(np.asarray(q), np.int8)
        _lib.check(L.sqb_dg_annealer_set_q(self._cobj, ptr(q), q.shape[0], self._dt))

    def set_qset(self, qset):
        if isinstance(qset, np.ndarray) and qset.ndim == 2 and qset.dtype == np.int8 and qset.flags.c_contiguous:
            q = qset            # already an m x N int8 matrix (possibly pinned): copied to the device as it is
        else:
            q = np.ascontiguousarray(np.stack([np.asarray(v, np.int8) for v in qset]))
        _lib.check(L.sqb_dg_annealer_set_qset(self._cobj, ptr(q), q.shape[0], q.shape[1], self._dt))

    def randomize_spin(self):
        _lib.check(L.sqb_dg_annealer_randomize_spin(self._cobj, self._dt))

    def get_system_E(self, G, beta):
        G, beta = self.dtype(G), self.dtype(beta)
        E = C.c_double(0)
        _lib.check(L.sqb_dg_annealer_get_system_E(self._cobj, C.c_double(float(G)), C.c_double(float(beta)), C.byref(E), self._dt))
        return self.dtype(E.value)

    def anneal_one_step(self, G, beta):
        G, beta = self.dtype(G), self.dtype(beta)   # dense_graph_annealer_base.py:99-101
        _lib.check(L.sqb_dg_annealer_anneal_one_step(self._cobj, C.c_double(float(G)), C.c_double(float(beta)), self._dt))

    # ---- ring sharding over several GPUs (sqaod_b200.multigpu.RingShardedDenseAnnealer drives these) ----
    def ring_configure(self, rank, world, m_global):
        _lib.check(L.sqb_dg_annealer_ring_configure(self._cobj, int(rank), int(world), int(m_global), self._dt))

    def ring_export(self):
        buf = (C.c_ubyte * 64)()
        _lib.check(L.sqb_dg_annealer_ring_export(self._cobj, buf, self._dt))
        return bytes(buf)

    def ring_attach(self, left, right):
        lb = (C.c_ubyte * 64).from_buffer_copy(left); rb = (C.c_ubyte * 64).from_buffer_copy(right)
        _lib.check(L.sqb_dg_annealer_ring_attach(self._cobj, lb, rb, self._dt))

    def ring_push_halos(self):
        _lib.check(L.sqb_dg_annealer_ring_push_halos(self._cobj, self._dt))

    def set_sweep_mode(self, mode='auto', field_refresh=0):
        """'classic': one J row per attempt; 'field': local fields in shared memory, one J row per accepted flip; 'auto'."""
        code = {'auto': -1, 'classic': 0, 'field': 1}[mode]
        _lib.check(L.sqb_dg_annealer_set_sweep_mode(self._cobj, code, int(field_refresh), self._dt))

    def get_sweep_mode(self):
        v = C.c_int(0)
        _lib.check(L.sqb_dg_annealer_get_sweep_mode(self._cobj, C.byref(v), self._dt))
        return 'field' if v.value else 'classic'

    def get_fields(self):
        """field mode with carried fields: H[y][j] = h[j] + 2 sum_i J[j][i] q[y][i] as the sweep left it, or None"""
        H = np.empty((self._m(), self.get_problem_size()), self.dtype); ok = C.c_int(0)
        _lib.check(L.sqb_dg_annealer_get_fields(self._cobj, ptr(H), H.shape[1], C.byref(ok), self._dt))
        return H if ok.value else None

    def get_cta_profile(self):
        """per-CTA profile of the last field-mode sweep: array (n_ctas, 8), see sqb_dg_annealer_get_cta_profile"""
        out = np.zeros((512, 16), np.uint64); n = C.c_int(0)
        _lib.check(L.sqb_dg_annealer_get_cta_profile(self._cobj, ptr(out), 512, C.byref(n), self._dt))
        return out[:n.value]

    def get_stats(self):
        a = C.c_ulonglong(0); w = C.c_ulonglong(0)
        _lib.check(L.sqb_dg_annealer_get_stats(self._cobj, C.byref(a), C.byref(w), self._dt))
        d = C.c_ulonglong(0); c = C.c_ulonglong(0)
        _lib.check(L.sqb_dg_annealer_get_barrier_cycles(self._cobj, C.byref(d), C.byref(c), self._dt))
        raw = (C.c_ulonglong * 8)()
        _lib.check(L.sqb_dg_annealer_get_counters(self._cobj, raw, self._dt))
        prof = (C.c_ulonglong * 16)()
        _lib.check(L.sqb_dg_annealer_get_profile(self._cobj, prof, self._dt))
        return {'accepted': a.value, 'flag_waits': w.value, 'barrier_cycles_dot': d.value, 'barrier_cycles_chain': c.value,
                'helper_cycles': raw[5], 'prep_cycles': raw[6],
                'chain_wait_rows_cycles': raw[4], 'chain_wait_neighbour_cycles': raw[7],
                'chain_gather_wait_cycles': prof[8], 'chain_idle_cycles': prof[9], 'chain_barrier_cycles': prof[10],
                'chain_eval_passes': prof[11], 'chain_uncertain_resolves': prof[12], 'chain_commit_resolves': prof[13],
                'chain_blocked_stops': prof[14], 'chain_barrier_cycles_others': prof[15]}


def dense_graph_annealer(W=None, optimize=minimize, dtype=np.float64, device=None, **prefs):
    """factory, as sqaod.cuda.dense_graph_annealer (cuda/dense_graph_annealer.py:18-31)."""
    return DenseGraphAnnealer(W, optimize, dtype, prefs, device)


# ------------------------------------------------------------------------------------------------ bipartite annealer
class BipartiteGraphAnnealer(_SolverBase):
    """sqaod.cuda.BipartiteGraphAnnealer (common/bipartite_graph_annealer_base.py)."""
    _prefix = 'bg_annealer'

    def __init__(self, b0=None, b1=None, W=None, optimize=minimize, dtype=np.float64, prefdict=None, device=None):
        self._init_native(dtype, device)
        self._optimize = optimize
        if W is not None:
            self.set_qubo(b0, b1, W, optimize)
        self.set_preferences(prefdict)

    def seed(self, seed):
        _lib.check(L.sqb_bg_annealer_seed(self._cobj, C.c_ulonglong(seed), self._dt))

    def set_qubo(self, b0, b1, W, optimize=minimize):
        b0, b1, W = common.fix_type([b0, b1, W], self.dtype)
        common.check_bipartite_qubo(b0, b1, W)
        _lib.check(L.sqb_bg_annealer_set_qubo(self._cobj, ptr(b0), ptr(b1), ptr(W), b0.shape[0], b1.shape[0],
                                              W.strides[0] // W.itemsize, int(optimize), self._dt))
        self._optimize = optimize

    def set_hamiltonian(self, h0, h1, J, c):
        h0, h1, J = common.fix_type([h0, h1, J], self.dtype)
        common.check_bipartite_qubo(h0, h1, J)
        _lib.check(L.sqb_bg_annealer_set_hamiltonian(self._cobj, ptr(h0), ptr(h1), ptr(J), h0.shape[0], h1.shape[0],
                                                     J.strides[0] // J.itemsize, C.c_double(float(self.dtype(c))), self._dt))
        self._optimize = minimize

    def get_problem_size(self):
        n0 = C.c_int(0); n1 = C.c_int(0)
        _lib.check(L.sqb_bg_annealer_get_problem_size(self._cobj, C.byref(n0), C.byref(n1), self._dt))
        return n0.value, n1.value

    def _m(self):
        m = C.c_int(0)
        _lib.check(L.sqb_bg_annealer_get_num_trotters(self._cobj, C.byref(m), self._dt))
        return m.value

    def get_hamiltonian(self):
        N0, N1 = self.get_problem_size()
        h0 = np.empty(N0, self.dtype); h1 = np.empty(N1, self.dtype)
        J = np.empty((N1, N0), self.dtype); c = np.empty(1, self.dtype)
        _lib.check(L.sqb_bg_annealer_get_hamiltonian(self._cobj, ptr(h0), ptr(h1), ptr(J), N0, ptr(c), self._dt))
        return h0, h1, J, c[0]

    def get_E(self):
        m = self._m()
        E = np.empty(m, self.dtype)
        _lib.check(L.sqb_bg_annealer_get_E(self._cobj, ptr(E), m, self._dt))
        return E

    def _pairs(self, fn):
        m = self._m(); N0, N1 = self.get_problem_size()
        a = np.empty((m, N0), np.int8); b = np.empty((m, N1), np.int8)
        _lib.check(fn(self._cobj, ptr(a), ptr(b), self._dt))
        return [(a[i], b[i]) for i in range(m)]   # list of m tuples (annealer.inc:553, :679)

    def get_x(self):
        return self._pairs(L.sqb_bg_annealer_get_x)

    def get_q(self):
        return self._pairs(L.sqb_bg_annealer_get_q)

    def set_q(self, qpair):
        q0 = common.fix_type(np.asarray(qpair[0]), np.int8); q1 = common.fix_type(np.asarray(qpair[1]), np.int8)
        _lib.check(L.sqb_bg_annealer_set_q(self._cobj, ptr(q0), ptr(q1), q0.shape[0], q1.shape[0], self._dt))

    def set_qset(self, qpairs):
        q0 = np.ascontiguousarray(np.stack([np.asarray(p[0], np.int8) for p in qpairs]))
        q1 = np.ascontiguousarray(np.stack([np.asarray(p[1], np.int8) for p in qpairs]))
        _lib.check(L.sqb_bg_annealer_set_qset(self._cobj, ptr(q0), ptr(q1), q0.shape[0], q0.shape[1], q1.shape[1], self._dt))

    def randomize_spin(self):
        _lib.check(L.sqb_bg_annealer_randomize_spin(self._cobj, self._dt))

    def get_system_E(self, G, beta):
        G, beta = self.dtype(G), self.dtype(beta)
        E = C.c_double(0)
        _lib.check(L.sqb_bg_annealer_get_system_E(self._cobj, C.c_double(float(G)), C.c_double(float(beta)), C.byref(E), self._dt))
        return self.dtype(E.value)

    def anneal_one_step(self, G, beta):
        G, beta = self.dtype(G), self.dtype(beta)
        _lib.check(L.sqb_bg_annealer_anneal_one_step(self._cobj, C.c_double(float(G)), C.c_double(float(beta)), self._dt))


def bipartite_graph_annealer(b0=None, b1=None, W=None, optimize=minimize, dtype=np.float64, device=None, **prefs):
    return BipartiteGraphAnnealer(b0, b1, W, optimize, dtype, prefs, device)


# ------------------------------------------------------------------------------------------------ dense BF searcher
class DenseGraphBFSearcher(_SolverBase):
    """sqaod.cuda.DenseGraphBFSearcher (common/dense_graph_bf_searcher_base.py)."""
    _prefix = 'dg_bf_searcher'

    def __init__(self, W=None, optimize=minimize, dtype=np.float64, prefdict=None, device=None):
        self._init_native(dtype, device)
        self._optimize = optimize
        if W is not None:
            self.set_qubo(W, optimize)
        self.set_preferences(prefdict)

    def set_qubo(self, W, optimize=minimize):
        common.check_dense_qubo(W)
        W = np.ascontiguousarray(common.symmetrize(common.fix_type(W, self.dtype)))
        _lib.check(L.sqb_dg_bf_searcher_set_qubo(self._cobj, ptr(W), W.shape[0], W.strides[0] // W.itemsize, int(optimize), self._dt))
        self._optimize = optimize

    def get_problem_size(self):
        n = C.c_int(0)
        _lib.check(L.sqb_dg_bf_searcher_get_problem_size(self._cobj, C.byref(n), self._dt))
        return n.value

    def _nsol(self):
        n = C.c_int(0)
        _lib.check(L.sqb_dg_bf_searcher_get_num_solutions(self._cobj, C.byref(n), self._dt))
        return n.value

    def get_x(self):
        n, N = self._nsol(), self.get_problem_size()
        x = np.empty((max(n, 1), N), np.int8)
        _lib.check(L.sqb_dg_bf_searcher_get_x(self._cobj, ptr(x), n, self._dt))
        return [x[i] for i in range(n)]

    def get_E(self):
        n = max(self._nsol(), 1)
        E = np.empty(n, self.dtype)
        _lib.check(L.sqb_dg_bf_searcher_get_E(self._cobj, ptr(E), n, self._dt))
        return E

    def search_range(self):
        done = C.c_int(0); x = C.c_ulonglong(0)
        _lib.check(L.sqb_dg_bf_searcher_search_range(self._cobj, C.byref(done), C.byref(x), self._dt))
        return bool(done.value), x.value

    def search(self):
        # the loop lives in Python in the reference too, so Ctrl-C works (dense_graph_bf_searcher_base.py:57-70)
        self.prepare()
        while not self.search_range()[0]:
            pass
        self.make_solution()

    # ---- sharding extras ----
    def set_range(self, x_begin, x_end):
        _lib.check(L.sqb_dg_bf_searcher_set_range(self._cobj, C.c_ulonglong(x_begin), C.c_ulonglong(x_end), self._dt))

    def get_Emin(self):
        e = C.c_double(0)
        _lib.check(L.sqb_dg_bf_searcher_get_Emin(self._cobj, C.byref(e), self._dt))
        return e.value

    def get_packed_x(self):
        n = C.c_int(0)
        _lib.check(L.sqb_dg_bf_searcher_get_packed_x(self._cobj, None, 0, C.byref(n), self._dt))
        x = np.empty(max(n.value, 1), np.uint64)
        _lib.check(L.sqb_dg_bf_searcher_get_packed_x(self._cobj, ptr(x), n.value, C.byref(n), self._dt))
        return x[:n.value]

    def set_packed_solutions(self, Emin, xs):
        xs = np.ascontiguousarray(xs, np.uint64)
        _lib.check(L.sqb_dg_bf_searcher_set_packed_solutions(self._cobj, C.c_double(Emin), ptr(xs), xs.shape[0], self._dt))


def dense_graph_bf_searcher(W=None, optimize=minimize, dtype=np.float64, device=None, **prefs):
    return DenseGraphBFSearcher(W, optimize, dtype, prefs, device)


# ------------------------------------------------------------------------------------------------ bipartite BF searcher
class BipartiteGraphBFSearcher(_SolverBase):
    """sqaod.cuda.BipartiteGraphBFSearcher (common/bipartite_graph_bf_searcher_base.py)."""
    _prefix = 'bg_bf_searcher'

    def __init__(self, b0=None, b1=None, W=None, optimize=minimize, dtype=np.float64, prefdict=None, device=None):
        self._init_native(dtype, device)
        self._optimize = optimize
        if W is not None:
            self.set_qubo(b0, b1, W, optimize)
        self.set_preferences(prefdict)

    def set_qubo(self, b0, b1, W, optimize=minimize):
        b0, b1, W = common.fix_type([b0, b1, W], self.dtype)
        common.check_bipartite_qubo(b0, b1, W)
        _lib.check(L.sqb_bg_bf_searcher_set_qubo(self._cobj, ptr(b0), ptr(b1), ptr(W), b0.shape[0], b1.shape[0],
                                                 W.strides[0] // W.itemsize, int(optimize), self._dt))
        self._optimize = optimize

    def get_problem_size(self):
        n0 = C.c_int(0); n1 = C.c_int(0)
        _lib.check(L.sqb_bg_bf_searcher_get_problem_size(self._cobj, C.byref(n0), C.byref(n1), self._dt))
        return n0.value, n1.value

    def _nsol(self):
        n = C.c_int(0)
        _lib.check(L.sqb_bg_bf_searcher_get_num_solutions(self._cobj, C.byref(n), self._dt))
        return n.value

    def get_x(self):
        n = self._nsol(); N0, N1 = self.get_problem_size()
        x0 = np.empty((max(n, 1), N0), np.int8); x1 = np.empty((max(n, 1), N1), np.int8)
        _lib.check(L.sqb_bg_bf_searcher_get_x(self._cobj, ptr(x0), ptr(x1), n, self._dt))
        return [(x0[i], x1[i]) for i in range(n)]

    def get_E(self):
        n = max(self._nsol(), 1)
        E = np.empty(n, self.dtype)
        _lib.check(L.sqb_bg_bf_searcher_get_E(self._cobj, ptr(E), n, self._dt))
        return E

    def search_range(self):
        done = C.c_int(0); x0 = C.c_ulonglong(0); x1 = C.c_ulonglong(0)
        _lib.check(L.sqb_bg_bf_searcher_search_range(self._cobj, C.byref(done), C.byref(x0), C.byref(x1), self._dt))
        return bool(done.value), x0.value, x1.value

    def search(self):
        self.prepare()
        while not self.search_range()[0]:
            pass
        self.make_solution()


def bipartite_graph_bf_searcher(b0=None, b1=None, W=None, optimize=minimize, dtype=np.float64, device=None, **prefs):
    return BipartiteGraphBFSearcher(b0, b1, W, optimize, dtype, prefs, device)
