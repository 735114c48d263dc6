"""ctypes front end of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package ``sqaod_b200`` never imports this module.

The classes mirror the reference's CPU solvers (sqaodc/cpu/CPU*.cpp) closely enough that parity tests read
like the reference's own (sqaodpy/tests/test_dense_graph_annealer.py etc.).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ALGO = {'naive': 2, 'coloring': 3, 'sa_naive': 6, 'sa_coloring': 7}
RNG_MT, RNG_PHILOX = 0, 1
DOM_DENSE_SWEEP, DOM_RANDOMIZE, DOM_BG_SIDE0, DOM_BG_SIDE1, DOM_RANDOMIZE1 = 0, 1, 2, 3, 4
FLT_MAX = float(np.finfo(np.float32).max)


def build(force=False):
    """Compile oracle/liboracle.so (and oracle/_ref when /root/reference exists)."""
    so = os.path.join(_HERE, 'liboracle.so')
    src = os.path.join(_HERE, 'oracle.cpp')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, 'liboracle.so'], stdout=subprocess.DEVNULL)
    if os.path.isdir('/root/reference/sqaodc'):
        ref = os.path.join(_HERE, '_ref', 'libsqaod_refparts.so')
        if force or not os.path.exists(ref):
            subprocess.check_call(['make', '-C', _HERE, 'ref'], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, 'liboracle.so')
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_dga_new_f32.restype = C.c_void_p
        _LIB.orc_dga_new_f64.restype = C.c_void_p
        _LIB.orc_bga_new_f32.restype = C.c_void_p
        _LIB.orc_bga_new_f64.restype = C.c_void_p
        _LIB.orc_dga_system_E_f32.restype = C.c_float
        _LIB.orc_dga_system_E_f64.restype = C.c_double
        _LIB.orc_bga_system_E_f32.restype = C.c_float
        _LIB.orc_bga_system_E_f64.restype = C.c_double
    return _LIB


def reflib():
    so = os.path.join(_HERE, '_ref', 'libsqaod_refparts.so')
    if not os.path.exists(so):
        return None
    return C.CDLL(so)


def _sfx(dtype):
    return 'f32' if np.dtype(dtype) == np.float32 else 'f64'


def _creal(dtype):
    return C.c_float if np.dtype(dtype) == np.float32 else C.c_double


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _fn(name, dtype):
    return getattr(lib(), '%s_%s' % (name, _sfx(dtype)))


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def num_threads():
    return lib().orc_num_threads()


def philox4x32_10(ctr, key):
    ctr = _arr(ctr, np.uint32); key = _arr(key, np.uint32); out = np.empty(4, np.uint32)
    lib().orc_philox4x32_10(_p(ctr), _p(key), _p(out))
    return out


def sqb_philox(seed, step, dom, idx, y):
    out = np.empty(4, np.uint32)
    lib().orc_sqb_philox(C.c_uint64(seed), C.c_uint64(step), C.c_uint32(dom), C.c_uint32(idx), C.c_uint32(y), _p(out))
    return out


def mt_stream(seed, n):
    out = np.empty(n, np.uint32)
    lib().orc_mt_stream(C.c_uint64(seed), n, _p(out))
    return out


def mt_reals(seed, n):
    f = np.empty(n, np.float32); d = np.empty(n, np.float64)
    lib().orc_mt_reals(C.c_uint64(seed), n, _p(f), _p(d))
    return f, d


# ---------------------------------------------------------------- formulas
def dense_graph_calculate_hamiltonian(W, dtype=np.float64):
    W = _arr(W, dtype); N = W.shape[0]
    h = np.empty(N, dtype); J = np.empty((N, N), dtype); c = np.empty(1, dtype)
    _fn('orc_dg_hamiltonian', dtype)(_p(h), _p(J), _p(c), _p(W), N)
    return h, J, c[0]


def dense_graph_batch_calculate_E(W, x, dtype=np.float64):
    W = _arr(W, dtype); x = _arr(np.atleast_2d(x), np.int8)
    E = np.empty(x.shape[0], dtype)
    _fn('orc_dg_energy_bits', dtype)(_p(E), _p(W), W.shape[0], _p(x), x.shape[0])
    return E


def dense_graph_batch_calculate_E_from_spin(h, J, c, q, dtype=np.float64):
    h = _arr(h, dtype); J = _arr(J, dtype); q = _arr(np.atleast_2d(q), np.int8)
    E = np.empty(q.shape[0], dtype)
    _fn('orc_dg_energy_spins', dtype)(_p(E), _p(h), _p(J), _creal(dtype)(c), J.shape[0], _p(q), q.shape[0])
    return E


def bipartite_graph_calculate_hamiltonian(b0, b1, W, dtype=np.float64):
    b0 = _arr(b0, dtype); b1 = _arr(b1, dtype); W = _arr(W, dtype)
    N1, N0 = W.shape
    h0 = np.empty(N0, dtype); h1 = np.empty(N1, dtype); J = np.empty((N1, N0), dtype); c = np.empty(1, dtype)
    _fn('orc_bg_hamiltonian', dtype)(_p(h0), _p(h1), _p(J), _p(c), _p(b0), _p(b1), _p(W), N0, N1)
    return h0, h1, J, c[0]


def bipartite_graph_batch_calculate_E(b0, b1, W, x0, x1, dtype=np.float64):
    b0 = _arr(b0, dtype); b1 = _arr(b1, dtype); W = _arr(W, dtype)
    x0 = _arr(np.atleast_2d(x0), np.int8); x1 = _arr(np.atleast_2d(x1), np.int8)
    N1, N0 = W.shape
    E = np.empty(x0.shape[0], dtype)
    _fn('orc_bg_energy_bits', dtype)(_p(E), _p(b0), _p(b1), _p(W), N0, N1, _p(x0), _p(x1), x0.shape[0])
    return E


def bipartite_graph_batch_calculate_E_2d(b0, b1, W, x0, x1, dtype=np.float64):
    b0 = _arr(b0, dtype); b1 = _arr(b1, dtype); W = _arr(W, dtype)
    x0 = _arr(np.atleast_2d(x0), np.int8); x1 = _arr(np.atleast_2d(x1), np.int8)
    N1, N0 = W.shape
    E = np.empty((x1.shape[0], x0.shape[0]), dtype)
    _fn('orc_bg_energy_bits_2d', dtype)(_p(E), _p(b0), _p(b1), _p(W), N0, N1, _p(x0), x0.shape[0], _p(x1), x1.shape[0])
    return E


def bipartite_graph_batch_calculate_E_from_spin(h0, h1, J, c, q0, q1, dtype=np.float64):
    h0 = _arr(h0, dtype); h1 = _arr(h1, dtype); J = _arr(J, dtype)
    q0 = _arr(np.atleast_2d(q0), np.int8); q1 = _arr(np.atleast_2d(q1), np.int8)
    N1, N0 = J.shape
    E = np.empty(q0.shape[0], dtype)
    _fn('orc_bg_energy_spins', dtype)(_p(E), _p(h0), _p(h1), _p(J), _creal(dtype)(c), N0, N1, _p(q0), _p(q1), q0.shape[0])
    return E


# ---------------------------------------------------------------- annealers
class DenseGraphAnnealer(object):
    """Restated sqaod.cpu dense-graph annealer (CPUDenseGraphAnnealer.cpp)."""

    def __init__(self, W=None, optimize=0, dtype=np.float64, n_trotters=None, algorithm='coloring',
                 n_workers=1, rng='mt'):
        self.dtype = np.dtype(dtype).type
        self._o = C.c_void_p(_fn('orc_dga_new', dtype)())
        self.algorithm = algorithm
        _fn('orc_dga_config', dtype)(self._o, ALGO[algorithm], n_workers, RNG_PHILOX if rng == 'philox' else RNG_MT)
        self.n_workers = _fn('orc_dga_workers', dtype)(self._o)
        self.N = 0; self.m = n_trotters
        if W is not None:
            self.set_qubo(W, optimize)

    def __del__(self):
        try:
            _fn('orc_dga_delete', self.dtype)(self._o)
        except Exception:
            pass

    def seed(self, s):
        _fn('orc_dga_seed', self.dtype)(self._o, C.c_uint64(s))

    def set_qubo(self, W, optimize=0):
        W = _arr(W, self.dtype); self.N = W.shape[0]
        _fn('orc_dga_set_qubo', self.dtype)(self._o, _p(W), self.N, int(optimize))
        if self.m is None:
            self.m = self.N // 4

    def set_hamiltonian(self, h, J, c):
        h = _arr(h, self.dtype); J = _arr(J, self.dtype); self.N = J.shape[0]
        _fn('orc_dga_set_hamiltonian', self.dtype)(self._o, _p(h), _p(J), _creal(self.dtype)(c), self.N)
        if self.m is None:
            self.m = self.N // 4

    def get_hamiltonian(self):
        h = np.empty(self.N, self.dtype); J = np.empty((self.N, self.N), self.dtype); c = np.empty(1, self.dtype)
        _fn('orc_dga_get_hamiltonian', self.dtype)(self._o, _p(h), _p(J), _p(c))
        return h, J, c[0]

    def prepare(self, n_trotters=None):
        if n_trotters is not None:
            self.m = n_trotters
        _fn('orc_dga_prepare', self.dtype)(self._o, int(self.m))

    def randomize_spin(self):
        _fn('orc_dga_randomize', self.dtype)(self._o)

    def set_qset(self, q):
        q = _arr(q, np.int8)
        if q.shape[0] != self.m:
            self.prepare(q.shape[0])
        _fn('orc_dga_set_q', self.dtype)(self._o, _p(q))

    def set_q(self, q):
        self.set_qset(np.tile(_arr(q, np.int8), (self.m, 1)))

    def get_q(self):
        q = np.empty((self.m, self.N), np.int8)
        _fn('orc_dga_get_q', self.dtype)(self._o, _p(q))
        return q

    def get_x(self):
        return ((self.get_q() + 1) // 2).astype(np.int8)

    def anneal_one_step(self, G, beta):
        r = _creal(self.dtype)
        _fn('orc_dga_anneal_one_step', self.dtype)(self._o, r(G), r(beta))

    def anneal_rounds(self, G, beta, r0, r1):
        r = _creal(self.dtype)
        _fn('orc_dga_anneal_rounds', self.dtype)(self._o, r(G), r(beta), int(r0), int(r1))

    def get_E(self):
        E = np.empty(self.m, self.dtype)
        _fn('orc_dga_calculate_E', self.dtype)(self._o, _p(E))
        return E

    def get_system_E(self, G, beta):
        r = _creal(self.dtype)
        return self.dtype(_fn('orc_dga_system_E', self.dtype)(self._o, r(G), r(beta)))

    def stats(self):
        a = C.c_longlong(0); b = C.c_longlong(0)
        _fn('orc_dga_stats', self.dtype)(self._o, C.byref(a), C.byref(b))
        return a.value, b.value


class BipartiteGraphAnnealer(object):
    """Restated sqaod.cpu bipartite-graph annealer (CPUBipartiteGraphAnnealer.cpp)."""

    def __init__(self, b0=None, b1=None, W=None, optimize=0, dtype=np.float64, n_trotters=None,
                 algorithm='coloring', n_workers=1, rng='mt'):
        self.dtype = np.dtype(dtype).type
        self._o = C.c_void_p(_fn('orc_bga_new', dtype)())
        self.algorithm = algorithm
        _fn('orc_bga_config', dtype)(self._o, ALGO[algorithm], n_workers, RNG_PHILOX if rng == 'philox' else RNG_MT)
        self.N0 = self.N1 = 0; self.m = n_trotters
        if W is not None:
            self.set_qubo(b0, b1, W, optimize)

    def __del__(self):
        try:
            _fn('orc_bga_delete', self.dtype)(self._o)
        except Exception:
            pass

    def seed(self, s):
        _fn('orc_bga_seed', self.dtype)(self._o, C.c_uint64(s))

    def set_qubo(self, b0, b1, W, optimize=0):
        b0 = _arr(b0, self.dtype); b1 = _arr(b1, self.dtype); W = _arr(W, self.dtype)
        self.N1, self.N0 = W.shape
        _fn('orc_bga_set_qubo', self.dtype)(self._o, _p(b0), _p(b1), _p(W), self.N0, self.N1, int(optimize))
        if self.m is None:
            self.m = (self.N0 + self.N1) // 4

    def set_hamiltonian(self, h0, h1, J, c):
        h0 = _arr(h0, self.dtype); h1 = _arr(h1, self.dtype); J = _arr(J, self.dtype)
        self.N1, self.N0 = J.shape
        _fn('orc_bga_set_hamiltonian', self.dtype)(self._o, _p(h0), _p(h1), _p(J), _creal(self.dtype)(c), self.N0, self.N1)
        if self.m is None:
            self.m = (self.N0 + self.N1) // 4

    def prepare(self, n_trotters=None):
        if n_trotters is not None:
            self.m = n_trotters
        _fn('orc_bga_prepare', self.dtype)(self._o, int(self.m))

    def randomize_spin(self):
        _fn('orc_bga_randomize', self.dtype)(self._o)

    def set_qset(self, q0, q1):
        q0 = _arr(q0, np.int8); q1 = _arr(q1, np.int8)
        if q0.shape[0] != self.m:
            self.prepare(q0.shape[0])
        _fn('orc_bga_set_q', self.dtype)(self._o, _p(q0), _p(q1))

    def get_q(self):
        q0 = np.empty((self.m, self.N0), np.int8); q1 = np.empty((self.m, self.N1), np.int8)
        _fn('orc_bga_get_q', self.dtype)(self._o, _p(q0), _p(q1))
        return q0, q1

    def anneal_one_step(self, G, beta):
        r = _creal(self.dtype)
        _fn('orc_bga_anneal_one_step', self.dtype)(self._o, r(G), r(beta))

    def get_E(self):
        E = np.empty(self.m, self.dtype)
        _fn('orc_bga_calculate_E', self.dtype)(self._o, _p(E))
        return E

    def get_system_E(self, G, beta):
        r = _creal(self.dtype)
        return self.dtype(_fn('orc_bga_system_E', self.dtype)(self._o, r(G), r(beta)))

    def stats(self):
        """(accepted flips, accept tests that sat within a rounding error of the threshold)"""
        a = C.c_longlong(0); b = C.c_longlong(0)
        _fn('orc_bga_stats', self.dtype)(self._o, C.byref(a), C.byref(b))
        return a.value, b.value


# ---------------------------------------------------------------- brute force
def dense_graph_bf_search(W, optimize=0, dtype=np.float64, tile_size=1024, x_begin=0, x_end=None):
    """Restated CPUDenseGraphBFSearcher (single worker): returns (E, sorted packed x list)."""
    W = _arr(W, dtype); N = W.shape[0]
    x_end = (1 << N) if x_end is None else x_end
    tile_size = min(tile_size, 1 << N)
    Emin = np.array([FLT_MAX], dtype); sols = np.empty(max(tile_size, 1), np.uint64); n = C.c_int(0)
    x = x_begin
    while x < x_end:
        xe = min(x + tile_size, x_end)
        _fn('orc_dg_bf_range', dtype)(_p(W), N, int(optimize), C.c_uint64(x), C.c_uint64(xe), tile_size,
                                      _p(Emin), _p(sols), C.byref(n))
        x = xe
    xs = np.sort(sols[:n.value])
    E = -Emin[0] if optimize == 1 else Emin[0]
    return np.dtype(dtype).type(E), xs


def bipartite_graph_bf_search(b0, b1, W, optimize=0, dtype=np.float64, tile_size_0=1024, tile_size_1=1024):
    """Restated CPUBipartiteGraphBFSearcher (single worker): returns (E, list of (x0, x1) packed pairs)."""
    b0 = _arr(b0, dtype); b1 = _arr(b1, dtype); W = _arr(W, dtype)
    N1, N0 = W.shape
    t0 = min(tile_size_0, 1 << N0); t1 = min(tile_size_1, 1 << N1)
    cap = N0 + N1   # CPUBipartiteGraphBatchSearch.cpp:44 (maxNSolutions = rows + cols)
    Emin = np.array([FLT_MAX], dtype); sols = np.empty(2 * max(cap, 1), np.uint64); n = C.c_int(0)
    for x0 in range(0, 1 << N0, t0):
        for x1 in range(0, 1 << N1, t1):
            _fn('orc_bg_bf_range', dtype)(_p(b0), _p(b1), _p(W), N0, N1, int(optimize),
                                          C.c_uint64(x0), C.c_uint64(min(x0 + t0, 1 << N0)),
                                          C.c_uint64(x1), C.c_uint64(min(x1 + t1, 1 << N1)), cap,
                                          _p(Emin), _p(sols), C.byref(n))
    pairs = [(int(sols[2 * i]), int(sols[2 * i + 1])) for i in range(n.value)]
    E = -Emin[0] if optimize == 1 else Emin[0]
    return np.dtype(dtype).type(E), pairs


def unpack_bits(x, N):
    """Common.cpp:86-93: bit `pos` of the vector is (x >> (N-1-pos)) & 1 (MSB first)."""
    return np.array([(int(x) >> (N - 1 - pos)) & 1 for pos in range(N)], np.int8)
