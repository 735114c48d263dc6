"""Stateless formulas on the device; same functions as sqaodpy/sqaod/cuda/formulas.py (each one pyglue/formulas.inc call)."""
import ctypes as C
import numpy as np
from . import _lib, common
from . import device as _device

L = _lib.lib
ptr = _lib.ptr
_objs = {}


def _get(kind, dtype):
    dt = _lib.dtype_code(dtype)
    key = (kind, dt)
    if key not in _objs:
        h = C.c_void_p()
        _lib.check(getattr(L, 'sqb_%s_formulas_new' % kind)(C.byref(h), dt))
        _lib.check(getattr(L, 'sqb_%s_formulas_assign_device' % kind)(h, _device.active_device()._cobj, dt))
        _objs[key] = h
    return _objs[key], dt


def _bits(x):
    return np.ascontiguousarray(np.atleast_2d(np.asarray(x)), np.int8)


def dense_graph_calculate_E(W, x, dtype):
    return dense_graph_batch_calculate_E(W, x, dtype)[0]


def dense_graph_batch_calculate_E(W, x, dtype):
    f, dt = _get('dg', dtype)
    W = np.ascontiguousarray(common.symmetrize(common.fix_type(W, dtype))); x = _bits(x)
    E = np.empty(x.shape[0], dtype)
    _lib.check(L.sqb_dg_formulas_calculate_E(f, ptr(E), ptr(W), W.shape[0], W.shape[1], ptr(x), x.shape[0], dt))
    return E


def dense_graph_calculate_hamiltonian(W, dtype):
    f, dt = _get('dg', dtype)
    W = np.ascontiguousarray(common.symmetrize(common.fix_type(W, dtype))); N = W.shape[0]
    h = np.empty(N, dtype); J = np.empty((N, N), dtype); c = np.empty(1, dtype)
    _lib.check(L.sqb_dg_formulas_calculate_hamiltonian(f, ptr(h), ptr(J), N, ptr(c), ptr(W), N, N, dt))
    return h, J, c[0]


def dense_graph_calculate_E_from_spin(h, J, c, q, dtype):
    return dense_graph_batch_calculate_E_from_spin(h, J, c, q, dtype)[0]


def dense_graph_batch_calculate_E_from_spin(h, J, c, q, dtype):
    f, dt = _get('dg', dtype)
    h, J = common.fix_type([h, J], dtype); J = np.ascontiguousarray(common.symmetrize(J)); q = _bits(q)
    E = np.empty(q.shape[0], dtype)
    _lib.check(L.sqb_dg_formulas_calculate_E_from_spin(f, ptr(E), ptr(h), ptr(J), J.shape[0], J.shape[1],
                                                       C.c_double(float(c)), ptr(q), q.shape[0], dt))
    return E


def bipartite_graph_calculate_E(b0, b1, W, x0, x1, dtype):
    return bipartite_graph_batch_calculate_E(b0, b1, W, x0, x1, dtype)[0]


def bipartite_graph_batch_calculate_E(b0, b1, W, x0, x1, dtype):
    f, dt = _get('bg', dtype)
    b0, b1, W = common.fix_type([b0, b1, W], dtype); x0 = _bits(x0); x1 = _bits(x1)
    E = np.empty(x0.shape[0], dtype)
    _lib.check(L.sqb_bg_formulas_calculate_E(f, ptr(E), ptr(b0), ptr(b1), ptr(W), b0.shape[0], b1.shape[0], W.shape[1],
                                             ptr(x0), ptr(x1), x0.shape[0], dt))
    return E


def bipartite_graph_batch_calculate_E_2d(b0, b1, W, x0, x1, dtype):
    f, dt = _get('bg', dtype)
    b0, b1, W = common.fix_type([b0, b1, W], dtype); x0 = _bits(x0); x1 = _bits(x1)
    E = np.empty((x1.shape[0], x0.shape[0]), dtype)
    _lib.check(L.sqb_bg_formulas_calculate_E_2d(f, ptr(E), ptr(b0), ptr(b1), ptr(W), b0.shape[0], b1.shape[0], W.shape[1],
                                                ptr(x0), x0.shape[0], ptr(x1), x1.shape[0], dt))
    return E


def bipartite_graph_calculate_hamiltonian(b0, b1, W, dtype):
    f, dt = _get('bg', dtype)
    b0, b1, W = common.fix_type([b0, b1, W], dtype)
    N0, N1 = b0.shape[0], b1.shape[0]
    h0 = np.empty(N0, dtype); h1 = np.empty(N1, dtype); J = np.empty((N1, N0), dtype); c = np.empty(1, dtype)
    _lib.check(L.sqb_bg_formulas_calculate_hamiltonian(f, ptr(h0), ptr(h1), ptr(J), N0, ptr(c), ptr(b0), ptr(b1), ptr(W),
                                                       N0, N1, N0, dt))
    return h0, h1, J, c[0]


def bipartite_graph_calculate_E_from_spin(h0, h1, J, c, q0, q1, dtype):
    return bipartite_graph_batch_calculate_E_from_spin(h0, h1, J, c, q0, q1, dtype)[0]


def bipartite_graph_batch_calculate_E_from_spin(h0, h1, J, c, q0, q1, dtype):
    f, dt = _get('bg', dtype)
    h0, h1, J = common.fix_type([h0, h1, J], dtype); q0 = _bits(q0); q1 = _bits(q1)
    E = np.empty(q0.shape[0], dtype)
    _lib.check(L.sqb_bg_formulas_calculate_E_from_spin(f, ptr(E), ptr(h0), ptr(h1), ptr(J), h0.shape[0], h1.shape[0], J.shape[1],
                                                       C.c_double(float(c)), ptr(q0), ptr(q1), q0.shape[0], dt))
    return E
