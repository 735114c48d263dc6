"""cuda_formulas (method table: sqaodc/pyglue/formulas.inc:586-606).  The handle of a formulas object is a pair
{float instance, double instance} in the reference (formulas.inc:5-36); here a small Python object holds the two C-ABI
handles and is passed around as `obj`.  Outputs are caller-allocated numpy arrays passed second."""
import ctypes as C
import numpy as np
from ._glue import L, check, ptr, dt, h

_KEEP = {}


class _Pair(object):
    def __init__(self, kind):
        self.kind = kind
        self.h = {}
        for code in (0, 1):
            p = C.c_void_p()
            check(getattr(L, 'sqb_%s_formulas_new' % kind)(C.byref(p), code))
            self.h[code] = p


def _new(kind):
    pair = _Pair(kind)
    key = np.uint64(id(pair))
    _KEEP[int(key)] = pair
    return key


def _get(obj, dtype):
    pair = _KEEP[int(obj)]
    return pair.h[dt(dtype)], dt(dtype)


def _delete(obj):
    pair = _KEEP.pop(int(obj))
    for code, p in pair.h.items():
        check(getattr(L, 'sqb_%s_formulas_delete' % pair.kind)(p, code))


def _assign(obj, dev):
    pair = _KEEP[int(obj)]
    for code, p in pair.h.items():
        check(getattr(L, 'sqb_%s_formulas_assign_device' % pair.kind)(p, h(dev), code))


def dg_formulas_new():
    return _new('dg')


def dg_formulas_delete(obj):
    _delete(obj)


def dg_formulas_assign_device(obj, dev):
    _assign(obj, dev)


def bg_formulas_new():
    return _new('bg')


def bg_formulas_delete(obj):
    _delete(obj)


def bg_formulas_assign_device(obj, dev):
    _assign(obj, dev)


def _b(x):
    return np.ascontiguousarray(np.atleast_2d(x), np.int8)


def dense_graph_calculate_E(obj, E, W, x, dtype):
    f, d = _get(obj, dtype); x = _b(x)
    check(L.sqb_dg_formulas_calculate_E(f, ptr(E), ptr(W), W.shape[0], W.strides[0] // W.itemsize, ptr(x), 1, d))


def dense_graph_batch_calculate_E(obj, E, W, x, dtype):
    f, d = _get(obj, dtype); x = _b(x)
    check(L.sqb_dg_formulas_calculate_E(f, ptr(E), ptr(W), W.shape[0], W.strides[0] // W.itemsize, ptr(x), x.shape[0], d))


def dense_graph_calculate_hamiltonian(obj, hvec, J, c, W, dtype):
    f, d = _get(obj, dtype)
    check(L.sqb_dg_formulas_calculate_hamiltonian(f, ptr(hvec), ptr(J), J.strides[0] // J.itemsize, ptr(c), ptr(W), W.shape[0],
                                                  W.strides[0] // W.itemsize, d))


def dense_graph_calculate_E_from_spin(obj, E, hvec, J, c, q, dtype):
    dense_graph_batch_calculate_E_from_spin(obj, E, hvec, J, c, q, dtype)


def dense_graph_batch_calculate_E_from_spin(obj, E, hvec, J, c, q, dtype):
    f, d = _get(obj, dtype); q = _b(q)
    check(L.sqb_dg_formulas_calculate_E_from_spin(f, ptr(E), ptr(hvec), ptr(J), J.shape[0], J.strides[0] // J.itemsize,
                                                  C.c_double(float(c)), ptr(q), q.shape[0], d))


def bipartite_graph_calculate_E(obj, E, b0, b1, W, x0, x1, dtype):
    bipartite_graph_batch_calculate_E(obj, E, b0, b1, W, x0, x1, dtype)


def bipartite_graph_batch_calculate_E(obj, E, b0, b1, W, x0, x1, dtype):
    f, d = _get(obj, dtype); x0 = _b(x0); x1 = _b(x1)
    check(L.sqb_bg_formulas_calculate_E(f, ptr(E), ptr(b0), ptr(b1), ptr(W), b0.shape[0], b1.shape[0], W.strides[0] // W.itemsize,
                                        ptr(x0), ptr(x1), x0.shape[0], d))


def bipartite_graph_batch_calculate_E_2d(obj, E, b0, b1, W, x0, x1, dtype):
    f, d = _get(obj, dtype); x0 = _b(x0); x1 = _b(x1)
    check(L.sqb_bg_formulas_calculate_E_2d(f, ptr(E), ptr(b0), ptr(b1), ptr(W), b0.shape[0], b1.shape[0], W.strides[0] // W.itemsize,
                                           ptr(x0), x0.shape[0], ptr(x1), x1.shape[0], d))


def bipartite_graph_calculate_hamiltonian(obj, h0, h1, J, c, b0, b1, W, dtype):
    f, d = _get(obj, dtype)
    check(L.sqb_bg_formulas_calculate_hamiltonian(f, ptr(h0), ptr(h1), ptr(J), J.strides[0] // J.itemsize, ptr(c), ptr(b0), ptr(b1),
                                                  ptr(W), b0.shape[0], b1.shape[0], W.strides[0] // W.itemsize, d))


def bipartite_graph_calculate_E_from_spin(obj, E, h0, h1, J, c, q0, q1, dtype):
    bipartite_graph_batch_calculate_E_from_spin(obj, E, h0, h1, J, c, q0, q1, dtype)


def bipartite_graph_batch_calculate_E_from_spin(obj, E, h0, h1, J, c, q0, q1, dtype):
    f, d = _get(obj, dtype); q0 = _b(q0); q1 = _b(q1)
    check(L.sqb_bg_formulas_calculate_E_from_spin(f, ptr(E), ptr(h0), ptr(h1), ptr(J), h0.shape[0], h1.shape[0],
                                                  J.strides[0] // J.itemsize, C.c_double(float(c)), ptr(q0), ptr(q1), q0.shape[0], d))
