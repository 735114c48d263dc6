"""cuda_bg_bf_searcher (method table: sqaodc/pyglue/bf_searcher.inc:478-494, bipartite-graph variant)"""
import ctypes as C
import numpy as np
from ._glue import L, check, ptr, dt, h, new_handle, stride, simple
from . import _glue

_P = 'bg_bf_searcher'


def new(dtype):
    return new_handle(L.sqb_bg_bf_searcher_new, dtype)


def delete(obj, dtype):
    check(L.sqb_bg_bf_searcher_delete(h(obj), dt(dtype)))


def assign_device(obj, dev, dtype):
    check(L.sqb_bg_bf_searcher_assign_device(h(obj), h(dev), dt(dtype)))


def set_qubo(obj, b0, b1, W, opt, dtype):
    check(L.sqb_bg_bf_searcher_set_qubo(h(obj), ptr(b0), ptr(b1), ptr(W), b0.shape[0], b1.shape[0], stride(W), int(opt), dt(dtype)))


def get_problem_size(obj, dtype):
    n0 = C.c_int(0); n1 = C.c_int(0)
    check(L.sqb_bg_bf_searcher_get_problem_size(h(obj), C.byref(n0), C.byref(n1), dt(dtype)))
    return n0.value, n1.value


def set_preferences(obj, prefs, dtype):
    _glue.set_preferences(_P, obj, prefs, dtype)


def get_preferences(obj, dtype):
    return _glue.get_preferences(_P, obj, dtype)


def _n(obj, dtype):
    n = C.c_int(0)
    check(L.sqb_bg_bf_searcher_get_num_solutions(h(obj), C.byref(n), dt(dtype)))
    return n.value


def get_x(obj, dtype):
    n = _n(obj, dtype); N0, N1 = get_problem_size(obj, dtype)
    x0 = np.empty((max(n, 1), N0), np.int8); x1 = np.empty((max(n, 1), N1), np.int8)
    check(L.sqb_bg_bf_searcher_get_x(h(obj), ptr(x0), ptr(x1), n, dt(dtype)))
    return [(x0[i], x1[i]) for i in range(n)]


def get_E(obj, dtype):
    n = max(_n(obj, dtype), 1)
    E = np.empty(n, dtype)
    check(L.sqb_bg_bf_searcher_get_E(h(obj), ptr(E), n, dt(dtype)))
    return E


prepare = simple(_P, 'prepare')
calculate_E = simple(_P, 'calculate_E')
make_solution = simple(_P, 'make_solution')
search = simple(_P, 'search')


def search_range(obj, dtype):
    done = C.c_int(0); x0 = C.c_ulonglong(0); x1 = C.c_ulonglong(0)
    check(L.sqb_bg_bf_searcher_search_range(h(obj), C.byref(done), C.byref(x0), C.byref(x1), dt(dtype)))
    return bool(done.value), x0.value, x1.value
