"""cuda_dg_bf_searcher (method table: sqaodc/pyglue/bf_searcher.inc:478-494, dense-graph variant)"""
import ctypes as C
import numpy as np
from ._glue import L, check, ptr, dt, h, new_handle, stride, simple
from . import _glue

_P = 'dg_bf_searcher'


def new(dtype):
    return new_handle(L.sqb_dg_bf_searcher_new, dtype)


def delete(obj, dtype):
    check(L.sqb_dg_bf_searcher_delete(h(obj), dt(dtype)))


def assign_device(obj, dev, dtype):
    check(L.sqb_dg_bf_searcher_assign_device(h(obj), h(dev), dt(dtype)))


def set_qubo(obj, W, opt, dtype):
    check(L.sqb_dg_bf_searcher_set_qubo(h(obj), ptr(W), W.shape[0], stride(W), int(opt), dt(dtype)))


def get_problem_size(obj, dtype):
    n = C.c_int(0)
    check(L.sqb_dg_bf_searcher_get_problem_size(h(obj), C.byref(n), dt(dtype)))
    return n.value


def set_preferences(obj, prefs, dtype):
    _glue.set_preferences(_P, obj, prefs, dtype)


def get_preferences(obj, dtype):
    return _glue.get_preferences(_P, obj, dtype)


def _n(obj, dtype):
    n = C.c_int(0)
    check(L.sqb_dg_bf_searcher_get_num_solutions(h(obj), C.byref(n), dt(dtype)))
    return n.value


def get_x(obj, dtype):
    n, N = _n(obj, dtype), get_problem_size(obj, dtype)
    x = np.empty((max(n, 1), N), np.int8)
    check(L.sqb_dg_bf_searcher_get_x(h(obj), ptr(x), n, dt(dtype)))
    return [x[i] for i in range(n)]


def get_E(obj, dtype):
    n = max(_n(obj, dtype), 1)
    E = np.empty(n, dtype)
    check(L.sqb_dg_bf_searcher_get_E(h(obj), ptr(E), n, dt(dtype)))
    return E


prepare = simple(_P, 'prepare')
calculate_E = simple(_P, 'calculate_E')
make_solution = simple(_P, 'make_solution')
search = simple(_P, 'search')


def search_range(obj, dtype):
    done = C.c_int(0); x = C.c_ulonglong(0)
    check(L.sqb_dg_bf_searcher_search_range(h(obj), C.byref(done), C.byref(x), dt(dtype)))
    return bool(done.value), x.value
