"""cuda_bg_annealer (method table: sqaodc/pyglue/annealer.inc:884-910, bipartite-graph variant)"""
import ctypes as C
import numpy as np
from ._glue import L, check, ptr, dt, h, new_handle, stride, simple, vec
from . import _glue

_P = 'bg_annealer'


def new(dtype):
    return new_handle(L.sqb_bg_annealer_new, dtype)


def delete(obj, dtype):
    check(L.sqb_bg_annealer_delete(h(obj), dt(dtype)))


def assign_device(obj, dev, dtype):
    check(L.sqb_bg_annealer_assign_device(h(obj), h(dev), dt(dtype)))


def seed(obj, seed, dtype):
    check(L.sqb_bg_annealer_seed(h(obj), C.c_ulonglong(int(seed)), dt(dtype)))


def set_qubo(obj, b0, b1, W, opt, dtype):
    check(L.sqb_bg_annealer_set_qubo(h(obj), ptr(b0), ptr(b1), ptr(W), b0.shape[0], b1.shape[0], stride(W), int(opt), dt(dtype)))


def set_hamiltonian(obj, h0, h1, J, c, dtype):
    check(L.sqb_bg_annealer_set_hamiltonian(h(obj), ptr(h0), ptr(h1), ptr(J), h0.shape[0], h1.shape[0], stride(J),
                                            C.c_double(float(c)), dt(dtype)))


def get_hamiltonian(obj, h0, h1, J, c, dtype):
    check(L.sqb_bg_annealer_get_hamiltonian(h(obj), ptr(h0), ptr(h1), ptr(J), stride(J), ptr(c), dt(dtype)))


def get_problem_size(obj, dtype):
    n0 = C.c_int(0); n1 = C.c_int(0)
    check(L.sqb_bg_annealer_get_problem_size(h(obj), C.byref(n0), C.byref(n1), dt(dtype)))
    return n0.value, n1.value


def set_preferences(obj, prefs, dtype):
    _glue.set_preferences(_P, obj, prefs, dtype)


def get_preferences(obj, dtype):
    return _glue.get_preferences(_P, obj, dtype)


def _m(obj, dtype):
    m = C.c_int(0)
    check(L.sqb_bg_annealer_get_num_trotters(h(obj), C.byref(m), dt(dtype)))
    return m.value


def get_E(obj, dtype):
    m = _m(obj, dtype)
    E = np.empty(m, dtype)
    check(L.sqb_bg_annealer_get_E(h(obj), ptr(E), m, dt(dtype)))
    return E


def _pairs(fn, obj, dtype):
    m = _m(obj, dtype); N0, N1 = get_problem_size(obj, dtype)
    a = np.empty((m, N0), np.int8); b = np.empty((m, N1), np.int8)
    check(fn(h(obj), ptr(a), ptr(b), dt(dtype)))
    return [(a[i], b[i]) for i in range(m)]


def get_x(obj, dtype):
    return _pairs(L.sqb_bg_annealer_get_x, obj, dtype)


def get_q(obj, dtype):
    return _pairs(L.sqb_bg_annealer_get_q, obj, dtype)


def set_q(obj, qpair, dtype):
    q0, q1 = vec(qpair[0]), vec(qpair[1])
    check(L.sqb_bg_annealer_set_q(h(obj), ptr(q0), ptr(q1), q0.shape[0], q1.shape[0], dt(dtype)))


def set_qset(obj, qpairs, dtype):
    q0 = np.ascontiguousarray(np.stack([p[0] for p in qpairs]), np.int8)
    q1 = np.ascontiguousarray(np.stack([p[1] for p in qpairs]), np.int8)
    check(L.sqb_bg_annealer_set_qset(h(obj), ptr(q0), ptr(q1), q0.shape[0], q0.shape[1], q1.shape[1], dt(dtype)))


randomize_spin = simple(_P, 'randomize_spin')
calculate_E = simple(_P, 'calculate_E')
prepare = simple(_P, 'prepare')
make_solution = simple(_P, 'make_solution')


def get_system_E(obj, G, beta, dtype):
    E = C.c_double(0)
    check(L.sqb_bg_annealer_get_system_E(h(obj), C.c_double(float(G)), C.c_double(float(beta)), C.byref(E), dt(dtype)))
    return dtype(E.value)


def anneal_one_step(obj, G, beta, dtype):
    check(L.sqb_bg_annealer_anneal_one_step(h(obj), C.c_double(float(G)), C.c_double(float(beta)), dt(dtype)))
