"""cuda_dg_annealer (method table: sqaodc/pyglue/annealer.inc:884-910, dense-graph variant)"""
import ctypes as C
import numpy as np
from ._glue import L, check, ptr, dt, h, new_handle, stride, simple, vec
from . import _glue

_P = 'dg_annealer'


def new(dtype):
    return new_handle(L.sqb_dg_annealer_new, dtype)


def delete(obj, dtype):
    check(L.sqb_dg_annealer_delete(h(obj), dt(dtype)))


def assign_device(obj, dev, dtype):
    check(L.sqb_dg_annealer_assign_device(h(obj), h(dev), dt(dtype)))


def seed(obj, seed, dtype):
    check(L.sqb_dg_annealer_seed(h(obj), C.c_ulonglong(int(seed)), dt(dtype)))


def set_qubo(obj, W, opt, dtype):
    check(L.sqb_dg_annealer_set_qubo(h(obj), ptr(W), W.shape[0], stride(W), int(opt), dt(dtype)))


def set_hamiltonian(obj, hvec, J, c, dtype):
    check(L.sqb_dg_annealer_set_hamiltonian(h(obj), ptr(hvec), ptr(J), J.shape[0], stride(J), C.c_double(float(c)), dt(dtype)))


def get_hamiltonian(obj, hvec, J, c, dtype):
    check(L.sqb_dg_annealer_get_hamiltonian(h(obj), ptr(hvec), ptr(J), stride(J), ptr(c), dt(dtype)))


def get_problem_size(obj, dtype):
    n = C.c_int(0)
    check(L.sqb_dg_annealer_get_problem_size(h(obj), C.byref(n), dt(dtype)))
    return n.value


def set_preferences(obj, prefs, dtype):
    _glue.set_preferences(_P, obj, prefs, dtype)


def get_preferences(obj, dtype):
    return _glue.get_preferences(_P, obj, dtype)


def _rows(obj, dtype):
    m = C.c_int(0); r = C.c_int(1)
    check(L.sqb_dg_annealer_get_num_trotters(h(obj), C.byref(m), dt(dtype)))
    check(L.sqb_dg_annealer_get_num_replicas(h(obj), C.byref(r), dt(dtype)))
    return m.value * r.value


def get_E(obj, dtype):
    m = _rows(obj, dtype)
    E = np.empty(m, dtype)
    check(L.sqb_dg_annealer_get_E(h(obj), ptr(E), m, dt(dtype)))
    return E


def _bits(fn, obj, dtype):
    m, N = _rows(obj, dtype), get_problem_size(obj, dtype)
    out = np.empty((m, N), np.int8)
    check(fn(h(obj), ptr(out), dt(dtype)))
    return [out[i] for i in range(m)]


def get_x(obj, dtype):
    return _bits(L.sqb_dg_annealer_get_x, obj, dtype)


def get_q(obj, dtype):
    return _bits(L.sqb_dg_annealer_get_q, obj, dtype)


def set_q(obj, q, dtype):
    q = vec(q)
    check(L.sqb_dg_annealer_set_q(h(obj), ptr(q), q.shape[0], dt(dtype)))


def set_qset(obj, qlist, dtype):
    q = np.ascontiguousarray(np.stack(qlist), np.int8)
    check(L.sqb_dg_annealer_set_qset(h(obj), ptr(q), q.shape[0], q.shape[1], dt(dtype)))


randomize_spin = simple(_P, 'randomize_spin')
calculate_E = simple(_P, 'calculate_E')
prepare = simple(_P, 'prepare')
make_solution = simple(_P, 'make_solution')


def get_system_E(obj, G, beta, dtype):
    E = C.c_double(0)
    check(L.sqb_dg_annealer_get_system_E(h(obj), C.c_double(float(G)), C.c_double(float(beta)), C.byref(E), dt(dtype)))
    return dtype(E.value)


def anneal_one_step(obj, G, beta, dtype):
    check(L.sqb_dg_annealer_anneal_one_step(h(obj), C.c_double(float(G)), C.c_double(float(beta)), dt(dtype)))
