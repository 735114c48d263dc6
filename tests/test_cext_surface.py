"""The cext-compatible modules (sqaod_b200/cext) expose the method tables of the reference's compiled extension modules
and work when driven exactly the way the reference's *_base.py classes drive `self._cext`."""
import numpy as np
import pytest
from conftest import quantized_symmetric_W, quantized_bipartite

# method tables of the reference glue: annealer.inc:884-910, bf_searcher.inc:478-494, formulas.inc:586-606, cuda_device.cpp:65-72
ANNEALER = ['new', 'delete', 'assign_device', 'seed', 'set_qubo', 'set_hamiltonian', 'get_problem_size', 'set_preferences',
            'get_preferences', 'get_E', 'get_x', 'get_hamiltonian', 'get_q', 'set_q', 'set_qset', 'randomize_spin', 'calculate_E',
            'prepare', 'make_solution', 'get_system_E', 'anneal_one_step']
SEARCHER = ['new', 'delete', 'assign_device', 'set_qubo', 'get_problem_size', 'set_preferences', 'get_preferences', 'get_x', 'get_E',
            'prepare', 'calculate_E', 'make_solution', 'search_range', 'search']
FORMULAS = ['dg_formulas_new', 'dg_formulas_delete', 'dg_formulas_assign_device', 'dense_graph_calculate_E',
            'dense_graph_batch_calculate_E', 'dense_graph_calculate_hamiltonian', 'dense_graph_calculate_E_from_spin',
            'dense_graph_batch_calculate_E_from_spin', 'bg_formulas_new', 'bg_formulas_delete', 'bg_formulas_assign_device',
            'bipartite_graph_calculate_E', 'bipartite_graph_batch_calculate_E', 'bipartite_graph_batch_calculate_E_2d',
            'bipartite_graph_calculate_hamiltonian', 'bipartite_graph_calculate_E_from_spin',
            'bipartite_graph_batch_calculate_E_from_spin']


def test_method_tables():
    from sqaod_b200 import cext
    for mod in (cext.cuda_dg_annealer, cext.cuda_bg_annealer):
        assert [n for n in ANNEALER if not callable(getattr(mod, n, None))] == []
    for mod in (cext.cuda_dg_bf_searcher, cext.cuda_bg_bf_searcher):
        assert [n for n in SEARCHER if not callable(getattr(mod, n, None))] == []
    assert [n for n in FORMULAS if not callable(getattr(cext.cuda_formulas, n, None))] == []
    assert [n for n in ('new', 'delete', 'initialize', 'finalize') if not callable(getattr(cext.cuda_device, n, None))] == []
    # host-side calls work without a GPU and carry the reference's conventions (uint64 handle, dtype last)
    h = cext.cuda_dg_annealer.new(np.float64)
    assert isinstance(h, np.uint64)
    cext.cuda_dg_annealer.set_preferences(h, {'n_trotters': 6, 'algorithm': 'sa_default'}, np.float64)
    p = cext.cuda_dg_annealer.get_preferences(h, np.float64)
    assert p == {'algorithm': 'sa_naive', 'n_trotters': 6, 'precision': 'double', 'device': 'cuda'}
    with pytest.raises(RuntimeError):
        cext.cuda_dg_annealer.prepare(h, np.float64)
    with pytest.raises(RuntimeError):
        cext.cuda_dg_annealer.new(np.int32)
    cext.cuda_dg_annealer.delete(h, np.float64)


class RefStyleDenseGraphAnnealer(object):
    """drives a cext module the way sqaodpy/sqaod/common/dense_graph_annealer_base.py:7-101 does"""

    def __init__(self, cext, cobj, dtype, W, optimize, prefdict):
        self._cext, self._cobj, self.dtype = cext, cobj, dtype
        self._cext.set_qubo(self._cobj, np.ascontiguousarray(W, dtype), optimize, self.dtype)
        self._cext.set_preferences(self._cobj, prefdict, self.dtype)

    def run(self, seed, schedule, beta):
        c, o, d = self._cext, self._cobj, self.dtype
        c.seed(o, seed, d); c.prepare(o, d); c.randomize_spin(o, d)
        for G in schedule:
            c.anneal_one_step(o, d(G), d(beta), d)
        c.make_solution(o, d)
        N = c.get_problem_size(o, d)
        h = np.empty(N, d); J = np.empty((N, N), d); cc = np.empty(1, d)
        c.get_hamiltonian(o, h, J, cc, d)
        return c.get_E(o, d), c.get_x(o, d), c.get_q(o, d), c.get_system_E(o, d(schedule[-1]), d(beta), d), (h, J, cc[0])


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_reference_style_driver_matches_package(dtype):
    import sqaod_b200 as sq
    from sqaod_b200 import cext
    dev = cext.cuda_device.new()
    cext.cuda_device.initialize(dev, 0)
    W = quantized_symmetric_W(48, 9, dtype)
    obj = cext.cuda_dg_annealer.new(dtype)
    cext.cuda_dg_annealer.assign_device(obj, dev, dtype)
    ref = RefStyleDenseGraphAnnealer(cext.cuda_dg_annealer, obj, dtype, W, 0, {'n_trotters': 12})
    sched = [2.0, 1.0, 0.5, 0.1]
    E, x, q, sysE, (h, J, c) = ref.run(7, sched, 10.0)
    ann = sq.dense_graph_annealer(W, sq.minimize, dtype, n_trotters=12)
    ann.seed(7); ann.prepare(); ann.randomize_spin()
    for G in sched:
        ann.anneal_one_step(G, 10.0)
    ann.make_solution()
    assert np.array_equal(E, ann.get_E()) and np.array_equal(np.stack(q), np.stack(ann.get_q()))
    assert np.array_equal(np.stack(x), (np.stack(q) + 1) // 2)
    assert sysE == ann.get_system_E(sched[-1], 10.0)
    h2, J2, c2 = ann.get_hamiltonian()
    assert np.array_equal(h, h2) and np.array_equal(J, J2) and c == c2
    cext.cuda_dg_annealer.delete(obj, dtype)
    # searcher + formulas through the cext surface
    s = cext.cuda_dg_bf_searcher.new(dtype)
    cext.cuda_dg_bf_searcher.assign_device(s, dev, dtype)
    W8 = np.full((8, 8), 4.0, dtype); np.fill_diagonal(W8, -32.0)
    cext.cuda_dg_bf_searcher.set_qubo(s, W8, 0, dtype)
    cext.cuda_dg_bf_searcher.prepare(s, dtype)
    while not cext.cuda_dg_bf_searcher.search_range(s, dtype)[0]:
        pass
    cext.cuda_dg_bf_searcher.make_solution(s, dtype)
    assert len(cext.cuda_dg_bf_searcher.get_x(s, dtype)) == 126 and np.all(cext.cuda_dg_bf_searcher.get_E(s, dtype) == -80)
    cext.cuda_dg_bf_searcher.delete(s, dtype)
    f = cext.cuda_formulas.dg_formulas_new()
    cext.cuda_formulas.dg_formulas_assign_device(f, dev)
    xs = np.array([[1, 0, 1, 1, 0, 0, 1, 0], [1] * 8], np.int8)
    Eb = np.empty(2, dtype)
    cext.cuda_formulas.dense_graph_batch_calculate_E(f, Eb, W8, xs, dtype)
    assert Eb[0] == 4 * 16 - 36 * 4 and Eb[1] == 4 * 64 - 36 * 8
    cext.cuda_formulas.dg_formulas_delete(f)
    b0, b1, Wb = quantized_bipartite(6, 5, 3, dtype)
    bg = cext.cuda_bg_annealer.new(dtype)
    cext.cuda_bg_annealer.assign_device(bg, dev, dtype)
    cext.cuda_bg_annealer.set_qubo(bg, b0, b1, Wb, 0, dtype)
    cext.cuda_bg_annealer.set_preferences(bg, {'n_trotters': 4}, dtype)
    cext.cuda_bg_annealer.seed(bg, 3, dtype); cext.cuda_bg_annealer.prepare(bg, dtype); cext.cuda_bg_annealer.randomize_spin(bg, dtype)
    cext.cuda_bg_annealer.anneal_one_step(bg, dtype(1.0), dtype(5.0), dtype)
    qp = cext.cuda_bg_annealer.get_q(bg, dtype)
    assert len(qp) == 4 and qp[0][0].shape == (6,) and qp[0][1].shape == (5,)
    cext.cuda_bg_annealer.delete(bg, dtype)
    cext.cuda_device.finalize(dev); cext.cuda_device.delete(dev)
