"""Generate tests/golden/*.npz from the reference's own pure-Python solvers (sqaod.py).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The outputs are committed; nothing under tests/ reads /root/reference at test time.

The reference's top-level package imports its compiled CPU extension (sqaodpy/sqaod/__init__.py:32),
which cannot be built here (Eigen absent), so a stub `sqaod` package pointing at the reference tree is
registered first; sqaod.common and sqaod.py then import unmodified.
"""
import os
import sys
import types
import warnings
import numpy as np

warnings.simplefilter('ignore')
REF = '/root/reference/sqaodpy/sqaod'
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    pkg = types.ModuleType('sqaod')
    pkg.__path__ = [REF]
    sys.modules['sqaod'] = pkg
    import sqaod.common as c
    from sqaod.common.preference import algorithm, minimize, maximize
    pkg.algorithm, pkg.minimize, pkg.maximize = algorithm, minimize, maximize
    for k in dir(c):
        if not k.startswith('_'):
            setattr(pkg, k, getattr(c, k))
    import sqaod.py as py
    return pkg, py


def quantize(W):  # sqaodpy/tests/example_problems.py:16-18
    return np.rint(W * 16384) / 16384.


def all_bits(N):
    return np.array([[(v >> (N - 1 - p)) & 1 for p in range(N)] for v in range(1 << N)], np.int8)


def main():
    sq, py = import_reference()
    f = py.formulas
    rng_state = 20261017
    np.random.seed(rng_state)

    # ---------------- dense graph
    dense = {}
    W8 = np.full((8, 8), 4.0); np.fill_diagonal(W8, -32.0)      # example_problems.py:4-14
    Wr16 = quantize(sq.generate_random_symmetric_W(16))           # example_problems.py:20-22
    Wr12 = quantize(sq.generate_random_symmetric_W(12))
    for name, W in (('W8', W8), ('Wr16', Wr16), ('Wr12', Wr12)):
        N = W.shape[0]
        h, J, c = f.dense_graph_calculate_hamiltonian(W.copy())
        dense[name] = W
        dense[name + '_h'], dense[name + '_J'], dense[name + '_c'] = h, J, np.float64(c)
        x = all_bits(N) if N <= 12 else np.random.randint(0, 2, (64, N)).astype(np.int8)
        q = (2 * x - 1).astype(np.int8)
        dense[name + '_x'] = x
        dense[name + '_E_x'] = f.dense_graph_batch_calculate_E(W, x)
        dense[name + '_E_q'] = f.dense_graph_batch_calculate_E_from_spin(h, J, c, q)
        dense[name + '_E_x0'] = np.float64(f.dense_graph_calculate_E(W, x[3]))
        dense[name + '_E_q0'] = np.float64(f.dense_graph_calculate_E_from_spin(h, J, c, q[3]))
    for name, W in (('W8', W8), ('Wr12', Wr12)):
        for opt, tag in ((sq.minimize, 'min'), (sq.maximize, 'max')):
            s = py.dense_graph_bf_searcher(W, opt, tile_size=1 << W.shape[0])  # py search_range() only terminates when one tile covers the range
            s.search()
            dense['%s_bf_%s_E' % (name, tag)] = np.asarray(s.get_E(), np.float64)
            dense['%s_bf_%s_x' % (name, tag)] = np.asarray(s.get_x(), np.int8)
    # get_system_E (py/dense_graph_annealer.py:263-279), n_trotters = 6, random spins
    for opt, tag in ((sq.minimize, 'min'), (sq.maximize, 'max')):
        ann = py.dense_graph_annealer(Wr16, opt, n_trotters=6)
        ann.prepare()
        qs = (2 * np.random.randint(0, 2, (6, 16)) - 1).astype(np.int8)
        ann.set_qset(qs)
        dense['Wr16_sys_%s_q' % tag] = qs
        dense['Wr16_sys_%s_E' % tag] = np.asarray(ann.get_E(), np.float64)
        dense['Wr16_sys_%s_sysE' % tag] = np.float64(ann.get_system_E(0.7, 1.0 / 0.03))
    np.savez(os.path.join(OUT, 'dense_graph.npz'), **dense)

    # ---------------- bipartite graph
    bip = {}
    N0, N1 = 5, 4
    b0 = quantize(np.random.random(N0) - 0.5)                     # example_problems.py:25-29
    b1 = quantize(np.random.random(N1) - 0.5)
    W = quantize(np.random.random((N1, N0)) - 0.5)
    h0, h1, J, c = f.bipartite_graph_calculate_hamiltonian(b0, b1, W)
    bip.update(b0=b0, b1=b1, W=W, h0=h0, h1=h1, J=J, c=np.float64(c))
    x0 = all_bits(N0); x1 = all_bits(N1)
    bip['x0'], bip['x1'] = x0, x1
    # the py checker insists on len(x0) == len(x1) even for the 2-D form (common/checkers.py:139-142)
    bip['x0_2d'] = x0[7:7 + len(x1)]
    bip['E_2d'] = f.bipartite_graph_batch_calculate_E_2d(b0, b1, W, bip['x0_2d'], x1)
    idx0 = np.random.randint(0, 1 << N0, 24); idx1 = np.random.randint(0, 1 << N1, 24)
    bx0, bx1 = x0[idx0], x1[idx1]
    bip['bx0'], bip['bx1'] = bx0, bx1
    bip['E_x'] = f.bipartite_graph_batch_calculate_E(b0, b1, W, bx0, bx1)
    bip['E_q'] = f.bipartite_graph_batch_calculate_E_from_spin(h0, h1, J, c, 2 * bx0 - 1, 2 * bx1 - 1)
    bip['E_x0'] = np.float64(f.bipartite_graph_calculate_E(b0, b1, W, bx0[0], bx1[0]))
    for opt, tag in ((sq.minimize, 'min'), (sq.maximize, 'max')):
        s = py.bipartite_graph_bf_searcher(b0, b1, W, opt)
        s.search()
        bip['bf_%s_E' % tag] = np.asarray(s.get_E(), np.float64)
        xs = s.get_x()
        bip['bf_%s_x0' % tag] = np.asarray([p[0] for p in xs], np.int8)
        bip['bf_%s_x1' % tag] = np.asarray([p[1] for p in xs], np.int8)
        ann = py.bipartite_graph_annealer(b0, b1, W, opt, n_trotters=6)
        ann.prepare()
        q0 = (2 * np.random.randint(0, 2, (6, N0)) - 1).astype(np.int8)
        q1 = (2 * np.random.randint(0, 2, (6, N1)) - 1).astype(np.int8)
        ann.set_qset([(q0[i], q1[i]) for i in range(6)])
        bip['sys_%s_q0' % tag], bip['sys_%s_q1' % tag] = q0, q1
        bip['sys_%s_E' % tag] = np.asarray(ann.get_E(), np.float64)
        bip['sys_%s_sysE' % tag] = np.float64(ann.get_system_E(0.7, 1.0 / 0.03))
    np.savez(os.path.join(OUT, 'bipartite_graph.npz'), **bip)
    print('written:', sorted(os.listdir(OUT)))


if __name__ == '__main__':
    main()
