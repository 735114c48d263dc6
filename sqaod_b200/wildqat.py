"""wildqat-style front end over the B200 dense-graph annealer -- the counterpart of the reference's adapter
(sqaodpy/sqaod/wildqat/opt.py:5-67): an `opt` object carrying a QUBO and the wildqat schedule attributes, with `sa()` and
`sqa()` that drive `anneal_one_step` the way the reference adapter does.

    import sqaod_b200.wildqat as wq
    a = wq.opt()
    a.qubo = [[4, -4, -4], [0, 4, -4], [0, 0, 4]]
    x = a.sa()            # one bit vector
    xs = a.sqa()          # one bit vector per trotter; a.E holds the system energy per schedule point

If the `wildqat` package is importable its `opt` is the base class (as in the reference); otherwise a stand-in with the same
schedule attributes and defaults (wildqat 1.x: Ts=5, Tf=0.02, Gs=10, Gf=0.02, tro=8, ite=1000, R=0.95) is used."""
import numpy as np
from . import solvers

try:                                    # pragma: no cover - wildqat is not part of this image
    import wildqat as _wq
    _Base = _wq.opt
except Exception:
    class _Base(object):
        def __init__(self):
            self.Ts, self.Tf = 5.0, 0.02        # SA temperature schedule
            self.Gs, self.Gf = 10.0, 0.02       # SQA transverse-field schedule
            self.tro = 8                        # trotters
            self.ite = 1000                     # flips per schedule point
            self.R = 0.95                       # geometric factor
            self.qubo, self.J, self.E = [], [], []

        def qi(self):
            """wildqat converts the QUBO to Ising form here; the annealer does that itself (set_qubo)"""
            return None


def _symmetric(qubo, dtype):
    """wildqat QUBOs are usually upper triangular: x^T Q x is unchanged by symmetrising, which set_qubo requires"""
    Q = np.asarray(qubo, dtype=np.float64)
    return np.asarray((Q + Q.T) * 0.5, dtype)


class opt(_Base):
    def __init__(self, pkg=None, dtype=np.float32):
        _Base.__init__(self)
        self.dtype = dtype
        self.ann = (pkg or solvers).dense_graph_annealer(dtype=dtype)

    def sa(self):
        """Run SA with the provided QUBO (set `qubo` first); returns one bit vector.  opt.py:12-38."""
        self.E = []
        T = self.Ts
        if len(self.qubo):
            self.ann.set_qubo(_symmetric(self.qubo, self.dtype), solvers.minimize)
            self.ann.set_preferences(algorithm=solvers.algorithm.sa_naive, n_trotters=1)
        self.ann.prepare()
        self.ann.randomize_spin()
        N = self.ann.get_problem_size()
        n_iters_at_T = (self.ite + N - 1) // N
        while T > self.Tf:
            for _ in range(n_iters_at_T):
                self.ann.anneal_one_step(T, 1.)
                self.E.append(self.ann.get_system_E(0., 0.))   # parameters are ignored for SA
            T *= self.R
        return self.ann.get_x()[0]

    def sqa(self):
        """Run SQA with the provided QUBO (set `qubo` first); returns one bit vector per trotter.  opt.py:40-67."""
        self.E = []
        G = self.Gs
        if len(self.qubo):
            self.ann.set_qubo(_symmetric(self.qubo, self.dtype), solvers.minimize)
            self.ann.set_preferences(algorithm=solvers.algorithm.default, n_trotters=self.tro)
            self.qi()
        self.ann.prepare()
        self.ann.randomize_spin()
        N = self.ann.get_problem_size()
        n_flips_per_call = N * self.tro
        n_iters_at_G = (self.ite + n_flips_per_call - 1) // n_flips_per_call
        while G > self.Gf:
            for _ in range(n_iters_at_G):
                self.ann.anneal_one_step(G, 1. / self.Tf)
            self.E.append(self.ann.get_system_E(G, 1. / self.Tf))
            G *= self.R
        return self.ann.get_x()
