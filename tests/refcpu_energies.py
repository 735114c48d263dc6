#!/usr/bin/env python
"""Final energies of the REFERENCE'S OWN sqaod.cpu dense annealer (compiled from its sources, oracle/_ref) over the seeds 0..255, for
the statistical parity tests: `python tests/refcpu_energies.py N m steps algorithm` prints one JSON line {"E": [...]}.
Same problem, schedule and seeds as tests/test_annealer_statistics_gpu.py; one worker (pinned before the libraries load), so the
chain is MT19937(seed) on any machine.  No GPU, nothing of the product library."""
import json
import os
import sys

os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[0]})
os.environ['OMP_NUM_THREADS'] = '1'
import warnings  # noqa: E402
import numpy as np  # noqa: E402

warnings.simplefilter('ignore')
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refsuite_runner  # noqa: E402
from conftest import quantized_symmetric_W  # noqa: E402


def main():
    N, m, steps, algo = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    nseeds = int(sys.argv[5]) if len(sys.argv) > 5 else 256
    sq = refsuite_runner.assemble('cpu')
    W = quantized_symmetric_W(N, 2024, np.float32)
    G0, G1 = (5.0, 0.01) if algo == 'coloring' else (2.0, 0.02)
    tau = (G1 / G0) ** (1.0 / steps)
    Gs = [G0 * tau ** k for k in range(steps)]
    ann = sq.cpu.dense_graph_annealer(W, sq.minimize, np.float32, n_trotters=m, algorithm=algo)
    E = []
    for s in range(nseeds):
        ann.seed(s); ann.prepare(); ann.randomize_spin()
        for G in Gs:
            ann.anneal_one_step(G, 1. / 0.02)
        E.append(float(np.min(ann.get_E())))
    print('REFCPU_ENERGIES ' + json.dumps({'E': E}))


if __name__ == '__main__':
    main()
