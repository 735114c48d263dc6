"""Where a field-mode sweep step spends its time (chain warp 0 of every CTA, averaged): run on the GPU box.
    python benchmarks/chain_profile.py [--N 8192 --m 512 --steps 20 --G 0.01 --beta 50]"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqaod_b200 as sq
from bench import make_problem

ap = argparse.ArgumentParser()
ap.add_argument('--N', type=int, default=8192); ap.add_argument('--m', type=int, default=512)
ap.add_argument('--steps', type=int, default=20); ap.add_argument('--warmup', type=int, default=10)
ap.add_argument('--G', type=float, default=0.01); ap.add_argument('--beta', type=float, default=50.)
ap.add_argument('--mode', default='field'); ap.add_argument('--refresh', type=int, default=0)
a = ap.parse_args()
ann = sq.dense_graph_annealer(make_problem(a.N), sq.minimize, np.float32, n_trotters=a.m)
ann.seed(1000); ann.set_sweep_mode(a.mode, a.refresh); ann.prepare(); ann.randomize_spin()
for _ in range(a.warmup):
    ann.anneal_one_step(a.G, a.beta)
ann._device.synchronize()
s0 = ann.get_stats(); t0 = time.perf_counter()
for _ in range(a.steps):
    ann.anneal_one_step(a.G, a.beta)
ann._device.synchronize()
dt = (time.perf_counter() - t0) / a.steps * 1e3
s1 = ann.get_stats()
d = {k: s1[k] - s0[k] for k in s1}
ctas, mhz = min(148, a.m), 1.965e6
ms = lambda c: c / ctas / mhz / a.steps
nwin = (a.N + 15) // 16
print('N=%d m=%d mode=%s  %.3f ms/step  acceptance %.4f  windows/step %d' % (a.N, a.m, ann.get_sweep_mode(), dt, d['accepted'] / (a.N * a.m * a.steps), nwin))
print('chain warp 0 per step per CTA [ms]: busy %.3f (gather waits %.3f, idle polls %.3f, window barrier %.3f) + waiting for fields/tables %.3f'
      % (ms(d['barrier_cycles_chain']), ms(d['chain_gather_wait_cycles']), ms(d['chain_idle_cycles']), ms(d['chain_barrier_cycles']), ms(d['chain_wait_rows_cycles'])))
print('field warp 0 busy %.3f ms, helper busy %.3f ms' % (ms(d['barrier_cycles_dot']), ms(d['helper_cycles'])))
pw = lambda c: c / ctas / a.steps / nwin
print('per CTA and window: eval passes %.2f, uncertain resolves %.3f, second-commit resolves %.3f, blocked stops %.3f, flag/idle polls %.2f'
      % (pw(d['chain_eval_passes']), pw(d['chain_uncertain_resolves']), pw(d['chain_commit_resolves']), pw(d['chain_blocked_stops']), pw(d['flag_waits'])))
