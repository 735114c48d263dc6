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
print('chain warp 0 per step per CTA [ms]: busy %.3f (gather waits %.3f, idle polls %.3f, window barrier %.3f) + waiting for fields %.3f / neighbour CTAs %.3f / tables %.3f'
      % (ms(d['barrier_cycles_chain']), ms(d['chain_gather_wait_cycles']), ms(d['chain_idle_cycles']), ms(d['chain_barrier_cycles']), ms(d['chain_wait_rows_cycles']), ms(d['chain_wait_neighbour_cycles']), ms(d['prep_cycles'])))
print('chain warps 1-3: window barrier %.3f ms each' % (ms(d['chain_barrier_cycles_others']) / 3))
print('field warp 0 busy %.3f ms, helper busy %.3f ms' % (ms(d['barrier_cycles_dot']), ms(d['helper_cycles'])))
pw = lambda c: c / ctas / a.steps / nwin
print('per CTA and window: eval passes %.2f, flag/idle polls %.2f; chain warp 0 segments [ms/step]: window start %.3f, evaluation %.3f, window end %.3f'
      % (pw(d['chain_eval_passes']), pw(d['flag_waits']), ms(d['chain_uncertain_resolves']), ms(d['chain_commit_resolves']), ms(d['chain_blocked_stops'])))

if ann.get_sweep_mode() == 'field':
    pc = ann.get_cta_profile().astype(np.float64)
    t_end = pc[:, 5] - pc[:, 5].min()
    print('per CTA (last launch): cta T wait_fields[us] wait_nb[us] chain_work[us] loop[us] end[us] accepted passes')
    for i in list(range(0, len(pc), max(1, len(pc) // 24))) + [len(pc) - 1]:
        r = pc[i]
        print('%4d %d %8.1f %8.1f %8.1f %8.1f %8.1f %6d %6d' % (i, r[0], r[1] / mhz * 1e3, r[2] / mhz * 1e3, r[3] / mhz * 1e3, r[4] / mhz * 1e3, t_end[i] / 1e3, r[6], r[7]))
    for T in sorted(set(pc[:, 0].astype(int))):
        sel = pc[:, 0] == T
        us = lambda k: pc[sel, k].mean() / mhz * 1e3
        print('T=%d: %d CTAs, mean wait_fields %.1f us, wait_nb %.1f us, chain work %.1f us, loop %.1f us; neighbour warp waits: own chain %.1f us, tables %.1f us, remote words %.1f us (loop %.1f us); table warp: Philox %.1f us, masks %.1f us, waiting %.1f us' % (
            T, sel.sum(), us(1), us(2), us(3), us(4), us(8), us(9), us(10), us(11), us(12), us(13), us(14)))
