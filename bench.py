#!/usr/bin/env python
"""bench.py -- headline benchmark of sqaod_b200: dense-graph SQA sweeps, N=8192 spins x m=512 trotters, fp32.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the reference's CPU algorithm on the host cores)

metric  = spin-flip attempts per second (BASELINE.json); one step = one annealOneStep = N*m attempts.
value   = whole-job attempts/s with the problem resident in HBM, timed with CUDA events on the launching stream.
e2e     = the same metric through the public API (sqaod_b200 -> C ABI) with host buffers: every step uploads the spin
          matrix from pinned memory, anneals one step, evaluates the energies and reads spins + energies back.
At N > 1 every GPU anneals its own replica of the problem with its own seed ("replicas only"): scaling weak.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SPINS, M_TROTTERS = 8192, 512
G_FIXED, BETA = 0.01, 1.0 / 0.02          # sqaodpy/benchmark/benchmark.py:10-11
W_SEED = 1133557                          # sqaodc/tests/perf.cpp:21


def make_problem(N, seed=W_SEED):
    """symmetric W ~ U(-0.5, 0.5), as sqaod.generate_random_symmetric_W (common/common.py:70-77), fp32."""
    rng = np.random.default_rng(seed)
    A = rng.random((N, N), dtype=np.float32) - np.float32(0.5)
    W = np.triu(A) + np.triu(A, 1).T
    return np.ascontiguousarray(W, np.float32)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(',')])
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(s[0]) for s in self.samples if s and s[0].replace('.', '').isdigit())
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for s in self.samples:
            for k, nm in enumerate(names):
                if len(s) > 3 + k and s[3 + k].lower().startswith('active'):
                    reasons.add(nm)
        smax = float(self.samples[0][1]) if self.samples else None
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons), 'samples': len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def traffic_from_profile(mode):
    p = os.path.join(ROOT, 'profiles', 'dense_sweep_ncu_summary.json' if mode == 'classic' else 'r1_field_sweep_ncu_summary.json')
    if os.path.exists(p):
        try:
            return json.load(open(p)).get('dram_bytes_per_launch')
        except Exception:
            return None
    return None


def cpu_reference_run(steps, warmup, sample_rounds=None, budget_s=15.0):
    """The reference's CPU algorithm (algoColoring, OpenMP over trotters, AVX2 dot, per-thread MT19937) restated in
    oracle/ (the reference's own library needs Eigen and cannot be built here): attempts/s on a bounded sample."""
    from oracle import pyoracle as orc
    orc.build()
    cores = orc.num_threads()
    W = make_problem(N_SPINS)
    ann = orc.DenseGraphAnnealer(W, 0, np.float32, n_trotters=M_TROTTERS, algorithm='coloring', n_workers=cores, rng='mt')
    ann.seed(1)
    ann.prepare()
    ann.randomize_spin()
    t0 = time.perf_counter()
    ann.anneal_rounds(G_FIXED, BETA, 0, 8)
    rate = 8 * M_TROTTERS / (time.perf_counter() - t0)
    if sample_rounds is None:
        per_step_budget = budget_s / max(1, steps + warmup)
        sample_rounds = int(max(8, min(N_SPINS, rate * per_step_budget / M_TROTTERS)))
    for _ in range(warmup):
        ann.anneal_rounds(G_FIXED, BETA, 0, sample_rounds)
    t0 = time.perf_counter()
    for _ in range(steps):
        ann.anneal_rounds(G_FIXED, BETA, 0, sample_rounds)
    dt = time.perf_counter() - t0
    value = steps * sample_rounds * M_TROTTERS / dt
    sample = '%d of the %d rounds of one annealOneStep (%d attempts) per step, N=%d m=%d fp32, G=%g beta=%g' % (
        sample_rounds, N_SPINS, sample_rounds * M_TROTTERS, N_SPINS, M_TROTTERS, G_FIXED, BETA)
    return value, cores, sample, dt / steps * 1e3


def run_reference(args, rank):
    if rank != 0:
        return
    value, cores, sample, ms = cpu_reference_run(args.steps, args.warmup)
    line = {
        'impl': 'reference', 'metric': 'spin-flip attempts/sec (dense SQA N=8192 m=512)', 'value': value, 'unit': 'attempts/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'dense-graph SQA N=8192 m=512 fp32 random QUBO (BASELINE.json configs[1])', 'G': G_FIXED, 'beta': BETA},
        'cpu_baseline': {'value': value, 'unit': 'attempts/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'attempts/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=60)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--N', type=int, default=N_SPINS)
    ap.add_argument('--m', type=int, default=M_TROTTERS)
    ap.add_argument('--sweep-mode', default='auto', choices=['auto', 'classic', 'field'],
                    help="how the sweep gets its local fields: 'classic' streams one J row per attempt, 'field' keeps J.q in shared "
                         "memory and streams one row per accepted flip (same Markov chain); 'auto' = the library's choice")
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    import sqaod_b200 as sq
    dev = sq.Device(local_rank)
    sq.set_active_device(dev)
    stream = torch.cuda.Stream()                # a real (non-default) stream shared by torch events and our kernels
    torch.cuda.set_stream(stream)
    dev.set_stream(stream.cuda_stream)

    N, m = args.N, args.m
    W = make_problem(N)
    ann = sq.dense_graph_annealer(W, sq.minimize, np.float32, n_trotters=m, device=dev)
    ann.seed(1000 + rank)                       # independent replica per GPU
    ann.set_sweep_mode(args.sweep_mode)
    ann.prepare()
    mode = ann.get_sweep_mode()
    ann.randomize_spin()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ----------------
    for _ in range(args.warmup):
        ann.anneal_one_step(G_FIXED, BETA)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev.launch_count(reset=True)
    stats0 = ann.get_stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        ann.anneal_one_step(G_FIXED, BETA)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = dev.launch_count()
    stats1 = ann.get_stats()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    attempts_per_step = N * m
    value = world * attempts_per_step * args.steps / (ms_max * 1e-3)

    # ---------------- end to end through the public API with host buffers ----------------
    q_host = torch.empty((m, N), dtype=torch.int8, pin_memory=True).numpy()
    q_host[...] = ann.get_spins()
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ann.set_qset(q_host)                    # H2D of the spin matrix (pinned), m x N int8
        ann.anneal_one_step(G_FIXED, BETA)
        E = ann.get_E()                         # energy kernel + D2H of m reals
        ann.get_spins(out=q_host)               # D2H of the spin matrix into the pinned buffer
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * attempts_per_step * e2e_steps / float(t.item())

    if rank == 0:
        peak, peak_src = measured_peaks()
        algo_bytes = attempts_per_step * N * 4          # one J row per attempt
        achieved = algo_bytes / (ms / args.steps * 1e-3) / 1e9
        accepted = stats1['accepted'] - stats0['accepted']
        traffic = traffic_from_profile(mode) if (N == N_SPINS and m == M_TROTTERS) else None
        line = {
            'metric': 'spin-flip attempts/sec (dense SQA N=8192 m=512)', 'value': value, 'unit': 'attempts/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_max / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'dense-graph SQA N=%d m=%d fp32 random QUBO (BASELINE.json configs[1]); one independent '
                                   'replica per GPU' % (N, m), 'G': G_FIXED, 'beta': BETA, 'algorithm': 'coloring',
                       'sweep_mode': mode + (' (local fields J.q recomputed on the tensor cores every step, kept in shared memory, one J row '
                                             'streamed per ACCEPTED flip)' if mode == 'field' else ' (one J row streamed per attempt)'),
                       'l2': 'J is %d MiB > 126 MB L2 and rows are drawn at random, no flush needed' % (N * N * 4 >> 20),
                       'acceptance_rate': accepted / float(attempts_per_step * args.steps),
                       'flag_waits': stats1['flag_waits'] - stats0['flag_waits'],
                       'busy_ms_per_step_per_cta': {k: (stats1['barrier_cycles_' + k] - stats0['barrier_cycles_' + k]) / 148.0 / 1.965e6 / args.steps
                                               for k in ('dot', 'chain')},
                       # of the chain warp's time: waiting for the local fields of the next window / for the neighbouring CTAs' data
                       'chain_wait_ms_per_step_per_cta': {k: (stats1['chain_wait_%s_cycles' % k] - stats0['chain_wait_%s_cycles' % k]) / 148.0 / 1.965e6 / args.steps
                                                          for k in ('rows', 'neighbour')}},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'attempts/s', 'h2d_bytes_per_step': m * N, 'd2h_bytes_per_step': m * N + m * 4,
                    'steps': e2e_steps, 'E_min': float(np.min(E))},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'peak_source': peak_src,
                         'kernel': 'denseSweepKernel<float,true,16,%s>' % ('true' if mode == 'field' else 'false'),
                         'algorithmic_bytes_per_launch': algo_bytes,
                         # the part of the algorithmic bytes that really came from HBM (ncu dram bytes per launch, profiles/):
                         # frac > 1 on the algorithmic figure is L2 reuse of J rows (classic mode) or rows of rejected attempts
                         # that were never needed (field mode: incremental local fields), not skipped work
                         'note': ('field mode is paced by the latency of the accept chain (one warp per CTA, ~430 ns per round), not by HBM: '
                                  'achieved/frac are the SURVEY 8d algorithmic figure (one J row per attempt, rows of rejected attempts '
                                  'are never fetched), dram_GBps/dram_frac the bytes the kernel really moves (ncu, profiles/)') if mode == 'field'
                                 else 'classic mode streams one J row per attempt and is HBM-bound; achieved > dram_GBps is L2 reuse of rows',
                         'dram_GBps': (traffic / (ms / args.steps * 1e-3) / 1e9) if traffic else None,
                         'dram_frac': (traffic / (ms / args.steps * 1e-3) / 1e9 / peak) if traffic else None},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                v, cores, sample, _ = cpu_reference_run(3, 1, budget_s=12.0)
                line['cpu_baseline'] = {'value': v, 'unit': 'attempts/s', 'cores': cores, 'kind': 'port', 'sample': sample}
            except Exception as e:      # the baseline is a reported number, never a reason to lose the GPU line
                line['cpu_baseline'] = {'value': None, 'unit': 'attempts/s', 'cores': None, 'kind': 'port', 'sample': 'failed: %s' % e}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
