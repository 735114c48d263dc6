#!/usr/bin/env python
"""bench.py -- headline benchmark of sqaod_b200: dense-graph SQA sweeps, N=8192 spins x m=512 trotters, fp32.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the reference's own CPU back end, compiled from its sources into oracle/_ref,
                                                              on the host cores; the oracle port when that build is absent)

metric  = spin-flip attempts per second (BASELINE.json); one step = one annealOneStep = N*m attempts.
value   = whole-job attempts/s with the problem resident in HBM, timed with CUDA events on the launching stream, in the reference
          benchmark's protocol (sqaodpy/benchmark/benchmark.py:9-48): fixed operating point G=0.01, beta=50, randomize_spin, an
          untimed warm-up of >= 5 s of anneal_one_step at that point (the chain is stationary when the clock starts), then the timed
          steps.  Here: --equilibrate-seconds of untimed steps, the W warm-up steps, then exactly K timed steps.
e2e     = the same metric through the public API (sqaod_b200 -> C ABI) with host buffers: every step uploads the spin
          matrix from pinned memory, anneals one step, evaluates the energies and reads spins + energies back.
Further legs of the same line (SURVEY.md 8d asks for them because the field-mode sweep's cost follows the acceptance rate):
  transient              the K steps that follow randomize_spin + W warm-up steps WITHOUT the protocol's warm-up phase (acceptance
                         still falling: the sweep's worst case at this operating point)
  sustained              >= 2 s of back-to-back steps at the fixed point, with its own clock samples
  schedule_sweep         a fresh anneal over the reference example's whole schedule G 5 -> 0.01 (geometric), beta = 50
                         (sqaodpy/example/dense_graph_annealer.py:60-70), with the acceptance rate per fifth of the schedule
  classic                the one-J-row-per-attempt kernel (HBM-bound) on the same state
  secondary              the tensor-core rows on one GPU: calculate_E at C2 and the bipartite annealOneStep at C3 (N0=N1=4096, m=512), with
                         their tensor roofline (algorithmic flops / time against the measured dense-bf16 peak / 3); and config C1, the
                         reference's tutorial anneal (N=128, m=32, fp64, 619 steps) -- the same loop through the reference's own
                         sqaod.cpu is part of cpu_baseline
  comm                   what communicates: brute force N=40 sharded over the ranks + NCCL min/gather merge,
                         the ring-sharded N=32768 sweep (256 trotters per GPU, NVLink hand-off), 512 replicas per GPU of N=1024 m=128
At N > 1 every GPU anneals its own replica of the headline problem with its own seed ("replicas only"): scaling weak.
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


def workload_config(N=N_SPINS, m=M_TROTTERS):
    """What is measured, in words and numbers that do not depend on the run: printed identically by both arms (this library and
    --impl reference), so the two lines can be matched on it.  Everything a run finds out about itself goes under `run`."""
    return {'workload': 'dense-graph SQA N=%d m=%d fp32 random QUBO (BASELINE.json configs[1]); at N > 1 GPUs one independent replica per GPU' % (N, m),
            'G': G_FIXED, 'beta': BETA, 'algorithm': 'coloring',
            'protocol': 'sqaodpy/benchmark/benchmark.py:9-48: randomize_spin, an untimed warm-up phase of anneal_one_step at the operating point '
                        '(the reference: batches until one takes >= 5 s), then W warm-up steps and K timed steps',
            'l2': 'J is %d MiB > 126 MB L2 and rows are drawn at random, no flush needed' % (N * N * 4 >> 20)}


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


def measured_tensor_peak():
    """dense bf16 TFLOP/s of this pool's B200s (burst figure: the kernels below are timed alone, a few launches each)"""
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['bf16_tflops']), 'measured (MEASURED_PEAKS.json bf16_tflops)'
        except Exception:
            pass
    return 2250.0, 'fallback (B200_PROFILING.md nominal dense bf16)'


def ncu_traffic_note(mode):
    """DRAM bytes per launch of the sweep kernel in the committed ncu capture of THIS round (profiles/), with the acceptance rate it
    was captured at -- reported next to the live traffic model, never used to compute a number of the line."""
    p = os.path.join(ROOT, 'profiles', 'r2_classic_sweep_ncu_summary.json' if mode == 'classic' else 'r2_field_sweep_ncu_summary.json')
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {'dram_bytes_per_launch': d.get('dram_bytes_per_launch'), 'acceptance_rate_at_capture': d.get('acceptance_rate'), 'file': os.path.relpath(p, ROOT)}
        except Exception:
            return None
    return None


def cpu_port_run(steps, warmup, sample_rounds=None, budget_s=15.0):
    """The reference's CPU algorithm (algoColoring, OpenMP over trotters, AVX2 dot, per-thread MT19937) restated in oracle/oracle.cpp,
    on a bounded sample of rounds per step.  tests/test_oracle_vs_reference_cpu.py pins it bit for bit against the compiled reference."""
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
    return value, cores, sample, dt / steps * 1e3, 'port'


REFCPU_GLUE = os.path.join(ROOT, 'oracle', '_ref', 'refsuite', 'glue_cpu', 'cpu_dg_annealer.so')


def cpu_reference_run(steps, warmup, budget_s=15.0, max_total_s=150.0):
    """The reference's own CPU implementation of the path on the host cores.

    oracle/_ref holds the reference's CPU back end compiled from its own sources (`make -C oracle refcpu`: sqaodc/common, sqaodc/cpu and
    its CPython glue, unmodified; the absent Eigen replaced by oracle/eigen_standin, which the annealing loop does not touch).  It is
    driven through the reference's own Python API, sqaod.cpu.dense_graph_annealer(...).anneal_one_step(G, beta)
    (sqaodpy/sqaod/cpu/dense_graph_annealer.py): every step is one WHOLE annealOneStep (N rounds x m trotters), all host threads
    (CPUDenseGraphAnnealer.cpp:303-338).  That API has no partial step, so when whole steps would not fit the time limit -- or the build
    is absent -- the bounded-sample port above is timed instead and the line says kind = "port"."""
    if not os.path.exists(REFCPU_GLUE):
        return cpu_port_run(steps, warmup, budget_s=budget_s)
    try:
        return _cpu_reference_compiled(steps, warmup, budget_s, max_total_s)
    except Exception as e:      # e.g. the compiled glue does not load on this host: the port is the same algorithm (pinned against it)
        sys.stderr.write('compiled reference unavailable (%s: %s); timing the oracle port instead\n' % (type(e).__name__, e))
        return cpu_port_run(steps, warmup, budget_s=budget_s)


def _cpu_reference_compiled(steps, warmup, budget_s, max_total_s):
    cores = len(os.sched_getaffinity(0))
    if os.environ.get('OMP_NUM_THREADS') in (None, '', '1'):    # torchrun pins it to 1 for N > 1; the reference sizes its pool from the affinity mask
        os.environ['OMP_NUM_THREADS'] = str(cores)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import refsuite_runner
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')                          # the reference's docstrings predate Python 3.12's escape-sequence check
        sq = refsuite_runner.assemble('cpu')
        W = make_problem(N_SPINS)
        ann = sq.cpu.dense_graph_annealer(W, sq.minimize, np.float32, n_trotters=M_TROTTERS, algorithm=sq.algorithm.coloring)
    ann.seed(1)
    ann.prepare()
    ann.randomize_spin()
    t0 = time.perf_counter()
    ann.anneal_one_step(G_FIXED, BETA)                           # first warm-up step, also the time estimate
    t_step = time.perf_counter() - t0
    if t_step * (steps + max(warmup, 1) - 1) > max_total_s:
        del ann
        return cpu_port_run(steps, warmup, budget_s=budget_s)
    for _ in range(warmup - 1):
        ann.anneal_one_step(G_FIXED, BETA)
    t0 = time.perf_counter()
    for _ in range(steps):
        ann.anneal_one_step(G_FIXED, BETA)
    dt = time.perf_counter() - t0
    value = steps * N_SPINS * M_TROTTERS / dt
    sample = ('whole annealOneStep per step (%d rounds x %d trotters = %d attempts), N=%d m=%d fp32, G=%g beta=%g, sqaod.cpu '
              'dense_graph_annealer compiled from the reference sources (oracle/_ref)') % (
        N_SPINS, M_TROTTERS, N_SPINS * M_TROTTERS, N_SPINS, M_TROTTERS, G_FIXED, BETA)
    return value, cores, sample, dt / steps * 1e3, 'reference'


def run_reference(args, rank):
    if rank != 0:
        return
    value, cores, sample, ms, kind = cpu_reference_run(args.steps, args.warmup)
    line = {
        'impl': 'reference', 'metric': 'spin-flip attempts/sec (dense SQA N=8192 m=512)', 'value': value, 'unit': 'attempts/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(),
        'cpu_baseline': {'value': value, 'unit': 'attempts/s', 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': 'attempts/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def device_facts(torch, local_rank, clocks):
    props = torch.cuda.get_device_properties(local_rank)
    mhz = (clocks or {}).get('sm_mhz') or props.clock_rate / 1e3
    return props.multi_processor_count, float(mhz)


def timed_steps(torch, stream, ann, Gs, beta, barrier):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for G in Gs:
        ann.anneal_one_step(G, beta)
    ev1.record(stream)
    barrier()
    return ev0.elapsed_time(ev1)


def comm_legs(args, torch, dist, sq, dev, rank, local_rank, world, barrier):
    """The workloads of SURVEY.md 8e that communicate, at the job's world size (world == 1: the unsharded counterpart)."""
    import hashlib
    from sqaod_b200 import multigpu
    out = {}
    # ---- C4: dense brute force N = 40, x range sharded over the ranks, one NCCL exchange (all_reduce MIN + all_gather)
    try:
        Nbf = args.bf_N
        rng = np.random.default_rng(40)
        A = np.rint((rng.random((Nbf, Nbf)) - 0.5) * 16384) / 16384.
        Wbf = np.asarray(np.triu(A) + np.triu(A, 1).T, np.float32)
        multigpu.sharded_dense_bf_search(Wbf[:16, :16].copy(), 0, np.float32)           # warm-up (kernels, NCCL)
        barrier()
        t0 = time.perf_counter()
        E, xs = multigpu.sharded_dense_bf_search(Wbf, 0, np.float32)
        torch.cuda.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        h = hashlib.sha256(np.float64(E).tobytes() + np.asarray(xs, np.int8).tobytes()).hexdigest()[:16]
        out['bf_n40_sharded'] = {'N': Nbf, 'states': float(1 << Nbf), 'seconds': dt, 'states_per_s': float(1 << Nbf) / dt, 'E_min': float(E),
                                 'n_argmin': len(xs), 'result_sha16': h,
                                 'exchange': 'all_reduce(MIN) + all_gather of argmin lists (NCCL)' if world > 1 else 'none (1 rank)'}
    except Exception as e:
        out['bf_n40_sharded'] = {'error': str(e)[:300]}
    # ---- C5a: independent replicas N = 1024, m = 128, 512 per GPU, no data-path collective, final MIN reduce
    try:
        Nr, mr, per = 1024, 128, args.replicas_per_gpu
        Wr = make_problem(Nr, seed=1024)
        Gs = [5.0 * (0.01 / 5.0) ** (k / 7.0) for k in range(8)]
        multigpu.anneal_replicas(Wr, world, Gs[:2], BETA, np.float32, n_trotters=mr)     # warm-up
        barrier()
        t0 = time.perf_counter()
        best, _, _, _ = multigpu.anneal_replicas(Wr, per * world, Gs, BETA, np.float32, n_trotters=mr)
        torch.cuda.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        out['replicas_c5a'] = {'N': Nr, 'm': mr, 'replicas': per * world, 'steps_per_replica': len(Gs), 'seconds': dt,
                               'attempts_per_s': float(per * world) * len(Gs) * Nr * mr / dt, 'best_E': float(best),
                               'note': 'wall clock incl. prepare / randomize / get_E; J resident per GPU; final all_reduce(MIN)'}
    except Exception as e:
        out['replicas_c5a'] = {'error': str(e)[:300]}
    # ---- C5b: ONE instance N = 32768, trotter ring sharded over the ranks (256 trotters per GPU), NVLink P2P hand-off
    try:
        Nq, per_m, steps = args.ring_N, 256, 3
        if world > 1:
            ring = multigpu.RingShardedDenseAnnealer(('random', Nq, 32768), 0, np.float32, n_trotters=per_m * world)
            ring.seed(7); ring.prepare(); ring.randomize_spin()
            ann = ring.ann
            step = ring.anneal_one_step
        else:
            ann = sq.dense_graph_annealer(None, sq.minimize, np.float32, device=dev)
            ann.set_qubo_random(Nq, 32768)
            ann.set_preferences(n_trotters=per_m)
            ann.seed(7); ann.prepare(); ann.randomize_spin()
            step = ann.anneal_one_step
        step(G_FIXED, BETA)
        barrier()
        s0 = ann.get_stats()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(torch.cuda.current_stream())
        for _ in range(steps):
            step(G_FIXED, BETA)
        ev1.record(torch.cuda.current_stream())
        barrier()
        ms = ev0.elapsed_time(ev1) / steps
        s1 = ann.get_stats()
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        sms, mhz = device_facts(torch, local_rank, None)
        wait_ms = (s1['chain_wait_neighbour_cycles'] - s0['chain_wait_neighbour_cycles']) / float(min(sms, per_m)) / (mhz * 1e3) / steps
        out['ring_c5b'] = {'N': Nq, 'm': per_m * world, 'trotters_per_gpu': per_m, 'ms_per_step': ms_max,
                           'attempts_per_s': float(Nq) * per_m * world / (ms_max * 1e-3), 'sweep_mode': ann.get_sweep_mode(),
                           'neighbour_wait_ms_per_step_per_cta': wait_ms, 'neighbour_wait_frac': wait_ms / ms_max,
                           'flag_polls_per_step': (s1['flag_waits'] - s0['flag_waits']) / steps,
                           'exchange': ('accept words + conflict flags + per-sweep edge-trotter push over NVLink (CUDA IPC peer memory), no NCCL in the data path'
                                        if world > 1 else 'none (1 rank: the whole ring on one GPU)')}
        del ann
    except Exception as e:
        out['ring_c5b'] = {'error': str(e)[:300]}
    return out


def c1_problem():
    """BASELINE.json configs[0]: W ~ U(-0.5, 0.5) symmetric (seed 13255), and the tutorial's schedule G = 5, 5 * 0.99, ... >= 0.01"""
    rng = np.random.default_rng(13255)
    A = rng.random((128, 128)) - 0.5
    Gs, G = [], 5.0
    while 0.01 <= G:
        Gs.append(G)
        G *= 0.99
    return np.triu(A) + np.triu(A, 1).T, Gs


def c1_tutorial(factory, W, Gs, **kw):
    """the loop of sqaodpy/example/dense_graph_annealer.py:44-70 on any package with the reference's solver API; (seconds, E_min, E_mean)"""
    a = factory(W, kw.pop('optimize'), np.float64, n_trotters=32, **kw)
    a.seed(13255)
    a.prepare()
    a.randomize_spin()
    t0 = time.perf_counter()
    for g in Gs:
        a.anneal_one_step(g, BETA)
    E = np.asarray(a.get_E())                             # waits for the device
    return time.perf_counter() - t0, float(E.min()), float(E.mean())


def c1_reference_cpu():
    """C1 through the reference's own sqaod.cpu (compiled from its sources, oracle/_ref): part of the cpu_baseline leg"""
    if not os.path.exists(REFCPU_GLUE):
        return None
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import refsuite_runner
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = refsuite_runner.assemble('cpu')
    Wc, Gs = c1_problem()
    sec, emin, emean = c1_tutorial(ref.cpu.dense_graph_annealer, Wc, Gs, optimize=ref.minimize)
    return {'N': 128, 'm': 32, 'dtype': 'f64', 'steps': len(Gs), 'seconds': sec, 'attempts_per_s': len(Gs) * 128 * 32 / sec, 'E_min': emin,
            'E_mean': emean, 'cores': len(os.sched_getaffinity(0)), 'kind': 'reference'}


def secondary_legs(args, torch, sq, dev, stream, ann, rank):
    """The tensor-core rows of the path (a6, a7) on one GPU, so that they have a driver-side number next to the
    headline: calculate_E at C2 (on the headline annealer's state) and the bipartite annealOneStep at C3, both through the split-bf16
    tcgen05 spin GEMM.  Tensor roofline: algorithmic flops (the three bf16 passes of the split are NOT counted as extra
    flops) / time against the measured dense-bf16 peak / 3."""
    out = {}
    peak, peak_src = measured_tensor_peak()

    def timed(fn, reps, warm):      # per-GPU figures (rank 0's is printed): no collective in here, so a failure on one rank cannot hang the others
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def tensor_roofline(flops, ms):
        tf = flops / (ms * 1e-3) / 1e12
        return {'bound': 'tensor', 'achieved': tf, 'peak': peak / 3.0, 'unit': 'TFLOP/s', 'frac': tf / (peak / 3.0), 'peak_source': peak_src,
                'note': 'algorithmic flops / time; peak = dense bf16 peak / 3 (split-precision GEMM: hi + mid + lo planes of J, one pass each)'}
    try:    # a7: E_y = -c - h.q_y - q_y^T J q_y for all m trotters (GEMM m x N x N + row dots), result left on the device (async read-back)
        N, m = args.N, args.m
        ms = timed(ann.calculate_E, 10, 2)
        flops = 2.0 * m * N * N + 4.0 * m * N
        out['calculate_E_c2'] = {'N': N, 'm': m, 'ms': ms, 'algorithmic_flops': flops, 'kernel': 'tcSpinGemmKernel (tcgen05, bf16 x 3) + row-dot',
                                 'roofline': tensor_roofline(flops, ms)}
    except Exception as e:
        out['calculate_E_c2'] = {'error': str(e)[:300]}
    try:    # a6: C3, one step = two half steps, each dEmat = qFixed . J(^T) on the tensor cores + one fused coloured flip launch
        N0 = N1 = args.bipartite_N
        mb = args.m
        rng = np.random.default_rng(W_SEED + 3)
        b0 = rng.random(N0, dtype=np.float32) - np.float32(0.5)
        b1 = rng.random(N1, dtype=np.float32) - np.float32(0.5)
        Wb = rng.random((N1, N0), dtype=np.float32) - np.float32(0.5)
        bg = sq.bipartite_graph_annealer(b0, b1, Wb, sq.minimize, np.float32, n_trotters=mb, device=dev)
        bg.seed(2000 + rank); bg.prepare(); bg.randomize_spin()
        ms = timed(lambda: bg.anneal_one_step(G_FIXED, BETA), 50, 5)
        flops = 4.0 * mb * N0 * N1
        out['bipartite_c3'] = {'N0': N0, 'N1': N1, 'm': mb, 'ms_per_step': ms, 'attempts_per_s': (N0 + N1) * mb / (ms * 1e-3),
                               'algorithmic_flops_per_step': flops, 'E_min': float(np.min(bg.get_E())),
                               'kernels': 'per half step: tcSpinGemmKernel (tcgen05, bf16 x 3) + bgFlipFusedKernel',
                               'roofline': tensor_roofline(flops, ms)}
        del bg
    except Exception as e:
        out['bipartite_c3'] = {'error': str(e)[:300]}
    try:    # BASELINE.json configs[0] (C1): the reference's tutorial anneal, N = 128, m = 32, fp64, G 5 -> 0.01 with G *= 0.99 (619 steps),
            # beta = 50, seed 13255 (sqaodpy/example/dense_graph_annealer.py:22-70), the one config the reference itself runs on a CPU: wall
            # clock here; the cpu_baseline leg runs the same loop through the reference's own sqaod.cpu
        Wc, Gs = c1_problem()
        c1_tutorial(sq.dense_graph_annealer, Wc, Gs, optimize=sq.minimize, device=dev)        # first pass: allocations, module load
        sec, emin, emean = c1_tutorial(sq.dense_graph_annealer, Wc, Gs, optimize=sq.minimize, device=dev)
        out['c1_tutorial'] = {'N': 128, 'm': 32, 'dtype': 'f64', 'steps': len(Gs), 'seconds': sec, 'attempts_per_s': len(Gs) * 128 * 32 / sec,
                              'E_min': emin, 'E_mean': emean,
                              'note': 'wall clock of the whole tutorial loop incl. the final get_E; launch-latency bound at this size; the same '
                                      'loop through the reference CPU solver: cpu_baseline.c1_tutorial'}
    except Exception as e:
        out['c1_tutorial'] = {'error': str(e)[:300]}
    return out


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
    ap.add_argument('--equilibrate-seconds', type=float, default=5.0,
                    help="untimed steps at the operating point before the warm-up and timed steps: the reference protocol's warm-up "
                         "phase (benchmark.py:17-27, batches of steps until one takes >= 5 s); 0: time right after randomize_spin")
    ap.add_argument('--sustain-seconds', type=float, default=2.5, help='length of the sustained leg (0: skip)')
    ap.add_argument('--schedule-steps', type=int, default=100, help='steps of the G 5 -> 0.01 schedule leg (0: skip)')
    ap.add_argument('--no-classic-leg', action='store_true')
    ap.add_argument('--no-comm-legs', action='store_true', help='skip the brute-force / ring / replica legs')
    ap.add_argument('--no-secondary-legs', action='store_true', help='skip the calculate_E / bipartite (tensor-core) legs')
    ap.add_argument('--bipartite-N', type=int, default=4096, help='N0 = N1 of the bipartite leg (C3: 4096)')
    ap.add_argument('--bf-N', type=int, default=40)
    ap.add_argument('--ring-N', type=int, default=32768)
    ap.add_argument('--replicas-per-gpu', type=int, default=512)
    ap.add_argument('--leg-timeout', type=float, default=420.0,
                    help='seconds any one of the extra legs (sustained, schedule, classic, secondary, comm, cpu_baseline) may take before the '
                         'line is printed without it')
    ap.add_argument('--quick', action='store_true', help='headline + e2e legs only')
    args = ap.parse_args()
    if args.quick:
        args.sustain_seconds, args.schedule_steps, args.no_classic_leg, args.no_comm_legs, args.no_cpu_baseline = 0.0, 0, True, True, True
        args.equilibrate_seconds = min(args.equilibrate_seconds, 1.0)
        args.no_secondary_legs = True
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

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    attempts_per_step = N * m

    # ---------------- transient: the K steps right after randomize_spin + W warm-up steps (no protocol warm-up phase) ----------------
    for _ in range(args.warmup):
        ann.anneal_one_step(G_FIXED, BETA)
    barrier()
    t_acc0 = ann.get_stats()['accepted']
    ms_tr = max_over_ranks(timed_steps(torch, stream, ann, [G_FIXED] * args.steps, BETA, barrier))
    t_acc = ann.get_stats()['accepted'] - t_acc0
    transient = {'steps': args.steps, 'ms_per_step': ms_tr / args.steps, 'value': world * attempts_per_step * args.steps / (ms_tr * 1e-3),
                 'unit': 'attempts/s', 'acceptance_rate': t_acc / float(attempts_per_step * args.steps),
                 'note': 'steps %d..%d after randomize_spin: acceptance still falling' % (args.warmup, args.warmup + args.steps - 1)}
    if mode == 'field':     # same live traffic model as the headline's roofline (rank 0's counters): more accepted flips, more rows to move
        tr_traffic = (t_acc / float(args.steps)) * (N * 4 + 32 * 32) + m * N * 4 + 2 * m * N
        tr_peak = measured_peaks()[0]
        transient['roofline'] = {'bound': 'latency', 'traffic': tr_traffic, 'achieved': tr_traffic / (ms_tr / args.steps * 1e-3) / 1e9, 'peak': tr_peak,
                                 'unit': 'GB/s', 'frac': tr_traffic / (ms_tr / args.steps * 1e-3) / 1e9 / tr_peak}

    # ---------------- the reference protocol's warm-up phase: untimed steps at the operating point until the chain is stationary ----------------
    equil_steps = 0
    if args.equilibrate_seconds > 0:
        batch = max(10, int(0.25e3 / max(ms_tr / args.steps, 1e-3)))     # ~0.25 s of steps between clock reads
        t_eq = time.perf_counter()
        while True:
            for _ in range(batch):
                ann.anneal_one_step(G_FIXED, BETA)
            torch.cuda.synchronize()
            equil_steps += batch
            # every rank sees the same (max over ranks) elapsed time, so all ranks run the same number of batches
            if max_over_ranks(time.perf_counter() - t_eq) >= args.equilibrate_seconds:
                break

    # ---------------- headline: device-resident throughput at the fixed operating point ----------------
    for _ in range(args.warmup):
        ann.anneal_one_step(G_FIXED, BETA)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev.launch_count(reset=True)
    stats0 = ann.get_stats()
    ms = timed_steps(torch, stream, ann, [G_FIXED] * args.steps, BETA, barrier)
    launches = dev.launch_count()
    stats1 = ann.get_stats()
    clocks = sampler.stop() if rank == 0 else None
    ms_max = max_over_ranks(ms)
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
    e2e_value = world * attempts_per_step * e2e_steps / max_over_ranks(time.perf_counter() - t0)

    # ---------------- the line so far (rank 0): headline, e2e, roofline.  The legs below only add to it ----------------
    line = {}
    peak, peak_src = measured_peaks()
    algo_bytes = attempts_per_step * N * 4              # one J row per attempt
    if rank == 0:
        sms, mhz = device_facts(torch, local_rank, clocks)
        ctas = min(sms, m)
        step_s = ms / args.steps * 1e-3
        accepted = stats1['accepted'] - stats0['accepted']
        acc_rate = accepted / float(attempts_per_step * args.steps)
        cyc_ms = lambda key: (stats1[key] - stats0[key]) / float(ctas) / (mhz * 1e3) / args.steps
        if mode == 'field':
            # bytes the field-mode sweep has to move per launch, from this run's counters: one J row + 2K 32-byte sectors of cross
            # terms per ACCEPTED flip, the field rows in (once per step), spins in and out.  L2 hits on J can only lower it.
            traffic = (accepted / float(args.steps)) * (N * 4 + 32 * 32) + m * N * 4 + 2 * m * N
            bound = 'latency (accept chain + per-window hand-offs between the warps and CTAs), not HBM'
            traffic_src = 'live model from this run: accepted flips x (J row + cross-term sectors) + field rows + spins'
        else:
            traffic = float(algo_bytes)
            bound = 'hbm'
            traffic_src = 'one J row per attempt (L2 hits lower the DRAM share)'
        roofline = {'bound': bound, 'achieved': traffic / step_s / 1e9, 'peak': peak, 'unit': 'GB/s', 'frac': traffic / step_s / 1e9 / peak,
                    'traffic': traffic, 'traffic_source': traffic_src, 'traffic_ncu': ncu_traffic_note(mode), 'peak_source': peak_src,
                    'kernel': 'denseSweepKernel<float,true,16,%s>' % ('true' if mode == 'field' else 'false'),
                    'algorithmic_bytes_per_launch': algo_bytes, 'algorithmic_GBps': algo_bytes / step_s / 1e9,
                    'algorithmic_speedup': algo_bytes / step_s / 1e9 / peak,
                    'note': 'achieved/frac: bytes the kernel moves per launch / launch time / measured copy peak.  algorithmic_*: the SURVEY 8d '
                            'figure (one J row per attempt); in field mode rows of rejected attempts are never fetched, so algorithmic_speedup '
                            '> 1 is an algorithmic gain, not bandwidth'}
        line = {
            'metric': 'spin-flip attempts/sec (dense SQA N=8192 m=512)', 'value': value, 'unit': 'attempts/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_max / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(N, m),
            # what this run found out about itself
            'run': {'sweep_mode': mode + (' (local fields h + 2 J.q from the tensor-core spin GEMM, carried in shared memory, one J row streamed per '
                                          'ACCEPTED flip; one accept-chain warp per trotter)' if mode == 'field' else ' (one J row streamed per attempt)'),
                    'acceptance_rate': acc_rate,
                    'equilibration_steps': equil_steps, 'equilibration_seconds': args.equilibrate_seconds,
                    'flag_waits': stats1['flag_waits'] - stats0['flag_waits'],
                    'device': {'sms': sms, 'sm_mhz_used_for_cycle_conversion': mhz},
                    # chain warp 0 and field warp 0 of every CTA, averaged: where a step goes
                    'ms_per_step_per_cta': {'chain_busy': cyc_ms('barrier_cycles_chain'), 'chain_wait_fields': cyc_ms('chain_wait_rows_cycles'),
                                            'chain_wait_neighbour_ctas': cyc_ms('chain_wait_neighbour_cycles'),
                                            'field_warp_busy': cyc_ms('barrier_cycles_dot')}},
            'schedule_sweep': None,
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'attempts/s', 'h2d_bytes_per_step': m * N, 'd2h_bytes_per_step': m * N + m * 4,
                    'steps': e2e_steps, 'E_min': float(np.min(E))},
            'gpu_launches': int(launches),
            'roofline': roofline,
            'transient': transient,
            'sustained': None,
            'classic': None,
            'secondary': None,
            'comm': None,
        }

    # The legs below are reported extras.  None of them may cost the line: an exception is recorded in the leg, and if a leg does not
    # come back within --leg-timeout seconds (a rank lost inside a collective, a kernel that never ends) the watchdog prints the line
    # as it stands and ends the process.
    emitted = threading.Lock()
    leg_now = ['(none)']

    def emit():
        if rank == 0 and emitted.acquire(False):
            print(json.dumps(line), flush=True)

    def on_timeout():
        if rank == 0:
            line['watchdog'] = 'leg "%s" did not finish within %.0f s; the line was printed without it' % (leg_now[0], args.leg_timeout)
        emit()
        os._exit(0)

    def run_leg(name, fn):
        leg_now[0] = name
        dog = threading.Timer(args.leg_timeout, on_timeout)
        dog.daemon = True
        dog.start()
        try:
            return fn()
        except Exception as e:
            return {'error': '%s: %s' % (type(e).__name__, str(e)[:300])}
        finally:
            dog.cancel()

    # ---------------- sustained: >= 2 s back to back, own clock samples (power-cap behaviour on record) ----------------
    def leg_sustained():
        n_sus = int(max(args.steps, np.ceil(args.sustain_seconds * 1e3 / (ms / args.steps))))
        sampler2 = ClockSampler(local_rank)
        if rank == 0:
            sampler2.start()
        ms_sus = timed_steps(torch, stream, ann, [G_FIXED] * n_sus, BETA, barrier)
        clocks2 = sampler2.stop() if rank == 0 else None
        ms_sus_max = max_over_ranks(ms_sus)
        return {'steps': n_sus, 'seconds': ms_sus_max * 1e-3, 'ms_per_step': ms_sus_max / n_sus,
                'value': world * attempts_per_step * n_sus / (ms_sus_max * 1e-3), 'unit': 'attempts/s', 'clocks': clocks2}
    if args.sustain_seconds > 0:
        line['sustained'] = run_leg('sustained', leg_sustained)

    # ---------------- the whole annealing schedule: G 5 -> 0.01 geometric, beta = 50, from random spins ----------------
    def leg_schedule():
        S = args.schedule_steps
        Gs = [5.0 * (0.01 / 5.0) ** (k / float(S - 1)) for k in range(S)]
        ann.randomize_spin()
        parts, tot_ms, tot_acc = [], 0.0, 0
        n_part = 5
        for i in range(n_part):
            chunk = Gs[i * S // n_part:(i + 1) * S // n_part]
            a0 = ann.get_stats()['accepted']
            ms_c = max_over_ranks(timed_steps(torch, stream, ann, chunk, BETA, barrier))
            acc = ann.get_stats()['accepted'] - a0
            parts.append({'G_from': chunk[0], 'G_to': chunk[-1], 'steps': len(chunk), 'ms_per_step': ms_c / len(chunk),
                          'acceptance_rate': acc / float(attempts_per_step * len(chunk))})
            tot_ms += ms_c
            tot_acc += acc
        return {'protocol': 'randomize_spin, then G = 5 -> 0.01 geometric over %d steps at beta = 50 (the range of '
                            'sqaodpy/example/dense_graph_annealer.py:60-70)' % S,
                'steps': S, 'ms_per_step': tot_ms / S, 'value': world * attempts_per_step * S / (tot_ms * 1e-3), 'unit': 'attempts/s',
                'acceptance_rate': tot_acc / float(attempts_per_step * S), 'by_fifth': parts, 'E_min_final': float(np.min(ann.get_E()))}
    if args.schedule_steps > 0:
        schedule = run_leg('schedule_sweep', leg_schedule)
        if rank == 0:
            line['schedule_sweep'] = schedule

    # ---------------- the classic (one J row per attempt, HBM-bound) kernel on the equilibrated state ----------------
    def leg_classic():
        q_now = ann.get_spins()
        ann_c = sq.dense_graph_annealer(W, sq.minimize, np.float32, n_trotters=m, device=dev)
        ann_c.seed(1000 + rank)
        ann_c.set_sweep_mode('classic')
        ann_c.prepare()
        ann_c.set_qset(q_now)
        for _ in range(2):
            ann_c.anneal_one_step(G_FIXED, BETA)
        n_c = 6
        c0 = ann_c.get_stats()['accepted']
        ms_c = max_over_ranks(timed_steps(torch, stream, ann_c, [G_FIXED] * n_c, BETA, barrier))
        acc_c = ann_c.get_stats()['accepted'] - c0
        cs = ms_c / n_c * 1e-3
        return {'ms_per_step': ms_c / n_c, 'value': world * attempts_per_step * n_c / (ms_c * 1e-3), 'unit': 'attempts/s', 'steps': n_c,
                'acceptance_rate': acc_c / float(attempts_per_step * n_c), 'kernel': 'denseSweepKernel<float,true,16,false>',
                'algorithmic_bytes_per_launch': algo_bytes,
                'roofline': {'bound': 'hbm', 'achieved': algo_bytes / cs / 1e9, 'peak': peak, 'unit': 'GB/s', 'frac': algo_bytes / cs / 1e9 / peak,
                             'note': 'algorithmic bytes (one row per attempt) / time; the share that misses L2 is in profiles/ (ncu)',
                             'traffic_ncu': ncu_traffic_note('classic')}}
    if not args.no_classic_leg and mode != 'classic':
        line['classic'] = run_leg('classic', leg_classic)

    if not args.no_secondary_legs:
        line['secondary'] = run_leg('secondary', lambda: secondary_legs(args, torch, sq, dev, stream, ann, rank))

    if not args.no_comm_legs:
        line['comm'] = run_leg('comm', lambda: comm_legs(args, torch, dist, sq, dev, rank, local_rank, world, barrier))

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        def leg_cpu():
            v, cores, sample, _, kind = cpu_reference_run(3, 1, budget_s=12.0, max_total_s=15.0)
            cb = {'value': v, 'unit': 'attempts/s', 'cores': cores, 'kind': kind, 'sample': sample}
            if not args.no_secondary_legs:
                try:        # config C1 on the host, next to secondary.c1_tutorial
                    cb['c1_tutorial'] = c1_reference_cpu()
                except Exception as e:
                    cb['c1_tutorial'] = {'error': str(e)[:300]}
            return cb
        cb = run_leg('cpu_baseline', leg_cpu)
        line['cpu_baseline'] = cb if 'error' not in cb else {'value': None, 'unit': 'attempts/s', 'cores': None, 'kind': 'port',
                                                             'sample': 'failed: %s' % cb['error']}
    emit()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
