"""bench.py's control flow and JSON contract, run on the CPU against stand-ins for the device (no GPU here): every leg executes, the
line parses and carries the keys the driver reads.  The stand-ins only count calls and return plausible numbers -- nothing is measured;
what this guards is that an edit of bench.py cannot lose the round's GPU line to a typo."""
import io
import json
import os
import sys
import types
import contextlib
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeAnnealer(object):
    def __init__(self, N, m):
        self.N, self.m, self.steps, self.mode = N, m, 0, 'field'
        self.q = np.ones((m, N), np.int8)

    def seed(self, s): pass
    def set_sweep_mode(self, mode): self.mode = 'field' if mode == 'auto' else mode
    def get_sweep_mode(self): return self.mode
    def prepare(self): pass
    def randomize_spin(self): pass
    def set_qubo_random(self, N, seed): self.N = N
    def set_preferences(self, **kw): self.m = kw.get('n_trotters', self.m)
    def anneal_one_step(self, G, beta): self.steps += 1
    def set_qset(self, q): pass
    def get_E(self): return np.full(self.m, -1.0, np.float32)
    def calculate_E(self): pass

    def get_spins(self, out=None):
        if out is not None:
            out[...] = self.q
            return out
        return self.q.copy()

    def get_stats(self):
        return {'accepted': 10 * self.steps, 'flag_waits': self.steps, 'barrier_cycles_chain': 1000 * self.steps,
                'chain_wait_rows_cycles': 100 * self.steps, 'chain_wait_neighbour_cycles': 50 * self.steps, 'barrier_cycles_dot': 500 * self.steps}


class _Patcher(object):
    """the two calls of pytest's monkeypatch that _install_fakes uses, for the worker processes of the two-rank run"""
    def setattr(self, obj, name, value): setattr(obj, name, value)
    def setitem(self, mapping, key, value): mapping[key] = value


def _install_fakes(monkeypatch):
    import torch
    clock = {'ms': 0.0}

    class Event(object):
        def __init__(self, enable_timing=False): self.t = 0.0
        def record(self, stream=None):
            clock['ms'] += 7.0
            self.t = clock['ms']
        def elapsed_time(self, other): return other.t - self.t

    class Stream(object):
        cuda_stream = 0

    monkeypatch.setattr(torch.cuda, 'set_device', lambda i: None)
    monkeypatch.setattr(torch.cuda, 'Stream', Stream)
    monkeypatch.setattr(torch.cuda, 'set_stream', lambda s: None)
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a: Stream())
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a: None)
    monkeypatch.setattr(torch.cuda, 'Event', Event)
    monkeypatch.setattr(torch.cuda, 'get_device_properties', lambda i: types.SimpleNamespace(multi_processor_count=148, clock_rate=1965000))
    real_tensor, real_empty = torch.tensor, torch.empty
    monkeypatch.setattr(torch, 'tensor', lambda *a, **k: real_tensor(*a, **{x: y for x, y in k.items() if x != 'device'}))
    monkeypatch.setattr(torch, 'empty', lambda *a, **k: real_empty(*a, **{x: y for x, y in k.items() if x != 'pin_memory'}))

    sq = types.ModuleType('sqaod_b200')
    sq.minimize = 0

    class Device(object):
        def __init__(self, i): self.n = 0
        def set_stream(self, s): pass
        def launch_count(self, reset=False): return 0 if reset else 46
    sq.Device = Device
    sq.set_active_device = lambda d: None
    sq.dense_graph_annealer = lambda W, opt, dtype, n_trotters=None, device=None: _FakeAnnealer(0 if W is None else W.shape[0], n_trotters or 1)
    sq.maximize = 1
    sq.bipartite_graph_annealer = lambda b0, b1, W, opt, dtype, n_trotters=None, device=None: _FakeAnnealer(W.shape[1], n_trotters or 1)
    mg = types.ModuleType('sqaod_b200.multigpu')
    mg.sharded_dense_bf_search = lambda W, opt, dtype: (np.float32(-1.5), [np.zeros(W.shape[0], np.int8)])
    mg.anneal_replicas = lambda W, R, Gs, beta, dtype, n_trotters=None: (-2.0, None, None, None)

    class RingShardedDenseAnnealer(object):
        def __init__(self, problem, optimize, dtype, n_trotters=None):
            self.ann = _FakeAnnealer(problem[1], n_trotters)
        def seed(self, s): pass
        def prepare(self): pass
        def randomize_spin(self): pass
        def anneal_one_step(self, G, beta): self.ann.anneal_one_step(G, beta)
    mg.RingShardedDenseAnnealer = RingShardedDenseAnnealer
    sq.multigpu = mg
    monkeypatch.setitem(sys.modules, 'sqaod_b200', sq)
    monkeypatch.setitem(sys.modules, 'sqaod_b200.multigpu', mg)


def _run(monkeypatch, argv):
    _install_fakes(monkeypatch)
    for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE'):
        monkeypatch.delenv(k, raising=False)
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module('bench')
    monkeypatch.setattr(bench.ClockSampler, 'start', lambda self: None)
    monkeypatch.setattr(bench.ClockSampler, 'stop', lambda self: {'sm_mhz': 1965.0, 'sm_max_mhz': 1965.0, 'reasons': [], 'samples': 3})
    monkeypatch.setattr(sys, 'argv', ['bench.py'] + argv)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.main()
    lines = [l for l in buf.getvalue().splitlines() if l.startswith('{')]
    assert len(lines) == 1, buf.getvalue()
    return json.loads(lines[0])


def test_bench_line_carries_the_contract_keys(monkeypatch):
    line = _run(monkeypatch, ['--gpus', '1', '--steps', '4', '--warmup', '5', '--N', '64', '--m', '8', '--no-cpu-baseline',
                              '--equilibrate-seconds', '0.05', '--sustain-seconds', '0.01', '--schedule-steps', '10', '--bf-N', '12', '--ring-N', '256',
                              '--replicas-per-gpu', '2', '--bipartite-N', '32'])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data',
              'config', 'clocks', 'e2e', 'gpu_launches', 'roofline', 'transient', 'sustained', 'classic', 'secondary', 'comm'):
        assert k in line, k
    assert line['steps'] == 4 and line['warmup'] == 5 and line['n_gpus'] == 1 and line['vs_baseline'] is None
    assert line['ms_per_step'] == pytest.approx(7.0 / 4)                     # the stand-in clock advances 7 ms per event
    assert line['value'] == pytest.approx(64 * 8 * 4 / 7e-3)
    assert set(line['e2e']) >= {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'}
    assert line['e2e']['h2d_bytes_per_step'] == 64 * 8 and line['e2e']['d2h_bytes_per_step'] == 64 * 8 + 8 * 4
    assert set(line['roofline']) >= {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'}
    assert line['run']['equilibration_steps'] > 0 and 'workload' in line['config'] and line['schedule_sweep']['steps'] == 10
    import bench
    assert line['config'] == bench.workload_config(64, 8)          # nothing run-dependent in `config`: both arms print the same one
    assert line['transient']['steps'] == 4
    c1 = line['secondary']['c1_tutorial']
    assert 'error' not in c1 and c1['steps'] == 619 and c1['N'] == 128 and c1['m'] == 32
    import bench
    ref_c1 = bench.c1_reference_cpu()   # what the cpu_baseline leg adds (None without oracle/_ref): the compiled reference really runs the tutorial here
    if ref_c1 is not None:
        assert ref_c1['steps'] == 619 and ref_c1['E_min'] < -200 and ref_c1['kind'] == 'reference'
    for leg in ('calculate_E_c2', 'bipartite_c3'):
        assert 'error' not in line['secondary'][leg], line['secondary'][leg]
        assert line['secondary'][leg]['roofline']['bound'] == 'tensor' and line['secondary'][leg]['roofline']['frac'] > 0
    assert 'error' not in line['comm']['bf_n40_sharded'] and 'error' not in line['comm']['replicas_c5a'] and 'error' not in line['comm']['ring_c5b']


def test_quick_mode_and_reference_arm_keys(monkeypatch):
    line = _run(monkeypatch, ['--quick', '--steps', '3', '--N', '64', '--m', '8'])
    assert line['sustained'] is None and line['classic'] is None and line['comm'] is None and line['secondary'] is None and 'cpu_baseline' not in line
    import bench
    monkeypatch.setattr(bench, 'N_SPINS', 256)
    monkeypatch.setattr(bench, 'M_TROTTERS', 16)
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--impl', 'reference', '--steps', '2', '--warmup', '1'])
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.main()
    ref = json.loads([l for l in buf.getvalue().splitlines() if l.startswith('{')][0])
    assert ref['impl'] == 'reference' and ref['steps'] == 2 and ref['cpu_baseline']['kind'] in ('reference', 'port')
    assert ref['e2e'] == {'value': ref['value'], 'unit': 'attempts/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert ref['cpu_baseline']['cores'] >= 1 and ref['value'] > 0
    assert ref['config'] == bench.workload_config()


def test_cpu_baseline_leg_runs_inside_the_main_arm(monkeypatch):
    """the one leg of the GPU arm that executes oracle/ (permitted: cpu_baseline), at a reduced size: the C2 sample and config C1"""
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(bench, 'N_SPINS', 256)
    monkeypatch.setattr(bench, 'M_TROTTERS', 16)
    line = _run(monkeypatch, ['--steps', '3', '--N', '64', '--m', '8', '--equilibrate-seconds', '0.02', '--sustain-seconds', '0', '--schedule-steps', '0',
                              '--no-classic-leg', '--no-comm-legs', '--bipartite-N', '32'])
    cb = line['cpu_baseline']
    assert cb['value'] > 0 and cb['kind'] in ('reference', 'port') and cb['cores'] >= 1
    if cb['kind'] == 'reference':
        assert cb['c1_tutorial']['steps'] == 619 and cb['c1_tutorial']['E_min'] < -200
    assert line['secondary']['c1_tutorial']['steps'] == 619


def _worker():
    """one rank of the two-rank dry run: the same stand-ins, torch.distributed over gloo instead of NCCL"""
    import torch
    import torch.distributed as dist
    _install_fakes(_Patcher())
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend, **kw: real_init('gloo')
    sys.path.insert(0, ROOT)
    import bench
    bench.ClockSampler.start = lambda self: None
    bench.ClockSampler.stop = lambda self: {'sm_mhz': 1965.0, 'sm_max_mhz': 1965.0, 'reasons': [], 'samples': 3}
    sys.argv = ['bench.py', '--gpus', '2', '--steps', '4', '--warmup', '3', '--N', '64', '--m', '8', '--equilibrate-seconds', '0.05',
                '--sustain-seconds', '0.01', '--schedule-steps', '10', '--bf-N', '12', '--ring-N', '256', '--replicas-per-gpu', '2', '--bipartite-N', '32']
    bench.main()


def test_two_rank_run_does_not_deadlock():
    """every collective of bench.py is reached by both ranks the same number of times (world size 2, gloo, 127.0.0.1)"""
    import socket
    import subprocess
    with socket.socket() as so:
        so.bind(('127.0.0.1', 0))
        port = so.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), '--worker'], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=300))
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    assert [p.returncode for p in procs] == [0, 0], outs[0][1][-1500:] + outs[1][1][-1500:]
    lines = [l for l in outs[0][0].splitlines() if l.startswith('{')]
    assert len(lines) == 1 and not [l for l in outs[1][0].splitlines() if l.startswith('{')]      # rank 0 alone prints the line
    line = json.loads(lines[0])
    assert line['n_gpus'] == 2 and line['scaling'] == 'weak' and 'cpu_baseline' not in line
    assert line['value'] == pytest.approx(2 * 64 * 8 * 4 / 7e-3)                                 # whole-job aggregate over both ranks
    assert 'error' not in line['comm']['ring_c5b'] and line['comm']['ring_c5b']['m'] == 512


def _worker_hang():
    """single rank; the replica leg never comes back: the watchdog has to print the line without the comm legs and end the process"""
    import time
    _install_fakes(_Patcher())
    sys.modules['sqaod_b200.multigpu'].anneal_replicas = lambda *a, **k: time.sleep(3600)
    sys.path.insert(0, ROOT)
    import bench
    bench.ClockSampler.start = lambda self: None
    bench.ClockSampler.stop = lambda self: {'sm_mhz': 1965.0, 'sm_max_mhz': 1965.0, 'reasons': [], 'samples': 3}
    sys.argv = ['bench.py', '--steps', '4', '--warmup', '3', '--N', '64', '--m', '8', '--equilibrate-seconds', '0.05', '--sustain-seconds', '0.01',
                '--schedule-steps', '10', '--bf-N', '12', '--ring-N', '256', '--replicas-per-gpu', '2', '--bipartite-N', '32', '--no-cpu-baseline',
                '--leg-timeout', '2']
    bench.main()


def test_a_leg_that_hangs_cannot_cost_the_line():
    import subprocess
    env = {k: v for k, v in os.environ.items() if k not in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE')}
    out = subprocess.run([sys.executable, os.path.abspath(__file__), '--worker-hang'], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-1500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, out.stdout[-1500:]
    line = json.loads(lines[0])
    assert 'comm' in line['watchdog'] and line['comm'] is None
    assert line['value'] > 0 and line['e2e']['value'] > 0 and line['sustained']['steps'] >= 4 and line['secondary'] is not None


if __name__ == '__main__' and '--worker' in sys.argv:
    _worker()
if __name__ == '__main__' and '--worker-hang' in sys.argv:
    _worker_hang()
