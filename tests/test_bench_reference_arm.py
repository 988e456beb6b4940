"""bench.py --impl reference is a CPU-only arm (the oracle / its C port timed on the host): it must run without a GPU and
print one JSON line with the contract's keys.  (The GPU arm is exercised by the driver on a B200.)"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {'impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
        'dtype', 'data', 'config', 'cpu_baseline', 'e2e'}


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


@pytest.mark.parametrize('workload,metric,unit', [('pcor1024', 'bicgstab_mcell_iters_per_s', 'Mcell-iters/s'),
                                                  ('karman320', 'simple_outer_iters_per_s', 'outer-iters/s')])
def test_reference_arm_line(workload, metric, unit):
    d = run(['--impl', 'reference', '--workload', workload, '--steps', '1', '--warmup', '1'])
    assert KEYS <= set(d), KEYS - set(d)
    assert d['impl'] == 'reference' and d['metric'] == metric and d['unit'] == unit
    assert d['value'] > 0 and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['config']['workload'] == workload
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_ignores_torchrun_thread_cap():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the arm must still use every host core it may run on."""
    d = run(['--impl', 'reference', '--workload', 'pcor1024', '--steps', '1', '--warmup', '1'], env={'OMP_NUM_THREADS': '1'})
    want = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    assert d['cpu_baseline']['cores'] == want


def test_reference_arm_other_ranks_stay_silent():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'pcor1024', '--gpus', '2'],
                         capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK='1', WORLD_SIZE='2'), cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_roofline_traffic_comes_from_the_committed_ncu_captures():
    """bench.py never carries a traffic constant: roofline.traffic is looked up, by kernel-name prefix, in the ncu launch lists
    committed under profiles/ (with the commit they were taken at) and scaled by the run's plane words."""
    sys.path.insert(0, ROOT)
    import bench
    docs = bench.measured_traffic()
    assert docs and all(d.get('commit') and d.get('kernels') for d in docs)
    words = 32770 * 32832
    for solver in ('bicgstab', 'cg'):
        for i, (name, desc, alg_words, pattern) in bench.kernel_classes(solver, fused=False, store_s=False).items():
            if solver == 'cg':
                continue                                      # the CG kernels are not in the default command's capture
            t, src = bench.traffic_of(docs, pattern, words)
            assert t is not None and 'profiles/' in src, (name, pattern)
            assert 0.9 <= t / (alg_words * 8 * words) <= 1.1, (name, t / (alg_words * 8 * words))   # traffic ~ algorithmic bytes
    # config 4: the Jacobi-plane kernels come from the second capture (9 words per plane word)
    for pattern in ('k_stencil<double, 0, 1, 1', 'k_stencil<double, 1, 1, 3', 'k_bicg_xr<double, 1'):
        t, src = bench.traffic_of(docs, pattern, 16386 * 3328)
        assert t is not None and 'karman' in src
        assert 0.9 <= t / (9 * 8 * 16386 * 3328) <= 1.1
    assert bench.traffic_of(docs, 'k_no_such_kernel', words) == (None, None)


def test_ncu_summary_leaves_out_noop_launches(tmp_path):
    """tools/ncu_summary.py on the committed config-4 launch list: launches behind the solver's done flag (a few microseconds, no
    bytes) must not dilute the per-launch traffic -- 9 words x 8 B per plane word for the Jacobi-plane kernels."""
    out = tmp_path / 't.json'
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_summary.py'), os.path.join(ROOT, 'profiles', 'r02_launches_karman16384.csv'),
                          '--words', str(16386 * 3328), '--json', str(out), '--commit', 'test', '--command', 'test'],
                         capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    assert 'no-ops behind the done flag' in res.stdout
    doc = json.load(open(out))
    k = [v for n, v in doc['kernels'].items() if n.startswith('k_stencil<double, 1, 1, 3')][0]
    assert 68.0 <= k['dram_bytes_per_word'] <= 76.0          # 72 algorithmic; 63.8 with the no-ops averaged in
