"""CPU checks of bench.py's contract pieces that do not need a GPU: the reference arm's JSON line
(keys, e2e block, cpu_baseline incl. the literal gradient evaluation) and the clock sampler's
behaviour without NVML / nvidia-smi."""
import json
import os
import subprocess
import sys

from tests.conftest import ROOT


def test_reference_arm_line():
    cmd = [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
           '--ref-size', '6', '--ref-serial', '--ref-grad']
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d['impl'] == 'reference' and d['metric'] == 'MITC4 Ke+assembly+adjoint elements/s'
    assert d['unit'] == 'elements/s' and d['higher_is_better'] is True and d['dtype'] == 'f64'
    assert d['value'] > 0 and d['steps'] == 1 and d['vs_baseline'] is None
    assert d['e2e'] == {'value': d['value'], 'unit': 'elements/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] == 1 and cb['value'] == d['value']
    assert cb['grad_eval']['quads'] == 36 and cb['grad_eval']['seconds'] > 0
    assert [p['size'] for p in cb['grad_eval']['points']] == [6]
    assert 'sub-mesh of the same generator' in cb['sample'] and 'rate_at_64x64' in d['config']


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '0', '--ref-size', '4'], capture_output=True, text=True, timeout=120,
                       cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_clock_sampler_without_gpu_tools(monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setenv('PATH', '/nonexistent')          # no nvidia-smi; NVML has no driver in this container
    s = bench.ClockSampler(0)
    out = s.stop()
    assert set(out) >= {'sm_mhz', 'sm_max_mhz', 'samples', 'reasons'}
    assert out['samples'] == 0 or out['sm_mhz'] is not None
