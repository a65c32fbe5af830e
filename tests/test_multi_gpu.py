"""Multi-GPU parity (needs >= 2 GPUs; run with `gpurun --gpus 2 -- python -m pytest tests -m gpu`):
the partitioned forward + backward equals the single-GPU result."""
import json
import os
import subprocess
import sys

import pytest

from jaxsso_b200 import _native as nat
from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('world', [2, 4, 8])
def test_partitioned_equals_single(world):
    if nat.lib().jsso_device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
           '--master-addr', '127.0.0.1', '--master-port', str(29510 + world),
           os.path.join(ROOT, 'scripts', 'dist_check.py'), '48']
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith('DIST_CHECK')]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.loads(line[0].split(' ', 1)[1])
    assert res['u_err'] < 1e-8 and res['grad_err'] < 1e-7 and res['dprop_err'] < 1e-7


@pytest.mark.parametrize('p2p', ['', 'p2p', 'p2p setup'])
@pytest.mark.parametrize('world,min_dist', [(2, 500), (2, 100000), (4, 500), (8, 500)])
def test_distributed_multigrid_equals_single(world, min_dist, p2p):
    """Row-range distributed V-cycle PCG (jsso_mg_set_dist): same u and iteration count as the single-GPU
    multigrid solve; min_dist 500 distributes two levels at 96^2, 100000 only the fine one; 'p2p': exchanges over
    peer memory; 'setup': the numeric multigrid setup distributed as well (jsso_mg_set_dist_setup)."""
    if nat.lib().jsso_device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
           '--master-addr', '127.0.0.1', '--master-port', str(29530 + world),
           os.path.join(ROOT, 'scripts', 'dist_mg_check.py'), '96', str(min_dist), '1'] + p2p.split()
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith('DIST_MG_CHECK')]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.loads(line[0].split(' ', 1)[1])
    assert res['u_err_vs_single_mg'] < 1e-8 and res['same_on_all_ranks']
    assert abs(res['iters_dist'] - res['iters_single']) <= 2
