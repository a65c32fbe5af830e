"""The product's HOST DRIVER on the CPU: csrc/jsso_api.cu (handle management, assembly, scaled PCG, multigrid setup
and V-cycle PCG, the row-range distributed multigrid solve, adjoint) compiled by g++ on the SIMT emulator
(tests/emu), kernel launches rewritten into emulator launches, CUDA runtime = heap, NCCL = queues between rank
THREADS.  The package's own Python binding runs against it in a subprocess (JSSO_LIB=...).  This is test
infrastructure, not a fallback: libjsso.so is untouched and still refuses to run without a GPU
(tests/test_abi.py).  It verifies LOGIC (indexing, ordering, which exchange is needed where); speed and the
real memory model are the GPU tests' business."""
import json
import os
import subprocess
import sys

import pytest

from tests.conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))


@pytest.fixture(scope='module')
def emu_api():
    import build_emu
    return build_emu.build_api()


def run(emu_api, *args, env=None):
    e = dict(os.environ, JSSO_LIB=emu_api, **(env or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'emu', 'driver_check.py'), *map(str, args)],
                       capture_output=True, text=True, timeout=900, env=e, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith('EMU_RESULT ')]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads(line[0].split(' ', 1)[1])


def test_value_and_gradient_through_the_emulated_driver(emu_api):
    """jsso_value_and_grad_host end to end: Ke + assembly (warp tasks), BC, block-Jacobi scaling, 3-kernel CG with
    device-side stop flag, adjoint kernels, node gather -- against the oracle at the GPU tests' tolerances."""
    res = run(emu_api, 'grad', 6)
    assert res['u_err'] <= 1e-8 and res['c_err'] <= 1e-8
    assert res['g_err'] <= 1e-6 and res['dq_err'] <= 1e-6 and res['db_err'] <= 1e-6
    assert res['launches'] > 100
    # host buffers reported as pageable: staged through the handle's pinned buffers with the threaded memcpy
    # (default: reported as pinned, DMA straight from / into the caller's arrays) -- the same numbers
    pg = run(emu_api, 'grad', 6, env={'EMU_PAGEABLE': '1'})
    assert all(pg[k] == res[k] for k in ('u_err', 'c_err', 'g_err', 'dq_err', 'db_err', 'iterations'))


def test_spmv_and_values_after_a_solve(emu_api):
    """A solve scales the stored matrix in place (W K W^T); jsso_spmv and jsso_get_values_host must still return K."""
    res = run(emu_api, 'afterpcg', 5)
    assert res['K_err'] <= 1e-12 and res['y_err'] <= 1e-12


def test_chunked_host_pipeline_equals_unchunked(emu_api):
    """JSSO_E2E_CHUNKS=K (opt-in): u / lam uploaded in K node ranges, the quad adjoint in K quad ranges with offset
    pointers, d_prop_q returned range by range -- bitwise the same gradients as the single-launch path, and within
    1e-6 of the complex-step oracle.  (The emulator checks the range arithmetic; stream ordering is the GPU's.)"""
    import hashlib  # noqa: F401
    base = run(emu_api, 'e2e', 7, env={'JSSO_E2E_CHUNKS': '1', 'PYTHONHASHSEED': '0'})
    for k in ('3', '8'):
        res = run(emu_api, 'e2e', 7, env={'JSSO_E2E_CHUNKS': k, 'PYTHONHASHSEED': '0'})
        assert res['hash'] == base['hash'] and res['sum'] == base['sum']
        assert res['g_err'] <= 1e-6 and res['dq_err'] <= 1e-6 and res['db_err'] <= 1e-6


@pytest.mark.parametrize('fp16', ['0', '1'])
def test_multigrid_pcg_through_the_emulated_driver(emu_api, fp16):
    """mg_numeric_setup (power iteration, smoothed prolongator, Galerkin products, dense coarse inverse) + V-cycle
    PCG (fused iteration, device scalars); the fine-level V-cycle matrix is stored in binary16 by default,
    JSSO_MG_FP16=0 keeps FP32: same solution, iteration count within 2."""
    res = run(emu_api, 'mg', 12, 1, env={'JSSO_MG_FP16': fp16})
    assert res['mg_converged'] and res['bj_converged']
    assert res['mg_err'] <= 1e-8 and res['bj_err'] <= 1e-8
    assert res['mg_iters'] < res['bj_iters'] / 2
    if fp16 == '1':
        base = run(emu_api, 'mg', 12, 1, env={'JSSO_MG_FP16': '0'})
        assert abs(res['mg_iters'] - base['mg_iters']) <= 2


def test_blocked_dense_inverse_of_a_larger_coarsest_level(emu_api):
    """A coarsest level of more than 96 unknowns is inverted by a blocked Gauss-Jordan over all SMs (pivot blocks of 32
    rows, three kernels per block; a partial last block here: 36 nodes = 216 = 6 x 32 + 24 unknowns for max_coarse_nodes = 256):
    the same solve as with the one-CTA kernel."""
    blocked = run(emu_api, 'mg', 16, 1, env={'EMU_MAX_COARSE': '256'})
    single = run(emu_api, 'mg', 16, 1, env={'EMU_MAX_COARSE': '256', 'JSSO_MG_DENSE_SINGLE_MAX': '100000'})
    assert blocked['mg_converged'] and blocked['mg_err'] <= 1e-8
    assert blocked['mg_iters'] == single['mg_iters'] and abs(blocked['mg_err'] - single['mg_err']) <= 1e-10
    assert blocked['mg_launches'] > single['mg_launches'] + 10          # 7 pivot blocks x 3 kernels against 1


@pytest.mark.parametrize('world,size,min_dist,deg', [(2, 12, 10, 1), (2, 12, 1000, 2), (4, 16, 10, 2)])
def test_distributed_multigrid_through_the_emulated_driver(emu_api, world, size, min_dist, deg):
    """jsso_mg_set_dist + mg_solve_dist (the REAL C++ driver, not its Python replay) on `world` rank threads with
    the fake NCCL: same iteration count as the undistributed solve, bitwise identical u on every rank, u equal to
    the undistributed solve to rounding and to the oracle within 1e-8.  min_dist 10 distributes two levels (with an
    EMPTY row range on some ranks at the first replicated level), 1000 only the fine one."""
    res = run(emu_api, 'dist', world, size, min_dist, deg)
    assert res['converged'] and res['identical_on_all_ranks']
    assert all(abs(i - res['iters_single']) <= 1 for i in res['iters_dist'])
    assert res['err_vs_single'] <= 1e-10 and res['err_vs_oracle'] <= 1e-8
    assert res['plan']['n_dist'] == (2 if min_dist == 10 else 1)
    assert res['exchanges'] > 0 and res['allreduces'] >= 3 * res['iters_single']


@pytest.mark.parametrize('world,size,min_dist,deg,p2p', [(2, 12, 10, 1, ''), (4, 16, 10, 1, 'p2p'), (2, 12, 1000, 2, 'p2p'),
                                                       (8, 24, 10, 1, 'p2p')])
def test_distributed_numeric_setup(emu_api, world, size, min_dist, deg, p2p):
    """jsso_mg_set_dist_setup: every rank assembles / scales only the row hull it reads and computes only its share of
    the prolongators and Galerkin products (ghost rows recomputed, coarse matrices all-gathered).  Fresh "device"
    memory is poisoned (EMU_POISON: NaN / -1), so anything read but not computed would surface: same iteration count
    and solution as the undistributed solve, identical u on every rank -- and with ONE ghost block dropped from the
    plan (negative control) the solve must fail."""
    env = {'EMU_POISON': '1', 'EMU_DIST_SETUP': '1'}
    res = run(emu_api, 'dist', world, size, min_dist, deg, *([p2p] if p2p else []), env=env)
    assert res['converged'] and res['identical_on_all_ranks']
    assert all(abs(i - res['iters_single']) <= 1 for i in res['iters_dist'])
    assert res['err_vs_single'] <= 1e-10 and res['err_vs_oracle'] <= 1e-8
    if world == 2 and not p2p:
        e = dict(os.environ, JSSO_LIB=emu_api, EMU_DROP_GHOST='1', **env)
        r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'emu', 'driver_check.py'), 'dist', '2', '12', '10', '1'],
                           capture_output=True, text=True, timeout=900, env=e, cwd=ROOT)
        assert r.returncode != 0 or 'EMU_RESULT' not in r.stdout or '"converged": false' in r.stdout


@pytest.mark.parametrize('args,env', [(('grad', 4), {}), (('dist', 2, 6, 5, 1), {'JSSO_MG_FP16': '1', 'EMU_DIST_SETUP': '1'})])
def test_memcheck_under_address_sanitizer(args, env):
    """The emulated driver built with -fsanitize=address: every "device" buffer is a heap block, so an out-of-bounds
    access of a kernel (or of the host driver) aborts with a report naming the .cuh line -- compute-sanitizer
    memcheck for a box without a GPU.  Runs the gradient flow and the distributed multigrid solve with the binary16
    fine level (small meshes: the sanitizer costs 3-5x); a negative control (a gather index one row past the end)
    must be reported."""
    import build_emu
    rt = build_emu.asan_runtime()
    if rt is None:
        pytest.skip('libasan not available')
    lib = build_emu.build_api(asan=True)
    e = dict(os.environ, JSSO_LIB=lib, LD_PRELOAD=rt, ASAN_OPTIONS='detect_leaks=0:halt_on_error=1', **env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'emu', 'driver_check.py'), *map(str, args)],
                       capture_output=True, text=True, timeout=1200, env=e, cwd=ROOT)
    assert r.returncode == 0 and 'AddressSanitizer' not in r.stderr, r.stderr[-4000:]
    assert 'EMU_RESULT ' in r.stdout
    if args[0] == 'grad':
        neg = ("import sys; sys.path.insert(0, %r)\n"
               "import numpy as np\nfrom jaxsso_b200 import _native as nat\n"
               "s = nat.DeviceArray.from_host(np.arange(60.0)); i = nat.DeviceArray.from_host(np.array([0, 10], np.int32))\n"
               "nat.gather_rows(s, i, 6)\n" % ROOT)
        r2 = subprocess.run([sys.executable, '-c', neg], capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
        assert r2.returncode != 0 and 'heap-buffer-overflow' in r2.stderr and 'gather_rows_kernel' in r2.stderr


def test_bench_distributed_gradient_evaluation_on_four_rank_threads(emu_api):
    """The same leg on four rank threads (ranks with two neighbours, an empty range at the first replicated level),
    fresh "device" memory poisoned: the ghost-row recomputation of the corrected iterate reads only what the plan's
    exchanges delivered."""
    res = run(emu_api, 'benchleg', 4, 24, 10, 'natural', env={'EMU_POISON': '1'})
    assert res['leg']['pcg_iterations'] == res['iters_single'] and res['leg']['distributed']['n_dist'] == 2
    assert res['u_err_vs_single'] <= 1e-9 and res['u_err_vs_oracle'] <= 1e-8 and res['g_err'] <= 1e-6 and res['dq_err'] <= 1e-6


@pytest.mark.parametrize('kind', ['natural', 'rcb'])
def test_bench_distributed_gradient_evaluation_on_rank_threads(emu_api, kind):
    """bench.py's N > 1 gradient evaluation (`distributed_solve_handle`: whole-mesh handle, V-cycle PCG distributed by
    row ranges over peer memory incl. the all-gather of the first replicated level, gather to the partitioned handle,
    partitioned adjoint) on two rank threads.  'natural' cuts the plate into contiguous ranges of its own numbering
    (identical hierarchy: the iteration count is the single-GPU one); 'rcb' renumbers by owner and aggregates in
    the original order (`build_hierarchy_invariant`): within one iteration."""
    res = run(emu_api, 'benchleg', 2, 12, 10, kind)
    leg = res['leg']
    assert leg['partition'] == kind and leg['peer_memory']
    assert leg['distributed']['n_dist'] == 2 and leg['halo_exchanges'] > 0
    assert abs(leg['pcg_iterations'] - res['iters_single']) <= (0 if kind == 'natural' else 1)
    assert res['u_err_vs_single'] <= 1e-9 and res['u_err_vs_oracle'] <= 1e-8
    assert res['g_err'] <= 1e-6 and res['dq_err'] <= 1e-6


def test_device_scalar_pcg_poll_period(emu_api):
    """The PCG scalars live on the device; the host polls the residual every JSSO_MG_POLL iterations and the
    launches enqueued past convergence are no-ops: the iteration count and the solution do not depend on the poll
    period, on one GPU and on two rank threads."""
    base = run(emu_api, 'mg', 12, 1, env={'JSSO_MG_POLL': '1'})
    for k in ('4', '8'):
        res = run(emu_api, 'mg', 12, 1, env={'JSSO_MG_POLL': k})
        assert res['mg_converged'] and res['mg_err'] <= 1e-8
        assert res['mg_iters'] == base['mg_iters']
    d = run(emu_api, 'dist', 2, 12, 10, 1, env={'JSSO_MG_POLL': '3'})
    assert d['converged'] and d['identical_on_all_ranks'] and d['err_vs_oracle'] <= 1e-8
    assert all(i == d['iters_single'] for i in d['iters_dist'])


@pytest.mark.parametrize('p2p', ['', 'p2p'])
@pytest.mark.parametrize('world', [2, 4])
def test_partitioned_path_on_rank_threads(emu_api, world, p2p):
    """Calibration of the rank-thread emulation against a path that IS verified on hardware (tests/test_multi_gpu.py,
    scripts/dist_check.py): partitioned handles, `jsso_set_halo`, distributed block-Jacobi CG with NCCL halo
    exchanges and scalar all-reduces, partitioned adjoint -- u and gradients equal the oracle's, every rank
    takes the same number of iterations.  'p2p': the peer-memory variant (halo pushes into the neighbours' vectors,
    mailbox all-reduces, kernels spinning on flags another rank's kernel sets) -- the rank threads' launches run
    concurrently, "IPC" handles carry plain pointers."""
    res = run(emu_api, 'part', world, 8, p2p)
    assert res['u_err'] <= 1e-8 and res['g_err'] <= 1e-6 and res['dq_err'] <= 1e-6
    assert len(set(res['iterations'])) == 1


def test_full_size_property_tests_run_on_the_emulator(emu_api):
    """tests/test_zz_full_size_properties.py (GPU marker, 1024^2 on the B200) at size 10 against the emulated library:
    keeps the logic of those tests checked where there is no GPU."""
    e = dict(os.environ, JSSO_LIB=emu_api, JSSO_FULL_SIZE='10')
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(ROOT, 'tests', 'test_zz_full_size_properties.py'), '-q',
                        '-x', '-p', 'no:cacheprovider'], capture_output=True, text=True, timeout=900, env=e, cwd=ROOT)
    assert r.returncode == 0 and '2 passed' in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_graph_captured_vcycle(emu_api):
    """The coarse part of the V-cycle (levels >= 1; the whole V-cycle for Chebyshev degrees other than 1) is captured
    once per numeric setup and replayed as one graph launch (the emulator records the launches with their arguments
    by value and replays them); JSSO_MG_GRAPH=0 launches the kernels one by one: same iterations and solution, far
    fewer launches; on rank threads the replicated coarse levels are the captured part."""
    base = run(emu_api, 'mg', 12, 1, env={'JSSO_MG_GRAPH': '0'})
    g = run(emu_api, 'mg', 12, 1)
    assert g['mg_converged'] and g['mg_iters'] == base['mg_iters'] and g['mg_err'] <= 1e-8
    assert g['mg_launches'] <= base['mg_launches'] - 3 * g['mg_iters']      # the coarse levels (4 kernels + dense solve here) became one launch
    c = run(emu_api, 'mg', 12, 2, env={'JSSO_MG_POLL': '4', 'JSSO_MG_FP16': '0'})
    assert c['mg_converged'] and c['mg_err'] <= 1e-8
    d = run(emu_api, 'dist', 2, 12, 10, 1, env={'JSSO_MG_GRAPH': '0'})
    assert d['converged'] and d['identical_on_all_ranks'] and d['err_vs_oracle'] <= 1e-8
    assert all(i == d['iters_single'] for i in d['iters_dist'])


@pytest.mark.parametrize('world,size,min_dist,deg,env', [
    (2, 12, 10, 1, {}), (4, 12, 10, 1, {'EMU_JITTER': '3000'}),
    (4, 16, 10, 2, {'JSSO_MG_POLL': '3', 'EMU_JITTER': '1000'})])
def test_peer_memory_distributed_multigrid(emu_api, world, size, min_dist, deg, env):
    """jsso_mg_p2p_connect: halo exchanges as push + wait/unpack kernels over "peer memory" (stores into the peers'
    double-buffered receive arenas, release/acquire flags) and mailbox all-reduces, no NCCL on the iteration path.
    The rank threads' launches run concurrently, so the kernels really spin on flags set by another rank's kernel;
    EMU_JITTER pauses every rank at random before launches so that ranks run far ahead of each other (with a
    SINGLE-buffered arena this stress makes the solve fail -- checked by hand -- with the double buffer it stays
    bit-identical).  Same iterations as the undistributed solve, identical u on every rank."""
    res = run(emu_api, 'dist', world, size, min_dist, deg, 'p2p', env=env)
    assert res['peer_memory'] and res['converged'] and res['identical_on_all_ranks']
    assert all(abs(i - res['iters_single']) <= 1 for i in res['iters_dist'])
    assert res['err_vs_single'] <= 1e-10 and res['err_vs_oracle'] <= 1e-8
