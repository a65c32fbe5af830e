#!/usr/bin/env python
"""bench.py -- the hot-path benchmark (contract: see DESIGN.md section "Measurement").

Workload (N = 1): BASELINE.json configs[2], the synthetic 1024 x 1024 MITC4 shell
(1 048 576 quads, 6 303 750 dof), jittered as SURVEY.md 8(d).
One "step" = fused Ke + numeric assembly (BC imposed) + fused adjoint sensitivity
reduction over all quads, with coordinates, properties, u and lam resident in HBM:
metric M1 "MITC4 Ke+assembly+adjoint elements/s" (solve excluded by definition).
The full shape-gradient evaluation including the PCG solve (metric M2) is measured
once per run and reported under "grad_eval".

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference's algorithm on host cores
"""
from __future__ import annotations

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

from jaxsso_b200 import meshes  # noqa: E402

METRIC = 'MITC4 Ke+assembly+adjoint elements/s'
UNIT = 'elements/s'


# ------------------------------------------------------------------------------ helpers
def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """SM clock / throttle-reason samples during the timed region: NVML from a sampling thread
    (the same counters `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*`
    prints); falls back to an `nvidia-smi -lms` child writing to a file when NVML cannot be loaded."""
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index, period_s=0.05):
        self.sm, self.mx, self.reasons = [], [], set()
        self.period, self.stop_flag, self.source = period_s, threading.Event(), None
        self.proc = self.path = self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(',') if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.source = 'nvml'
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
        except Exception:
            self.nv = None
            try:
                self.path = os.path.join(ROOT, 'gpurun_out', f'clocks_{os.getpid()}.csv')
                os.makedirs(os.path.dirname(self.path), exist_ok=True)
                self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                              '--format=csv,noheader,nounits', '-lms', '100', '-f', self.path],
                                             stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                self.source = 'nvidia-smi'
            except Exception:
                self.proc = None

    def _poll_nvml(self):
        nv = self.nv
        bits = [(getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8), 'hw_slowdown'),
                (getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40), 'hw_thermal_slowdown'),
                (getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20), 'sw_thermal_slowdown'),
                (getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4), 'sw_power_cap')]
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mx.append(self.max_mhz)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for b, n in bits:
                    if r & b:
                        self.reasons.add(n)
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    @property
    def n_samples(self):
        return len(self.sm)

    def stop(self):
        if self.nv is not None:
            self.stop_flag.set()
            self.t.join(timeout=2.0)
        elif self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
            try:
                with open(self.path) as f:
                    rows = [l.strip() for l in f if l.strip()]
                os.remove(self.path)
            except Exception:
                rows = []
            for r in rows:
                f = [x.strip() for x in r.split(',')]
                try:
                    self.sm.append(float(f[0])); self.mx.append(float(f[1]))
                except (ValueError, IndexError):
                    continue
                for n, v in zip(self.NAMES, f[2:6]):
                    if v.lower().startswith('active'):
                        self.reasons.add(n)
        else:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'samples': 0, 'reasons': ['nvml and nvidia-smi unavailable']}
        return {'sm_mhz': float(np.median(self.sm)) if self.sm else None,
                'sm_max_mhz': max(self.mx) if self.mx else None, 'samples': len(self.sm),
                'reasons': sorted(self.reasons), 'source': self.source}


def synthetic_state(md):
    """A smooth displacement-like field u (zero at prescribed dofs) and lam = u/2."""
    rng = np.random.default_rng(7)
    x, y = md.crds[:, 0], md.crds[:, 1]
    sx, sy = x / max(x.max(), 1.0), y / max(y.max(), 1.0)
    u = np.zeros((md.n_node, 6))
    bump = np.sin(np.pi * sx) * np.sin(np.pi * sy)
    u[:, 2] = -1e-3 * bump
    u[:, 0] = 1e-5 * np.cos(np.pi * sx) * bump
    u[:, 1] = 1e-5 * np.cos(np.pi * sy) * bump
    u[:, 3] = 1e-4 * np.sin(np.pi * sx) * np.cos(np.pi * sy)
    u[:, 4] = -1e-4 * np.cos(np.pi * sx) * np.sin(np.pi * sy)
    u += 1e-7 * rng.standard_normal(u.shape)
    u = u.reshape(-1)
    u[md.known] = 0.0
    return u, 0.5 * u


# ------------------------------------------------------------------------------ reference arm
_REF = {}     # per worker process: the mesh arrays of the sample (set once by the pool initializer, not per job)


def _ref_init(crds, cnct, prop, u, lam, ndof):
    _REF.update(crds=crds, cnct=cnct, prop=prop, u=u, lam=lam, ndof=ndof)


def _oracle_chunk(rng):
    """Ke + COO->CSR (sum_duplicates) + adjoint element sensitivities of the quads [a, b) (worker process)."""
    import scipy.sparse as sp
    from oracle import jaxsso_oracle as orc
    a, b = rng
    crds, cnct, prop, u, lam = _REF['crds'], _REF['cnct'][a:b], _REF['prop'][a:b], _REF['u'], _REF['lam']
    n = cnct.shape[0]
    e = crds[cnct].reshape(n, 12)
    K = orc.element_K_quad(e, prop)
    r, c = orc.quad_indices(cnct)
    Kc = sp.coo_matrix((K.reshape(-1), (r, c)), shape=(_REF['ndof'], _REF['ndof'])).tocsr()
    dof = (6 * cnct.astype(np.int64)[:, :, None] + np.arange(6)[None, None, :]).reshape(n, 24)
    W = -lam[dof][:, :, None] * u[dof][:, None, :]
    dx = np.zeros((n, 12))
    dp = np.zeros((n, 5))
    h = 1e-30
    for k in range(12):
        ec = e.astype(complex); ec[:, k] += 1j * h
        dx[:, k] = np.sum(W * (orc.element_K_quad(ec, prop).imag / h), axis=(1, 2))
    for k in range(5):
        pc = prop.astype(complex); pc[:, k] += 1j * h
        dp[:, k] = np.sum(W * (orc.element_K_quad(e, pc).imag / h), axis=(1, 2))
    return Kc, dx, dp


def reference_step(md, pool, n_workers):
    """The reference's algorithm for the path on host cores: vmap(element_K_quad) -> raw COO -> sort /
    sum_duplicates (scipy tocsr per chunk, chunks added) -> element-wise adjoint reduction (complex-step dK_e)."""
    bounds = np.linspace(0, md.n_quad, n_workers + 1).astype(int)
    jobs = [(int(bounds[i]), int(bounds[i + 1])) for i in range(n_workers) if bounds[i + 1] > bounds[i]]
    res = list(pool.map(_oracle_chunk, jobs)) if pool else [_oracle_chunk(j) for j in jobs]
    Kg = res[0][0]
    for r_ in res[1:]:
        Kg = Kg + r_[0]
    d_crds = np.zeros((md.n_node, 3))
    dx = np.concatenate([r_[1] for r_ in res]).reshape(-1, 4, 3)
    np.add.at(d_crds, md.cnct_quads, dx)
    return Kg, d_crds


def _reference_rate(size, steps, warmup, n_workers):
    from concurrent.futures import ProcessPoolExecutor
    md = meshes.plate(size)
    u, lam = synthetic_state(md)
    init = (md.crds, md.cnct_quads, md.prop_quads, u, lam, md.ndof)
    pool = ProcessPoolExecutor(n_workers, initializer=_ref_init, initargs=init) if n_workers > 1 else None
    if pool is None:
        _ref_init(*init)
    try:
        for _ in range(warmup):
            reference_step(md, pool, n_workers)
        t0 = time.perf_counter()
        for _ in range(steps):
            reference_step(md, pool, n_workers)
        dt = (time.perf_counter() - t0) / steps
    finally:
        if pool:
            pool.shutdown()
    return md, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_workers = max(1, min(cores, 64))
    size = args.ref_size
    md, dt = _reference_rate(size, args.steps, args.warmup, n_workers)
    v = md.n_quad / dt
    # the rate is size independent (element-wise work): one step of the 64 x 64 sub-mesh of the same generator beside it
    rate64 = None
    if size != 64 and not args.ref_serial:
        md64, dt64 = _reference_rate(64, 1, 1, n_workers)
        rate64 = md64.n_quad / dt64
    # metric M2 on the host: full gradient evaluations of bounded samples in the reference's LITERAL formulation
    # (raw COO -> CSR, augmented Lagrange system, SuperLU twice, element derivatives), one process
    ge = None
    if args.ref_grad:
        from oracle import jaxsso_oracle as orc
        pts = []
        for gs in [int(x) for x in str(args.ref_grad_sizes or size).split(',') if x]:
            gmd = meshes.plate(gs)
            m = orc.Mesh(gmd.crds, gmd.cnct_quads, gmd.prop_quads, gmd.cnct_beams, gmd.prop_beams, gmd.known, gmd.loads)
            t0 = time.perf_counter()
            orc.value_and_grad(m, literal=True)
            pts.append({'size': gs, 'quads': int(gmd.n_quad), 'seconds': time.perf_counter() - t0})
        ge = {'seconds': pts[0]['seconds'], 'quads': pts[0]['quads'], 'points': pts,
              'what': 'oracle value_and_grad(literal=True): augmented SuperLU solve x2 + complex-step element '
                      'derivatives, one process; the points show how the CPU solve scales with the mesh'}
    sample = (f'{size}x{size} jittered plate ({md.n_quad} quads), a sub-mesh of the same generator as the {args.size}x{args.size} '
              f'workload, per step: NumPy/SciPy restatement of vmap(element_K_quad) + COO->CSR sum_duplicates + '
              f'complex-step adjoint reduction, {n_workers} worker processes (mesh arrays shared once per worker)')
    out = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus,
           'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
           'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
           'config': {'workload': f'synthetic {args.size}x{args.size} MITC4 shell plate (BASELINE configs[2]); the reference '
                                  f'arm (the oracle port: jax is not installable here) is timed on a bounded {size}x{size} '
                                  f'sub-mesh of the same generator; elements/s is size independent (rate_at_64x64 beside it)',
                      'rate_at_64x64': rate64},
           'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': n_workers, 'kind': 'port', 'sample': sample,
                            'grad_eval': ge},
           'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out))


def distributed_solve_handle(nat, gmd, rank, world, device, bcast, allgather, min_dist_nodes, use_p2p=True,
                             partition_kind='auto', dist_setup=True):
    """Whole-mesh handle whose multigrid V-cycle PCG is distributed by row ranges over the ranks (collective).

    The partition of the SOLVE is `dist_multigrid.solve_partition`: contiguous ranges of the mesh's own numbering when
    it is banded (the benchmark plate: no renumbering, so the hierarchy and the iteration count are exactly the
    single-GPU ones), else RCB + ownership renumbering with the aggregates still formed in the original order
    (`build_hierarchy_invariant`).  Exchanges and scalar all-reduces run over NVLink peer memory (CUDA IPC) unless
    use_p2p is false (NCCL send/recv).  Returns (handle, solve mesh, new->old node permutation or None, info, levels).
    (A function so that the CPU suite can run it on rank threads against the emulated library.)"""
    from jaxsso_b200 import dist_multigrid as dmg
    t0 = time.perf_counter()
    hd = nat.Handle(gmd.n_node, gmd.cnct_quads, gmd.cnct_beams, gmd.known, device=device)
    rp, ci = hd.pattern()
    if partition_kind == 'rcb':
        from jaxsso_b200 import partition
        perm, bounds = dmg.owner_permutation(partition.rcb_owner(gmd.crds[:, :2], world), world)
        kind = 'rcb'
    elif partition_kind == 'natural':
        perm, bounds, kind = None, dmg.natural_bounds(gmd.n_node, world), 'natural'
    else:
        perm, bounds, kind = dmg.solve_partition(gmd.crds, rp, ci, world)
    smd = gmd
    if kind == 'rcb':
        hd.close()
        smd = dmg.renumber_mesh(gmd, perm)
        hd = nat.Handle(smd.n_node, smd.cnct_quads, smd.cnct_beams, smd.known, device=device)
        rp, ci = hd.pattern()
        levels = hd.mg_setup(levels=dmg.build_hierarchy_invariant(rp, ci, perm))
    else:
        perm = None
        levels = hd.mg_setup()
    plan = dmg.build_plan(rp, ci, levels, bounds, min_dist_nodes=min_dist_nodes)
    hd.mg_set_dist(bcast(nat.nccl_unique_id() if rank == 0 else None), rank, world, plan)
    if dist_setup:
        hd.mg_set_dist_setup(rp, ci, levels, plan, rank)
    if use_p2p:
        hd.mg_p2p_connect(plan, allgather)
    info = {'nodes_per_level': [a for a, _ in hd.mg_levels] + [hd.mg_levels[-1][1]],
            'symbolic_setup_s': time.perf_counter() - t0, 'partition': kind,
            'numeric_setup': 'distributed by row ranges (ghost rows recomputed, coarse matrices all-gathered)' if dist_setup else 'replicated',
            'exchange': 'NVLink peer memory (push / wait kernels, mailbox all-reduce)' if use_p2p else 'NCCL send/recv',
            'distributed': dmg.plan_summary(plan)}
    return hd, smd, perm, info, levels


def batch_designs_eval(nat, n_designs, grid, rank, world, device, dist, barrier, max_over_ranks, rtol=1e-8):
    """BASELINE config 5: `n_designs` synthetic grid x grid beam-column gridshell designs, design k
    on rank k % world (one handle per GPU: same connectivity, the symbolic state and the multigrid hierarchy are built
    once); every design is one value + shape-gradient evaluation through the host-buffer entry point
    (jsso_value_and_grad_host: H2D, Ke + assembly, multigrid PCG, adjoint, D2H).  The compliances and the gradients
    (n_designs x n_node x 3) are all-gathered over NCCL at the end, inside the timed region."""
    md0 = meshes.gridshell(grid, 0)
    mine = list(range(rank, n_designs, world))
    t0 = time.perf_counter()
    h = nat.Handle(md0.n_node, md0.cnct_quads, md0.cnct_beams, md0.known, device=device)
    h.mg_setup()
    designs = [meshes.gridshell(grid, k) for k in mine]
    t_setup = time.perf_counter() - t0
    opts = nat.make_opts(rtol=rtol, precond='multigrid')
    if designs:
        h.value_and_grad_host(designs[0].crds, md0.prop_quads, md0.prop_beams, md0.loads, want=('crds',), opts=opts)   # warm-up
    n_max = (n_designs + world - 1) // world
    vals = np.zeros(n_max)
    grads = np.zeros((n_max, md0.n_node, 3))
    its = []
    barrier()
    t0 = time.perf_counter()
    for i, d in enumerate(designs):
        v, u, dc, _, _, fs, _ = h.value_and_grad_host(d.crds, md0.prop_quads, md0.prop_beams, md0.loads, want=('crds',), opts=opts)
        vals[i] = v
        grads[i] = dc
        its.append(int(fs.iterations))
    if dist is not None:
        import torch
        tv = torch.from_numpy(vals).cuda()
        tg = torch.from_numpy(grads).cuda()
        av = torch.empty((world,) + tuple(tv.shape), dtype=tv.dtype, device='cuda')
        ag = torch.empty((world,) + tuple(tg.shape), dtype=tg.dtype, device='cuda')
        dist.all_gather_into_tensor(av, tv)
        dist.all_gather_into_tensor(ag, tg)
        torch.cuda.synchronize()
        all_vals = av.cpu().numpy()            # [rank][i] = design rank + i * world
        g_bytes = int(ag.numel() * 8)
    else:
        all_vals, g_bytes = vals[None, :], 0
    dt = max_over_ranks(time.perf_counter() - t0)
    h.close()
    by_design = [float(all_vals[k % world][k // world]) for k in range(n_designs)]
    return {'designs': n_designs, 'designs_per_s': n_designs / dt, 'seconds': dt, 'ms_per_design_per_gpu': 1e3 * dt / max(len(mine), 1),
            'beams_per_design': int(md0.n_beam), 'dof_per_design': int(md0.ndof), 'pcg_iterations_first': its[:4],
            'compliance_first_last': [by_design[0], by_design[-1]], 'gathered_gradient_bytes': g_bytes,
            'setup_s': t_setup, 'sharding': f'design k on rank k % {world} (replicas; all-gather of compliances and gradients at the end)',
            'workload': f'BASELINE configs[4]: {n_designs} synthetic {grid}x{grid} gridshell designs'}


def pcg_iteration_bytes(nnzb0, levels):
    """Algorithmic HBM bytes of ONE fused multigrid-PCG iteration (Chebyshev-1), whole system, from the hierarchy's
    sizes.  Matrix blocks: 288 B (FP64), 144 B (FP32), 72 B (binary16) + 4 B column index; a vector pass = 48 B per
    node.  Fine level: q = A p in FP64 (nnzb*292 + n*100), two binary16 products (r0 = b - A b / theta,
    z = x + (b - A x) / theta: rowptr, gathered x, b, y), restriction and prolongation (FP32 P^T, P), direction
    update (3 passes) and x / r update (6 passes).  Coarse level l: two FP32 products with A_l, P_l^T, P_l, two
    smoother updates (3 + 4 passes incl. the 288-byte D^-1 block per node), b / x traffic."""
    n0 = levels[0]['n_f']
    fine = {'outer_spmv_fp64': nnzb0 * 292 + n0 * 100,
            'vcycle_fine_products_fp16': 2 * (nnzb0 * 76 + n0 * (4 + 3 * 48)),
            'restrict_prolong_fp32': 2 * levels[0]['nnz_p'] * 148 + (n0 + levels[0]['n_c']) * 2 * 48 + n0 * 48,
            'pcg_vector_updates': 9 * n0 * 48}
    coarse = 0
    for l in range(1, len(levels)):
        lv, nl = levels[l], levels[l]['n_f']
        nnz_a = levels[l - 1]['nnz_c']
        coarse += 2 * (nnz_a * 148 + nl * (4 + 3 * 48)) + 2 * lv['nnz_p'] * 148 + (nl + lv['n_c']) * 2 * 48 + nl * 48
        coarse += nl * (7 * 48 + 2 * 288)
    fine['coarse_levels'] = coarse
    fine['total'] = sum(fine.values())
    return fine


# ------------------------------------------------------------------------------ own arm
def run_b200(args):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        torch.cuda.set_device(local_rank)
        dist_.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        dist = dist_
    from jaxsso_b200 import _native as nat
    from jaxsso_b200 import build as jbuild
    from jaxsso_b200 import partition
    if rank == 0:
        jbuild.build()
    if dist:
        dist.barrier()
    L = nat.lib()
    L.jsso_set_device(local_rank)

    def barrier():
        L.jsso_stream_sync(None)
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if not dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    N = args.size
    t_setup = time.perf_counter()
    if args.scaling == 'weak' and world > 1:
        gmd = meshes.plate(N, M=N * world)
    else:
        gmd = meshes.plate(N)
    n_quad_total = gmd.n_quad
    if world > 1:
        owner = partition.rcb_owner(gmd.crds[:, :2], world)
        lm = partition.local_mesh(gmd, owner, rank, world)
        md, n_row = lm.md, lm.n_owned
        ug, lg = synthetic_state(gmd)
        u = ug.reshape(-1, 6)[lm.l2g].reshape(-1)
        lam = lg.reshape(-1, 6)[lm.l2g].reshape(-1)
    else:
        md, n_row = gmd, gmd.n_node
        u, lam = synthetic_state(md)
    h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=local_rank, n_row=n_row)
    if world > 1:
        idbuf = [nat.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(idbuf, src=0)
        h.set_halo(idbuf[0], rank, world, lm.peer_rank, lm.send_ptr, lm.send_idx, lm.recv_start, lm.recv_count)
        if args.p2p:
            # NVLink peer-memory path: halo pushes and scalar all-reduces inside the CG kernels
            hs = [None] * world
            dist.all_gather_object(hs, h.p2p_export())
            h.p2p_connect(hs, lm.remote_start)
    mg_info = None
    mg_levels = None
    hg = None
    smd, solve_perm = gmd, None     # the mesh the solve handle sees; new -> old node order if it is renumbered
    if world > 1 and args.solve and args.precond != 'block_jacobi':
        # N > 1: the solve is > 95 % of a gradient evaluation.  Every rank keeps a handle of the WHOLE mesh whose
        # multigrid V-cycle PCG is distributed by row ranges (jsso_mg_set_dist), halo exchanges and all-reduces over
        # NVLink peer memory; Ke+assembly+adjoint of the metric stay partitioned (RCB, ghost elements).
        # --replicated-solve: the same handle without the distribution (every rank solves redundantly).
        # --precond block_jacobi: the distributed CG over the RCB partition instead.
        def bcast(obj):
            box = [obj]
            dist.broadcast_object_list(box, src=0)
            return box[0]

        def allgather_obj(obj):
            box = [None] * world
            dist.all_gather_object(box, obj)
            return box

        if args.replicated_solve:
            t_mg = time.perf_counter()
            hg = nat.Handle(gmd.n_node, gmd.cnct_quads, gmd.cnct_beams, gmd.known, device=local_rank)
            mg_levels = hg.mg_setup()
            mg_info = {'nodes_per_level': [a for a, _ in hg.mg_levels] + [hg.mg_levels[-1][1]],
                       'symbolic_setup_s': time.perf_counter() - t_mg}
        else:
            hg, smd, solve_perm, mg_info, mg_levels = distributed_solve_handle(
                nat, gmd, rank, world, local_rank, bcast, allgather_obj, args.min_dist_nodes, use_p2p=args.p2p,
                partition_kind=args.solve_partition, dist_setup=args.dist_setup)
    if world == 1 and args.solve and args.precond != 'block_jacobi':
        # single GPU: smoothed-aggregation multigrid preconditioner (symbolic hierarchy, once per model)
        t_mg = time.perf_counter()
        mg_levels = h.mg_setup()
        mg_info = {'nodes_per_level': [a for a, _ in h.mg_levels] + [h.mg_levels[-1][1]],
                   'symbolic_setup_s': time.perf_counter() - t_mg}
    mg_bytes_it = None
    if mg_info is not None and args.cheb_degree == 1:
        mg_bytes_it = pcg_iteration_bytes(h.sizes.nnzb if world == 1 else hg.sizes.nnzb, mg_levels)
    if world == 1:
        mg_levels = None        # (at N > 1 the replicated check handle re-uses them)
    D = nat.DeviceArray
    crds_d, pq_d, pb_d = D.from_host(md.crds), D.from_host(md.prop_quads), D.from_host(md.prop_beams)
    u_d, lam_d, f_d = D.from_host(u), D.from_host(lam), D.from_host(md.loads)
    dc_d, dq_d = D((md.n_node, 3)), D((md.n_quad, 5))
    t_setup = time.perf_counter() - t_setup

    def step():
        h.assemble(crds_d, pq_d, pb_d, apply_bc=True)
        h.adjoint(crds_d, pq_d, pb_d, u_d, lam_d, dc_d, dq_d, None)

    ev = [L.jsso_event_create() for _ in range(4)]

    def timed(fn, reps):
        """CUDA events on the launching stream (the default stream of this process)."""
        L.jsso_event_record(ev[0], None)
        for _ in range(reps):
            fn()
        L.jsso_event_record(ev[1], None)
        import ctypes
        ms = ctypes.c_float()
        L.jsso_event_elapsed_ms(ev[0], ev[1], ctypes.byref(ms))
        return ms.value / reps

    sampler = ClockSampler(local_rank) if rank == 0 else None   # samples through warm-up + all timed legs
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    l0 = L.jsso_launch_count()
    ms_step = max_over_ranks(timed(step, args.steps))
    launches = L.jsso_launch_count() - l0
    barrier()
    # the stages on their own (roofline of each); the assembly's two kernels are timed separately with
    # CUDA events recorded between them on the launching stream (jsso_profile)
    ms_asm = max_over_ranks(timed(lambda: h.assemble(crds_d, pq_d, pb_d, apply_bc=True), args.steps))
    h.profile(True)
    g_ms, t_ms = [], []
    for _ in range(args.steps):
        h.assemble(crds_d, pq_d, pb_d, apply_bc=True)
        a, b = h.profile_read()
        g_ms.append(a); t_ms.append(b)
    h.profile(False)
    ms_geo = max_over_ranks(float(np.mean(g_ms)))
    ms_tasks = max_over_ranks(float(np.mean(t_ms)))
    ms_adj = max_over_ranks(timed(lambda: h.adjoint(crds_d, pq_d, pb_d, u_d, lam_d, dc_d, dq_d, None), args.steps))
    # the PCG's dominant kernel: block-CSR SpMV on the assembled (BC-imposed) matrix
    y_d = D((6 * n_row,))
    h.assemble(crds_d, pq_d, pb_d, apply_bc=True)
    for _ in range(3):
        h.spmv(u_d, y_d)
    ms_spmv = max_over_ranks(timed(lambda: h.spmv(u_d, y_d), max(args.steps, 20)))
    barrier()

    # end to end through the C ABI with host buffers (H2D + kernels + D2H inside the timed region)
    # host buffers are page-locked (cudaHostAlloc), as the contract asks: the H2D/D2H copies are DMA
    out_bufs = (nat.pinned_empty((md.n_node, 3)), nat.pinned_empty((md.n_quad, 5)), None)
    hc, hq, hb, hu, hl = (nat.pinned_copy(a) for a in (md.crds, md.prop_quads, md.prop_beams, u, lam))
    for _ in range(2):
        h.assemble_adjoint_host(hc, hq, hb, hu, hl, out_bufs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        h.assemble_adjoint_host(hc, hq, hb, hu, hl, out_bufs)
    barrier()
    s_e2e = max_over_ranks((time.perf_counter() - t0) / args.steps)
    if sampler and sampler.n_samples < 3:      # very short runs: keep the GPU busy until a few samples exist
        t_end = time.perf_counter() + 1.5
        while time.perf_counter() < t_end and sampler.n_samples < 3:
            step()
            L.jsso_stream_sync(None)
    clocks = sampler.stop() if sampler else None
    h2d = 8 * (md.crds.size + md.prop_quads.size + md.prop_beams.size + 2 * u.size)
    d2h = 8 * (out_bufs[0].size + out_bufs[1].size)
    # what PCIe allows for this step: pinned-memory copy bandwidth both ways, measured here (rank 0's device)
    pcie = None
    try:
        nb = 256 << 20
        hbuf = nat.pinned_empty((nb // 8,))
        dbuf = D((nb // 8,))
        bw = {}
        for name, fn in (('h2d', lambda: L.jsso_memcpy_h2d(dbuf.ptr, hbuf.ctypes.data, nb, None)),
                         ('d2h', lambda: L.jsso_memcpy_d2h(hbuf.ctypes.data, dbuf.ptr, nb, None))):
            fn(); L.jsso_stream_sync(None)
            t0 = time.perf_counter()
            for _ in range(4):
                fn()
            L.jsso_stream_sync(None)
            bw[name] = 4 * nb / (time.perf_counter() - t0) / 1e9
        pcie = {'h2d_gbs': bw['h2d'], 'd2h_gbs': bw['d2h'],
                'copy_bound_ms': 1e3 * max(h2d / (bw['h2d'] * 1e9), d2h / (bw['d2h'] * 1e9)),
                'copies_serial_ms': 1e3 * (h2d / (bw['h2d'] * 1e9) + d2h / (bw['d2h'] * 1e9)),
                'note': 'copy_bound_ms: both directions fully overlapped (full duplex); copies_serial_ms: upload then '
                        'download.  The gradients can only leave after the adjoint, which needs every upload.'}
        dbuf.free()
    except Exception as e:
        pcie = {'error': str(e)}

    # full shape-gradient evaluation incl. the solve (metric M2): one warm-up evaluation, then the mean of
    # --grad-evals timed ones (each: Ke+assembly, block-Jacobi scaling, numeric multigrid setup, PCG for u to the TRUE
    # relative residual rtol, lam = u/2, adjoint)
    grad_eval = None
    dist_fail = None
    if args.solve:
        precond = args.precond if (world == 1 or hg is not None) else 'block_jacobi'
        opts = nat.make_opts(rtol=args.rtol, maxiter=args.maxiter, check_every=100, compliance=True,
                             precond=precond, cheb_degree=args.cheb_degree)
        uu_d = D((md.ndof,))
        hs = hg if hg is not None else h           # the handle that solves
        distributed = hg is not None and not args.replicated_solve
        if hg is not None:   # whole-mesh solve handle (distributed or replicated), partitioned adjoint
            gc_d, gq_d, gb_d = D.from_host(smd.crds), D.from_host(smd.prop_quads), D.from_host(smd.prop_beams)
            gf_d, gu_d = D.from_host(smd.loads), D((smd.ndof,))
            if solve_perm is None:
                l2s = lm.l2g
            else:
                inv_perm = np.empty(gmd.n_node, np.int64)
                inv_perm[solve_perm] = np.arange(gmd.n_node)
                l2s = inv_perm[lm.l2g]
            l2g_d = D.from_host(l2s.astype(np.int32))
        else:
            gc_d, gq_d, gb_d, gf_d, gu_d = crds_d, pq_d, pb_d, f_d, uu_d

        def one_eval(o_):
            fs_ = hs.forward(gc_d, gq_d, gb_d, gf_d, gu_d, opts=o_)
            if hg is not None:
                nat.gather_rows(gu_d, l2g_d, 6, out=uu_d)
            h.backward(crds_d, pq_d, pb_d, uu_d, None, dc_d, dq_d, None, opts=o_)
            L.jsso_stream_sync(None)
            return fs_

        fs, ok, dts = None, True, []
        try:
            one_eval(opts)                       # warm-up (graph capture, lazy peer mappings, clocks)
            for _ in range(max(1, args.grad_evals)):
                barrier()
                t0 = time.perf_counter()
                fs = one_eval(opts)
                dts.append(max_over_ranks(time.perf_counter() - t0))
        except nat.JssoError as e:
            fs, ok = None, str(e)
        if fs is not None:
            dt = float(np.mean(dts))
            u_est = None
            # how accurate is u at this rtol?  The same system solved 100x tighter (or to the attainable accuracy of
            # FP64, whichever comes first): u_err_estimate = ||u(rtol) - u(tight)|| / ||u(tight)||
            if args.u_check and precond != 'block_jacobi':
                u_loose = gu_d.download()
                ot = nat.make_opts(rtol=args.rtol * 1e-2, maxiter=args.maxiter, compliance=True, precond=precond,
                                   cheb_degree=args.cheb_degree, use_x0=True)
                t0 = time.perf_counter()
                st_t = hs.pcg(gf_d, gu_d, opts=ot, allow_noconv=True)
                u_tight = gu_d.download()
                u_est = {
                    'value': float(np.linalg.norm(u_loose - u_tight) / np.linalg.norm(u_tight)),
                    'vs': f'the same system continued to rtol {args.rtol * 1e-2:g} or the attainable accuracy',
                    'tight_relres': st_t.relres, 'tight_extra_iterations': st_t.iterations,
                    'seconds': time.perf_counter() - t0, 'target': 1e-8}
            # stages of one evaluation (same call sequence, timed one by one)
            barrier()
            t0 = time.perf_counter()
            hs.assemble(gc_d, gq_d, gb_d, apply_bc=True)
            L.jsso_stream_sync(None)
            t_asm = max_over_ranks(time.perf_counter() - t0)
            o1 = nat.make_opts(rtol=args.rtol, maxiter=1, compliance=True, precond=precond, cheb_degree=args.cheb_degree)
            t0 = time.perf_counter()
            try:
                hs.pcg(gf_d, gu_d, opts=o1, allow_noconv=True)     # scaling + numeric multigrid setup + 1 iteration
            except nat.JssoError:
                pass
            L.jsso_stream_sync(None)
            t_setup1 = max_over_ranks(time.perf_counter() - t0)
            it = max(fs.iterations, 1)
            ms_it = 1e3 * max(dt - t_asm - t_setup1 - ms_adj * 1e-3, 0.0) / max(it - 1, 1)
            grad_eval = {'seconds': dt, 'evals_per_s': 1.0 / dt, 'evals_timed': len(dts), 'seconds_each': dts,
                         'pcg_iterations': fs.iterations, 'pcg_restarts': fs.restarts, 'true_relres': fs.relres,
                         'rtol': args.rtol,
                         'ms_per_pcg_iteration': ms_it,
                         'stage_s': {'assembly_whole_mesh': t_asm, 'scaling_numeric_mg_setup_first_iteration': t_setup1,
                                     'pcg_iterations': ms_it * 1e-3 * (it - 1), 'adjoint': ms_adj * 1e-3},
                         'preconditioner': (f'smoothed-aggregation multigrid (V-cycle, Chebyshev-{args.cheb_degree}; fine level '
                                            f'binary16, coarse levels FP32 storage, FP64 vectors and accumulation; fused iteration, '
                                            f'device-resident PCG scalars)'
                                            if precond != 'block_jacobi' else 'block-Jacobi'),
                         'solve': ('single GPU' if world == 1 else
                                   ('V-cycle PCG distributed by row ranges over ' + mg_info['exchange'] +
                                    (' (assembly + numeric setup distributed too)' if args.dist_setup else ' (assembly + numeric setup replicated)') + ', adjoint partitioned') if distributed else
                                   'replicated on every rank (whole-mesh handle), adjoint partitioned' if hg is not None
                                   else 'distributed CG over the partition'),
                         'multigrid': mg_info,
                         'u_err_estimate': u_est,
                         'note': 'Ke+assembly, PCG for u (numeric multigrid setup included), lam = u/2 '
                                 '(compliance), adjoint; mean of the timed evaluations after one warm-up evaluation'}
            if distributed:
                grad_eval['halo_exchanges'], grad_eval['scalar_allreduces'] = hg.mg_dist_counters()
            # N > 1: the distributed solve against the SAME handle logic without the distribution (replicated solve on
            # rank 0's device only would need another handle; every rank solves redundantly here) -- fail loudly
            if distributed and args.dist_check:
                hr = nat.Handle(smd.n_node, smd.cnct_quads, smd.cnct_beams, smd.known, device=local_rank)
                hr.mg_setup(levels=mg_levels)
                ur_d = D((smd.ndof,))
                fr = hr.forward(gc_d, gq_d, gb_d, gf_d, ur_d, opts=opts)
                hs.forward(gc_d, gq_d, gb_d, gf_d, gu_d, opts=opts)
                u_rep, u_dst = ur_d.download(), gu_d.download()
                diff = max_over_ranks(float(np.linalg.norm(u_dst - u_rep) / max(np.linalg.norm(u_rep), 1e-300)))
                grad_eval['u_rel_diff_vs_replicated_solve'] = diff
                grad_eval['replicated_pcg_iterations'] = fr.iterations
                hr.close()
                if not (diff <= 1e-8):
                    dist_fail = f'distributed solve differs from the replicated solve: {diff:.3e} > 1e-8'
        else:
            grad_eval = {'error': ok, 'seconds': float(np.mean(dts)) if dts else None}
            dist_fail = ok if world > 1 else None

    # FP64 peak of this device, measured in this run (DFMA chains, jsso_fp64_peak) -- the denominator of the FP64-bound
    # roofline; nominal 148 SM x 64 DFMA/clk x 1.965 GHz = 37.2 TFLOP/s if the measurement is skipped
    fp64_peak, fp64_src = 37.2, 'nominal B200 FP64 (148 SM x 64 DFMA/clk x 1.965 GHz)'
    if args.fp64_peak_seconds > 0:
        try:
            tf, sec = nat.fp64_peak(local_rank, args.fp64_peak_seconds)
            fp64_peak, fp64_src = tf, f'measured in this run: DFMA-chain microbenchmark (jsso_fp64_peak), {sec:.2f} s sustained'
        except Exception:
            pass

    # metric M2 end to end through the host-buffer entry point (N = 1): H2D of coordinates / properties / loads, the
    # whole gradient evaluation, D2H of u and the gradients inside the timed region
    m2_e2e = None
    if world == 1 and grad_eval and 'error' not in grad_eval and args.precond != 'block_jacobi':
        try:
            o_ = nat.make_opts(rtol=args.rtol, maxiter=args.maxiter, precond=args.precond, cheb_degree=args.cheb_degree)
            # page-locked host buffers on both sides, as for the M1 leg: the copies are DMA from / into the caller's arrays
            hin = [nat.pinned_copy(a) for a in (md.crds, md.prop_quads, md.prop_beams, md.loads)]
            hout = (nat.pinned_empty((md.ndof,)), nat.pinned_empty((md.n_node, 3)), nat.pinned_empty((md.n_quad, 5)), None)
            h.value_and_grad_host(*hin, want=('crds', 'prop_q'), opts=o_, out=hout)
            t0 = time.perf_counter()
            v_, u_, dc_, dq_, _, fs_, _ = h.value_and_grad_host(*hin, want=('crds', 'prop_q'), opts=o_, out=hout)
            dt_ = time.perf_counter() - t0
            # the same call with ordinary (pageable) NumPy arrays and freshly allocated results: staged by the library
            t0 = time.perf_counter()
            h.value_and_grad_host(md.crds, md.prop_quads, md.prop_beams, md.loads, want=('crds', 'prop_q'), opts=o_)
            dt_pageable = time.perf_counter() - t0
            m2_e2e = {'seconds': dt_, 'evals_per_s': 1.0 / dt_, 'call': 'jsso_value_and_grad_host (C ABI, host buffers)',
                      'host_buffers': 'page-locked', 'seconds_pageable_buffers': dt_pageable,
                      'h2d_bytes': int(8 * (md.crds.size + md.prop_quads.size + md.prop_beams.size + md.loads.size)),
                      'd2h_bytes': int(8 * (u_.size + dc_.size + dq_.size)), 'pcg_iterations': int(fs_.iterations),
                      'compliance': float(v_)}
        except nat.JssoError as e:
            m2_e2e = {'error': str(e)}

    # BASELINE config 5: a batch of independent gridshell designs (beam-columns) sharded over the ranks -- replicas,
    # no collective on the hot path; the 64 compliances and shape gradients are all-gathered at the end
    batch_eval = None
    if args.batch_designs > 0:
        try:
            batch_eval = batch_designs_eval(nat, args.batch_designs, args.batch_grid, rank, world, local_rank, dist,
                                            barrier, max_over_ranks, rtol=args.rtol)
        except nat.JssoError as e:
            batch_eval = {'error': str(e)}

    # BASELINE config 4 (N = 1 only): simultaneous shape + topology optimisation loop on a 512^2 shell through the
    # public host-buffer API (filters both sides, warm-started multigrid PCG); scripts/topo_shape_512.py
    topo_eval = None
    if world == 1 and args.topo_iters > 0:
        try:
            sys.path.insert(0, os.path.join(ROOT, 'scripts'))
            import topo_shape_512 as ts
            o_ = ts.run(args.topo_size, args.topo_iters, rtol=1e-6, device=local_rank)
            hist_, its_ = o_.pop('history'), o_.pop('pcg_iterations')
            topo_eval = dict(o_, strain_energy_first_last=[hist_[0], hist_[-1]], pcg_iterations_first_last=[its_[0], its_[-1]],
                             monotone_fraction=float(np.mean(np.diff(hist_) <= 0)) if len(hist_) > 1 else None,
                             workload=f'BASELINE configs[3] at {args.topo_size}^2, {args.topo_iters} of its 100 iterations',
                             note='the objective rises while the uniform starting density separates (first ~10 iterations) and '
                                  'then falls: 6.54e5 -> 2.03e5 over the full 100 iterations in 12.1 s '
                                  '(profiles/r2_topo100_512.json, scripts/topo_shape_512.py 512 100)')
        except Exception as e:
            topo_eval = {'error': f'{type(e).__name__}: {e}'}

    def build_out():
        hbm, peak_src = peaks()
        s = h.sizes
        # measured DRAM traffic per launch of the committed ncu capture of this exact workload (profiles/)
        ncu = {}
        try:
            with open(os.path.join(ROOT, 'profiles', 'ncu_traffic_1024.json')) as f:
                ncu = json.load(f) if (world == 1 and N == 1024) else {}
        except Exception:
            ncu = {}
        # dominant kernel: assemble_tasks_kernel.  Algorithmic bytes of that launch: every geometry record read
        # once (496 B per quad), task/item descriptors, every stored block written once
        REC_BYTES = 496
        tasks_bytes = s.n_quad * REC_BYTES + s.n_items * 2 + s.nnzb * (288 + 6)
        if ms_tasks > 0:
            roof = {'kernel': 'assemble_tasks_kernel', 'bound': 'hbm', 'achieved': tasks_bytes / (ms_tasks * 1e-3) / 1e9,
                    'peak': hbm, 'unit': 'GB/s', 'frac': tasks_bytes / (ms_tasks * 1e-3) / 1e9 / hbm,
                    'traffic': ncu.get('assemble_tasks_kernel'), 'peak_source': peak_src,
                    'algorithmic_bytes_per_launch': tasks_bytes, 'ms': ms_tasks,
                    'note': 'second bound: LSU data-pipe wavefronts (shared memory), see profiles/'}
        else:   # chunked single-kernel path (meshes the warp tasks cannot hold, or JSSO_ASM_CHUNKED=1)
            b_ = s.n_quad * (16 + 40) + s.n_node * 24 + s.nnzb * 288
            roof = {'kernel': 'assemble_fused_kernel', 'bound': 'hbm', 'achieved': b_ / (ms_asm * 1e-3) / 1e9,
                    'peak': hbm, 'unit': 'GB/s', 'frac': b_ / (ms_asm * 1e-3) / 1e9 / hbm, 'traffic': None,
                    'peak_source': peak_src, 'algorithmic_bytes_per_launch': b_, 'ms': ms_asm}
        # the whole assembly stage (quad_geometry_kernel + assemble_tasks_kernel) against SURVEY 8(d)'s
        # algorithmic bytes of a fused Ke+assembly: connectivity + properties + coordinates read once, every
        # stored block written once (the geometry records are overhead traffic of the two-kernel design)
        asm_bytes = s.n_quad * (16 + 40) + s.n_node * 24 + s.nnzb * 288
        tr = [ncu.get('quad_geometry_kernel'), ncu.get('assemble_tasks_kernel')]
        roof_asm = {'kernel': 'quad_geometry_kernel + assemble_tasks_kernel', 'bound': 'hbm',
                    'achieved': asm_bytes / (ms_asm * 1e-3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
                    'frac': asm_bytes / (ms_asm * 1e-3) / 1e9 / hbm,
                    'traffic': (tr[0] + tr[1]) if all(t is not None for t in tr) else None, 'peak_source': peak_src,
                    'algorithmic_bytes_per_launch': asm_bytes, 'ms': ms_asm, 'ms_geometry': ms_geo, 'ms_tasks': ms_tasks}
        spmv_bytes = s.nnzb * 292 + s.n_row * 100
        roof_spmv = {'kernel': 'bsr_spmv_kernel', 'bound': 'hbm', 'achieved': spmv_bytes / (ms_spmv * 1e-3) / 1e9,
                     'peak': hbm, 'unit': 'GB/s', 'frac': spmv_bytes / (ms_spmv * 1e-3) / 1e9 / hbm,
                     'traffic': ncu.get('bsr_spmv_kernel'),
                     'algorithmic_bytes_per_launch': spmv_bytes, 'ms': ms_spmv, 'peak_source': peak_src}
        adj_flops = 14000.0 * s.n_quad
        roof_adj = {'kernel': 'quad_adjoint_kernel(+node_gather)', 'bound': 'fp64', 'achieved': adj_flops / (ms_adj * 1e-3) / 1e12,
                    'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': adj_flops / (ms_adj * 1e-3) / 1e12 / fp64_peak,
                    'peak_source': fp64_src, 'ms': ms_adj,
                    'traffic': ncu.get('quad_adjoint_kernel'),
                    'algorithmic_flops_per_launch': adj_flops}
        value = n_quad_total / (ms_step * 1e-3)
        out = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
               'warmup': max(args.warmup, 3), 'ms_per_step': ms_step, 'higher_is_better': True,
               'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
               'config': {'workload': f'synthetic {N}x{N * world if args.scaling == "weak" else N} MITC4 shell plate '
                                      f'(BASELINE configs[2]), jittered; {n_quad_total} quads, '
                                      f'{6 * gmd.n_node} dof; step = Ke+assembly (geometry records + warp tasks) + adjoint reduction',
                          'parallelism': ((f'rcb{world} element work (ghost elements, no collective) + ' +
                                           ('distributed CG, ' if (hg is None and args.solve) else
                                            'replicated multigrid solve, ' if args.replicated_solve else
                                            'row-range distributed multigrid solve, ' if args.solve else '') +
                                           ('NVLink peer memory' if args.p2p else 'NCCL')) if world > 1 else 'single'),
                          'l2_note': 'working set per step (2.7 GB of block-CSR values at N=1) is larger than the 126 MB L2',
                          'rank0_local': {'n_quad': s.n_quad, 'n_node': s.n_node, 'n_row': s.n_row, 'nnzb': s.nnzb}},
               'e2e': {'value': n_quad_total / s_e2e, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                       'ms_per_step': s_e2e * 1e3, 'call': 'jsso_assemble_adjoint_host (C ABI, host buffers)', 'pcie': pcie},
               'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roof, 'roofline_assembly': roof_asm,
               'roofline_adjoint': roof_adj,
               'roofline_spmv': roof_spmv,
               'kernel_ms': {'assembly': ms_asm, 'quad_geometry': ms_geo, 'assemble_tasks': ms_tasks, 'adjoint': ms_adj,
                             'spmv': ms_spmv}, 'setup_s': t_setup,
               'grad_eval': grad_eval, 'batch_eval': batch_eval, 'topo_eval': topo_eval}
        if grad_eval and 'error' not in grad_eval:
            # metric M2 (the second half of BASELINE.json's metric) as a first-class entry of the line
            out['m2'] = {'metric': 'full shape-gradient evaluations/s (Ke, assembly, solve, adjoint) at '
                                   f'{n_quad_total} quads', 'value': grad_eval['evals_per_s'], 'unit': 'evals/s',
                         'seconds_per_eval': grad_eval['seconds'], 'n_gpus': world, 'rtol': args.rtol,
                         'pcg_iterations': grad_eval['pcg_iterations'],
                         'ms_per_pcg_iteration': grad_eval['ms_per_pcg_iteration'],
                         'u_err_estimate': (grad_eval.get('u_err_estimate') or {}).get('value')}
            if m2_e2e is not None:
                out['m2']['e2e'] = m2_e2e
            if mg_bytes_it:
                bi = mg_bytes_it['total'] / max(world, 1)     # row ranges: every rank streams 1/N of every level it distributes
                ms_it = grad_eval['ms_per_pcg_iteration']
                out['roofline_pcg_iteration'] = {
                    'kernel': 'one multigrid-PCG iteration (4 fine-level V-cycle products, coarse levels, direction, q = A p, update)',
                    'bound': 'hbm', 'achieved': bi / (ms_it * 1e-3) / 1e9 if ms_it > 0 else None, 'peak': hbm, 'unit': 'GB/s',
                    'frac': (bi / (ms_it * 1e-3) / 1e9 / hbm) if ms_it > 0 else None, 'peak_source': peak_src,
                    'algorithmic_bytes_per_iteration_per_gpu': bi, 'breakdown_bytes_whole_system': mg_bytes_it, 'ms': ms_it,
                    'traffic': ncu.get('pcg_iteration')}
        if world == 1 and args.cpu_baseline:
            cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                   '--ref-size', str(args.cpu_baseline_size), '--ref-serial', '--ref-grad', '--ref-grad-sizes', '64,128',
                   '--size', str(args.size)]
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
                ref = json.loads(r.stdout.strip().splitlines()[-1])
                out['cpu_baseline'] = ref['cpu_baseline']
            except Exception as e:  # the baseline is a report, never a reason to lose the bench line
                out['cpu_baseline'] = {'value': None, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'sample': f'failed: {e}'}
        return out

    out = build_out() if rank == 0 else None

    if dist_fail:
        if rank == 0:
            out['error'] = dist_fail
            print(json.dumps(out), flush=True)
        if dist:
            dist.destroy_process_group()
        sys.exit(1)
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    print(json.dumps(out))
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--size', type=int, default=1024, help='plate is size x size quads')
    ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'])
    ap.add_argument('--no-solve', dest='solve', action='store_false')
    ap.add_argument('--no-p2p', dest='p2p', action='store_false',
                    help='distributed CG over NCCL send/recv + all-reduce instead of peer-memory kernels')
    ap.add_argument('--replicated-solve', action='store_true',
                    help='N > 1: every rank solves the whole system redundantly instead of the row-range distributed '
                         'V-cycle PCG')
    ap.add_argument('--solve-partition', default='auto', choices=['auto', 'natural', 'rcb'],
                    help='N > 1: row partition of the distributed solve (auto: contiguous ranges of the mesh numbering '
                         'when it is banded, else RCB with renumbering)')
    ap.add_argument('--replicated-setup', dest='dist_setup', action='store_false',
                    help='N > 1: every rank assembles the whole matrix and builds the whole multigrid hierarchy (A/B)')
    ap.add_argument('--no-dist-check', dest='dist_check', action='store_false',
                    help='N > 1: skip the comparison of the distributed solve with a replicated solve of the same system')
    ap.add_argument('--topo-iters', type=int, default=10, help='BASELINE config 4 leg at N = 1: optimiser iterations (0: skip)')
    ap.add_argument('--topo-size', type=int, default=512)
    ap.add_argument('--batch-designs', type=int, default=64,
                    help='BASELINE config 5 leg: gridshell designs sharded over the ranks (0: skip)')
    ap.add_argument('--batch-grid', type=int, default=224, help='nodes per side of a gridshell design')
    ap.add_argument('--fp64-peak-seconds', type=float, default=1.0,
                    help='duration of the FP64 DFMA microbenchmark behind roofline_adjoint.peak (0: nominal peak)')
    ap.add_argument('--grad-evals', type=int, default=3, help='timed full gradient evaluations (after one warm-up)')
    ap.add_argument('--no-u-check', dest='u_check', action='store_false',
                    help='skip the tighter solve that estimates the error of u at --rtol')
    ap.add_argument('--min-dist-nodes', type=int, default=20000,
                    help='multigrid levels with fewer nodes run replicated in the distributed solve')
    ap.add_argument('--rtol', type=float, default=1e-8)
    ap.add_argument('--precond', default='auto', choices=['auto', 'block_jacobi', 'multigrid'])
    ap.add_argument('--maxiter', type=int, default=400000)
    ap.add_argument('--cheb-degree', type=int, default=1, help='Chebyshev smoother degree of the V-cycle')
    ap.add_argument('--no-cpu-baseline', dest='cpu_baseline', action='store_false')
    ap.add_argument('--ref-size', type=int, default=256, help='plate size of the bounded CPU sample (reference arm)')
    ap.add_argument('--cpu-baseline-size', type=int, default=64, help='plate size of the one-core cpu_baseline leg')
    ap.add_argument('--ref-grad-sizes', default=None,
                    help='comma-separated plate sizes of the literal CPU gradient evaluations (--ref-grad; default: --ref-size)')
    ap.add_argument('--ref-serial', action='store_true', help='reference arm on one core (cpu_baseline leg)')
    ap.add_argument('--ref-grad', action='store_true', help='also time one literal gradient evaluation of the sample')
    args = ap.parse_args()
    if args.impl == 'reference':
        if args.ref_serial:
            os.cpu_count = lambda: 1
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
