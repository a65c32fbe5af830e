"""BASELINE config 4: simultaneous shape & topology optimisation (nodal z + element density, E = mu^7 E0)
on a synthetic N x N MITC4 shell (default 512), 100 iterations of projected gradient descent through the
hot path: filters (sparse hat filter, both sides) -> Ke + assembly -> multigrid PCG warm-started from the
previous iterate -> adjoint sensitivities.  The optimiser is the host loop the reference's examples run
(Examples/shells_topo_shape.ipynb: filter radius, SIMP exponent 7, mu0); only `fun(x) -> (value, grad)`
is the hot path.   python scripts/topo_shape_512.py [N] [ITERS]
`run()` is what tests/test_large_parity.py and bench.py (--config4) call."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxsso_b200 import _native as nat, meshes
from jaxsso_b200.filters import HatFilter


def run(N=512, iters=100, mu0=None, dz0=None, rtol=1e-6, keep_first=False, radius=2.5, device=0):
    md = meshes.plate(N)
    t0 = time.perf_counter()
    h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=device)
    h.mg_setup()
    t_setup = time.perf_counter() - t0
    E0, p_simp, mu_min = float(md.prop_quads[0, 1]), 7.0, 0.1
    centroids = md.crds[md.cnct_quads].mean(1)
    t0 = time.perf_counter()
    Fz = HatFilter(md.crds[:, :2], radius, device=device)        # shape filter over nodes
    Fm = HatFilter(centroids[:, :2], radius, device=device)      # density filter over quads
    t_filter = time.perf_counter() - t0
    z0 = md.crds[:, 2].copy()
    sup = md.known[md.known % 6 == 2] // 6                       # supported nodes keep their height
    dz = np.zeros(md.n_node) if dz0 is None else np.array(dz0, float)   # design offsets of z
    mu = np.full(md.n_quad, 0.5) if mu0 is None else np.array(mu0, float)
    mean_mu = float(mu.mean())
    u_prev, hist, its = None, [], []
    first = {}
    step_z, step_mu = 0.05 * N * 0.01, 0.05
    t_loop = time.perf_counter()
    for it in range(iters):
        zf = Fz.apply(dz)
        zf[sup] = 0.0
        muf = np.clip(Fm.apply(mu), mu_min, 1.0)
        crds = md.crds.copy(); crds[:, 2] = z0 + zf
        pq = md.prop_quads.copy(); pq[:, 1] = E0 * muf ** p_simp
        opts = nat.make_opts(rtol=rtol, use_x0=u_prev is not None, cheb_degree=1)
        val, u, dc, dq, _, fs, _ = h.value_and_grad_host(crds, pq, md.prop_beams, md.loads, want=('crds', 'prop_q'),
                                                          opts=opts, u0=u_prev)
        u_prev = u
        gz = dc[:, 2].copy(); gz[sup] = 0.0
        gz = Fz.apply_T(gz)
        gm = Fm.apply_T(dq[:, 1] * p_simp * E0 * muf ** (p_simp - 1.0))
        hist.append(float(val)); its.append(int(fs.iterations))
        if it == 0 and keep_first:
            first = {'first_gz': gz.copy(), 'first_gm': gm.copy()}
        # projected steepest descent with max-norm scaled steps; densities keep the mean (volume) by a shift
        dz -= step_z * gz / max(np.abs(gz).max(), 1e-300)
        mu_new = mu - step_mu * gm / max(np.abs(gm).max(), 1e-300)
        mu = np.clip(mu_new + (mean_mu - mu_new.mean()), mu_min, 1.0)
    t_loop = time.perf_counter() - t_loop
    h.close()
    out = {'config': f'{N}x{N} shell, {md.n_quad} quads, {md.ndof} dof, {md.design_nodes.shape[0]} z + {md.n_quad} density variables',
           'iterations': iters, 'seconds_loop': t_loop, 's_per_iteration': t_loop / max(iters, 1),
           'setup_s': t_setup, 'filter_setup_s': t_filter, 'history': hist, 'pcg_iterations': its}
    out.update(first)
    return out


if __name__ == '__main__':
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    o = run(N, iters)
    hist, its = o.pop('history'), o.pop('pcg_iterations')
    o.update({'pcg_iterations_first_mean_last': [its[0], float(np.mean(its)), its[-1]],
              'strain_energy_first_last': [hist[0], hist[-1]],
              'monotone_fraction': float(np.mean(np.diff(hist) <= 0))})
    print(json.dumps(o))
