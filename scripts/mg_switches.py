"""A/B of the multigrid-PCG opt-in switches on one mesh, one process: the symbolic hierarchy is built once and
re-uploaded into a fresh handle per configuration (the switches are read by jsso_mg_setup).
  python scripts/mg_switches.py [N] [rtol]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxsso_b200 import _native as nat, meshes, multigrid

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rtol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-8
md = meshes.plate(N)
D = nat.DeviceArray
crds, pq, pb, f = D.from_host(md.crds), D.from_host(md.prop_quads), D.from_host(md.prop_beams), D.from_host(md.loads)
u = D((md.ndof,))
levels = None
u_ref = None
CONFIGS = [{}, {'JSSO_MG_FP16': '0'}, {'JSSO_MG_GRAPH': '0'}] if len(sys.argv) <= 3 else [{}]
if os.environ.get('MG_SWITCH_CONFIGS'):        # a JSON list of environments, e.g. '[{}, {"JSSO_MG_POLL": "4"}]'
    CONFIGS = json.loads(os.environ['MG_SWITCH_CONFIGS'])
DEGREES = tuple(int(d) for d in os.environ.get('MG_SWITCH_DEGREES', '1,2').split(','))
KEYS = tuple(sorted({k for c in CONFIGS for k in c})) + ('JSSO_MG_FP16', 'JSSO_MG_POLL', 'JSSO_MG_GRAPH', 'JSSO_MG_FP64')
for cfg in CONFIGS:
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update({k: v for k, v in cfg.items() if k != 'MAX_COARSE'})
    h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
    mc = int(cfg.get('MAX_COARSE', 64))      # not an environment switch: nodes of the coarsest (dense) level
    if levels is None or mc not in levels:
        rp, ci = h.pattern()
        t0 = time.perf_counter()
        levels = dict(levels or {})
        levels[mc] = multigrid.build_hierarchy(rp, ci, max_coarse_nodes=mc)
        print('symbolic hierarchy (max_coarse_nodes %d): %.1f s, %d levels' % (mc, time.perf_counter() - t0, len(levels[mc])), flush=True)
    h.mg_setup(levels=levels[mc])
    for deg in DEGREES:
        best = None
        for rep in range(3):
            t0 = time.perf_counter()
            st = h.forward(crds, pq, pb, f, u, opts=nat.make_opts(rtol=rtol, precond='multigrid', cheb_degree=deg))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        # setup share: one iteration only
        t0 = time.perf_counter()
        try:
            h.forward(crds, pq, pb, f, u, opts=nat.make_opts(rtol=rtol, precond='multigrid', cheb_degree=deg, maxiter=1))
        except nat.JssoError:
            pass
        t_setup = time.perf_counter() - t0
        st = h.forward(crds, pq, pb, f, u, opts=nat.make_opts(rtol=rtol, precond='multigrid', cheb_degree=deg))
        uh = u.download()
        if u_ref is None:
            u_ref = uh.copy()
        print('MG_SWITCH ' + json.dumps({'cfg': cfg, 'cheb': deg, 'forward_s': best, 'setup_plus_1it_s': t_setup,
                                         'iterations': st.iterations, 'relres': st.relres,
                                         'ms_per_iteration': 1e3 * (best - t_setup) / max(st.iterations - 1, 1),
                                         'u_rel_diff_vs_default': float(np.linalg.norm(uh - u_ref) / np.linalg.norm(u_ref))}),
              flush=True)
    h.close()
