"""Wall-clock of one value+gradient evaluation on the reference's own small cases
(BASELINE.md 1a/1b), through the host-buffer C ABI."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
from jaxsso_b200 import _native as nat, meshes
cases = {'barrel_arch_361q': meshes.barrel_arch(), 'beam_arch_99b': meshes.beam_arch(),
         'frames_f10_100 (C1)': meshes.frames(10, 100),
         'mannheim_457q (C2)': meshes.mannheim_quad(dict(np.load('tests/golden/mannheim_quad.npz'))),
         'gridshell_224 (C5, one design)': meshes.gridshell(224, 0)}
out = {}
for name, md in cases.items():
    h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
    opts = nat.make_opts(rtol=1e-10, check_every=int(sys.argv[1]) if len(sys.argv) > 1 else 50)
    ts = []
    for rep in range(6):
        t0 = time.perf_counter()
        val, u, dc, dq, db, fs, bs = h.value_and_grad_host(md.crds, md.prop_quads, md.prop_beams, md.loads, opts=opts)
        ts.append(time.perf_counter() - t0)
    out[name] = {'ms_best': 1e3 * min(ts[1:]), 'pcg_iterations': fs.iterations, 'relres': fs.relres,
                 'us_per_iteration': 1e6 * min(ts[1:]) / max(fs.iterations, 1), 'value': val}
    print(name, json.dumps(out[name]), flush=True)
