"""BASELINE config 5: a batch of synthetic 224x224 beam-column gridshell designs, sharded over the
ranks (replicas: one handle per GPU, no collective on the hot path).  Prints designs/s.
  python scripts/gridshell_batch.py [N_DESIGNS]          (or under torchrun for several GPUs)"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxsso_b200 import _native as nat, meshes
rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
n_designs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nat.lib().jsso_set_device(local)
mine = list(range(rank, n_designs, world))
md0 = meshes.gridshell(224, 0)
h = nat.Handle(md0.n_node, md0.cnct_quads, md0.cnct_beams, md0.known, device=local)
h.mg_setup()
designs = [meshes.gridshell(224, k) for k in mine]
opts = nat.make_opts(rtol=1e-8)
h.value_and_grad_host(designs[0].crds, md0.prop_quads, md0.prop_beams, md0.loads, opts=opts)   # warm-up
t0 = time.perf_counter()
out = []
for d in designs:
    val, u, dc, _, _, fs, _ = h.value_and_grad_host(d.crds, md0.prop_quads, md0.prop_beams, md0.loads,
                                                    want=('crds',), opts=opts)
    out.append((val, fs.iterations))
dt = time.perf_counter() - t0
print(json.dumps({'rank': rank, 'designs': len(mine), 'seconds': dt, 'designs_per_s_this_rank': len(mine) / dt,
                  'beams': md0.n_beam, 'dof': md0.ndof, 'pcg_iterations': [o[1] for o in out][:4],
                  'compliance': [o[0] for o in out][:2]}))
