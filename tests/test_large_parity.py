"""Parity against the CPU oracle at sizes it needs minutes for: the CUDA path (through the C ABI and the reference-
facing facade) against fixtures generated once by tests/golden/make_large_fixtures.py.

* plate160: 25 921 nodes >= Model.multigrid_min_nodes, i.e. the facade picks the smoothed-aggregation multigrid PCG
  by itself (the tier the 1024^2 benchmark runs on) -- u and compliance <= 1e-8, gradients <= 1e-6 (north_star's
  tolerances), at the solver tolerance bench.py times (rtol 1e-8) and the facade's default;
* gridshell96: one design of the beam-column gridshell generator (BASELINE config 5);
* topo128: the first iterate of BASELINE config 4 (density + shape, hat filters both sides) and ten iterations of
  the optimiser loop (scripts/topo_shape_512.py) on top."""
import os
import sys

import numpy as np
import pytest

import jaxsso_b200 as jb
from jaxsso_b200 import meshes
from tests.conftest import GOLDEN_DIR, ROOT

pytestmark = pytest.mark.gpu


def _fix(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name)))


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b))


def _relmax(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(b).max())


@pytest.mark.parametrize('rtol', [1e-8, 1e-10])
def test_plate160_multigrid_tier_matches_the_oracle(rtol):
    from jaxsso_b200 import _native as nat
    g = _fix('plate160.npz')
    md = meshes.plate(int(g['N']))
    h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
    h.mg_setup()
    val, u, dc, dq, _, fs, _ = h.value_and_grad_host(md.crds, md.prop_quads, md.prop_beams, md.loads,
                                                     want=('crds', 'prop_q'), opts=nat.make_opts(rtol=rtol, cheb_degree=1))
    assert fs.converged and fs.relres <= 1.5 * rtol
    assert _rel(u, g['u']) <= 1e-8, (rtol, _rel(u, g['u']))
    assert abs(val - float(g['value'])) <= 1e-8 * abs(float(g['value']))
    assert _relmax(dc, g['d_crds']) <= 1e-6
    assert _relmax(dq[:, 0], g['d_t']) <= 1e-6 and _relmax(dq[:, 1], g['d_E']) <= 1e-6
    h.close()


def test_plate160_through_the_facade_auto_multigrid():
    """The reference-facing API: Model.solve + SSO_model.value_grad_params on a model large enough for the facade to
    choose the multigrid tier by itself (no solver options passed)."""
    from tests.test_api_gpu import build_model
    g = _fix('plate160.npz')
    md = meshes.plate(int(g['N']))
    model = build_model(md)
    model.model_ready()
    assert md.n_node >= model.multigrid_min_nodes and model.handle.mg_levels
    model.solve(which_solver='sparse', enforce_scipy_sparse=True)
    assert model.last_stats['converged'] and model.last_stats['iterations'] < 400     # block-Jacobi CG needs > 10 000 here
    assert _rel(model.u, g['u']) <= 1e-8
    assert abs(model.strain_energy() - float(g['value'])) <= 1e-8 * float(g['value'])
    sso = jb.SSO_model(model)
    nodes = md.design_nodes[::97]
    for node in nodes:
        sso.add_nodeparameter(jb.NodeParameter(int(node), 2))
    quads = np.arange(0, md.n_quad, 211)
    for q in quads:
        sso.add_eleparameter(jb.ElementParameter(int(q), 1, 0))      # thickness of quad q
    sso.initialize_parameters_values()
    sso.set_objective(objective='strain energy', func=None, func_args=None)
    C, sens = sso.value_grad_params(which_solver='sparse', enforce_scipy_sparse=True)
    assert abs(C - float(g['value'])) <= 1e-8 * float(g['value'])
    ref = np.concatenate([g['d_crds'][nodes, 2], g['d_t'][quads]])
    scale = np.concatenate([np.full(nodes.size, np.abs(g['d_crds']).max()), np.full(quads.size, np.abs(g['d_t']).max())])
    assert np.max(np.abs(sens - ref) / scale) <= 1e-6


def test_gridshell96_design_matches_the_oracle():
    from jaxsso_b200 import _native as nat
    g = _fix('gridshell96.npz')
    md = meshes.gridshell(int(g['n']), int(g['k']))
    h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
    h.mg_setup()
    val, u, dc, _, db, fs, _ = h.value_and_grad_host(md.crds, md.prop_quads, md.prop_beams, md.loads,
                                                     want=('crds', 'prop_b'), opts=nat.make_opts(rtol=1e-9, precond='multigrid'))
    assert fs.converged
    assert _rel(u, g['u']) <= 1e-8
    assert abs(val - float(g['value'])) <= 1e-8 * abs(float(g['value']))
    assert _relmax(dc, g['d_crds']) <= 1e-6
    assert _relmax(db[:, 5], g['d_A']) <= 1e-6 and _relmax(db[:, 2], g['d_Iy']) <= 1e-6
    h.close()


def test_topo128_first_iterate_and_ten_iterations():
    """BASELINE config 4 at 128^2 through scripts/topo_shape_512.py's own loop: the first objective and the FILTERED
    gradients (dC/dz through B_z^T, dC/dmu through the SIMP chain and B_mu^T) against the oracle with the same
    filters applied on the host; then ten projected-gradient iterations must decrease the objective."""
    sys.path.insert(0, os.path.join(ROOT, 'scripts'))
    import topo_shape_512 as ts
    g = _fix('topo128.npz')
    out = ts.run(int(g['N']), 1, mu0=g['mu'], dz0=g['dz'], rtol=1e-9, keep_first=True, radius=float(g['R']))
    assert abs(out['history'][0] - float(g['value'])) <= 1e-8 * float(g['value'])
    assert _relmax(out['first_gz'], g['gz']) <= 1e-6
    assert _relmax(out['first_gm'], g['gm']) <= 1e-6
    out = ts.run(int(g['N']), 10, mu0=g['mu'], dz0=g['dz'], rtol=1e-6)
    hist = out['history']
    assert len(hist) == 10 and hist[-1] < hist[0] and np.mean(np.diff(hist) <= 0) >= 0.8
    assert out['pcg_iterations'][-1] <= out['pcg_iterations'][0]        # warm start pays
