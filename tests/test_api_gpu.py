"""The host-side mirror of the reference API (jaxsso_b200.Model / SSO_model) on the GPU path,
written the way the reference's own notebooks use it (Test/shells_ad_validation.ipynb,
Test/beamcols_ad_validation.ipynb, Examples/shells_topo_shape.ipynb)."""
import numpy as np
import pytest

import jaxsso_b200 as jb
from jaxsso_b200 import meshes
from oracle import jaxsso_oracle as orc
from tests.conftest import to_oracle_mesh

pytestmark = pytest.mark.gpu


def build_model(md):
    model = jb.Model()
    for i in range(md.n_node):
        model.add_node(i, *md.crds[i])
    act = np.zeros((md.n_node, 6), int)
    act.reshape(-1)[md.known] = 1
    for i in np.flatnonzero(act.any(1)):
        model.add_support(int(i), act[i].tolist())
    for i in np.flatnonzero(np.abs(md.loads.reshape(-1, 6)).sum(1)):
        model.add_nodal_load(int(i), md.loads.reshape(-1, 6)[i].tolist())
    for e in range(md.n_quad):
        model.add_quad(e, *[int(x) for x in md.cnct_quads[e]], *md.prop_quads[e])
    for e in range(md.n_beam):
        model.add_beamcol(e, *[int(x) for x in md.cnct_beams[e]], *md.prop_beams[e])
    return model


def test_model_solve_shell_arch(golden):
    md = meshes.barrel_arch()
    model = build_model(md)
    model.model_ready()
    model.solve(which_solver='sparse', enforce_scipy_sparse=True)   # reference call signature
    uz = model.u[md.design_nodes * 6 + 2]
    assert abs(uz.min() - golden['shell_arch_min_uz']['dense']) / 25.0 < 1e-8
    assert abs(model.strain_energy() - golden['shell_arch_strain_energy']['value']) / 1.2e6 < 1e-8


def test_sso_shape_gradient_shell_arch(golden):
    md = meshes.barrel_arch()
    model = build_model(md)
    sso = jb.SSO_model(model)
    for node in md.design_nodes:
        sso.add_nodeparameter(jb.NodeParameter(int(node), 2))
    sso.initialize_parameters_values()
    sso.set_objective(objective='strain energy', func=None, func_args=None)
    C, sens = sso.value_grad_params(which_solver='sparse', enforce_scipy_sparse=True)
    g = golden['shell_arch_grad_node201']
    k = int(np.where(md.design_nodes == g['design_i'])[0][0])
    assert abs(sens[k] - g['dense']) / abs(g['dense']) < 1e-6
    assert abs(C - golden['shell_arch_strain_energy']['value']) / C < 1e-8
    assert sens.shape == (md.design_nodes.shape[0],)
    assert abs(sso.params_to_objective() - C) / C < 1e-10
    assert np.allclose(sso.grad_params(), sens, rtol=1e-9)


def test_sso_beam_gradient(golden):
    md = meshes.beam_arch()
    model = build_model(md)
    sso = jb.SSO_model(model)
    for node in md.design_nodes:
        sso.add_nodeparameter(jb.NodeParameter(int(node), 2))
    sso.initialize_parameters_values()
    sso.set_objective('strain energy')
    sso.rtol = 1e-12
    C, sens = sso.value_grad_params()
    g = golden['beam_arch_grad_node49']
    k = int(np.where(md.design_nodes == g['design_i'])[0][0])
    assert abs(sens[k] - g['dense']) / abs(g['dense']) < 1e-5
    assert abs(C - golden['beam_arch']['dense_strain_energy']) / C < 5e-8


def test_sso_mixed_parameters_and_update():
    """Node z + quad E (topology) + quad t (size) parameters in one vector, in add order
    (SSO_model.py:125-173), and update_*parameter between evaluations."""
    md = meshes.plate(12)
    model = build_model(md)
    sso = jb.SSO_model(model)
    for node in md.design_nodes[::3]:
        sso.add_nodeparameter(jb.NodeParameter(int(node), 2))
    for e in range(0, md.n_quad, 2):
        sso.add_eleparameter(jb.ElementParameter(e, ele_type=1, prop_type=1))
    for e in range(1, md.n_quad, 2):
        sso.add_eleparameter(jb.ElementParameter(e, ele_type=1, prop_type=0))
    sso.initialize_parameters_values()
    sso.set_objective('strain energy')
    sso.rtol = 1e-12
    z = sso.nodeparameters_values + 0.03
    ep = sso.eleparameters_values * 1.1
    sso.update_nodeparameter(z)
    sso.update_eleparameter(ep)
    C, sens = sso.value_grad_params()
    crds = md.crds.copy()
    crds[md.design_nodes[::3], 2] = z
    pq = md.prop_quads.copy()
    pq[0::2, 1] *= 1.1
    pq[1::2, 0] *= 1.1
    m = to_oracle_mesh(md)
    rv, ru, rl, rdc, rdq, _ = orc.value_and_grad(m, crds=crds, prop_quads=pq)
    ref = np.concatenate([rdc[md.design_nodes[::3], 2], rdq[0::2, 1], rdq[1::2, 0]])
    assert abs(C - rv) / rv < 1e-8
    nn = md.design_nodes[::3].shape[0]
    assert np.abs(sens[:nn] - ref[:nn]).max() / np.abs(ref[:nn]).max() < 1e-6
    assert np.abs(sens[nn:] - ref[nn:]).max() / np.abs(ref[nn:]).max() < 1e-6


def test_user_objective():
    """A user objective L(u) with its own dL/du (the adjoint right-hand side is arbitrary,
    Examples/Shells_Mannheim_Multihalle_Size.ipynb uses a displacement penalty)."""
    md = meshes.plate(10)
    model = build_model(md)
    sso = jb.SSO_model(model)
    for node in md.design_nodes:
        sso.add_nodeparameter(jb.NodeParameter(int(node), 2))
    sso.initialize_parameters_values()
    w = np.random.default_rng(0).uniform(0.5, 1.5, md.ndof)
    w[md.known] = 0

    def penalty(sso_model, u, weight):
        return float(0.5 * np.sum(weight * u * u)), weight * u

    sso.set_objective('user', func=penalty, func_args=(w,))
    sso.rtol = 1e-12
    L, sens = sso.value_grad_params()
    m = to_oracle_mesh(md)
    rv, ru, rl, rdc, _, _ = orc.value_and_grad(m, g_fn=lambda u: (0.5 * np.sum(w * u * u), w * u))
    assert abs(L - rv) / rv < 1e-8
    ref = rdc[md.design_nodes, 2]
    assert np.abs(sens - ref).max() / np.abs(ref).max() < 1e-6


def test_user_objective_scalar_valued_as_in_the_reference():
    """The reference's signature: func(sso_model, u, *args) -> scalar, differentiated by reverse mode
    (SSO_model.py:303-306, 333-339).  Without jax in this image the gradient dL/du comes from PyTorch autograd: the
    same penalty written with torch ops gives the same value and sensitivities as the (value, gradient) form."""
    import torch
    md = meshes.plate(10)
    w = np.random.default_rng(0).uniform(0.5, 1.5, md.ndof)
    w[md.known] = 0
    out = []
    for form in ('scalar', 'pair'):
        sso = jb.SSO_model(build_model(md))
        for node in md.design_nodes:
            sso.add_nodeparameter(jb.NodeParameter(int(node), 2))
        sso.initialize_parameters_values()
        if form == 'scalar':
            sso.set_objective('user', func=lambda m_, u, weight: 0.5 * torch.sum(torch.as_tensor(weight) * u * u),
                              func_args=(w,))
        else:
            sso.set_objective('user', func=lambda m_, u, weight: (float(0.5 * np.sum(weight * u * u)), weight * u),
                              func_args=(w,))
        sso.rtol = 1e-12
        out.append(sso.value_grad_params())
    assert abs(out[0][0] - out[1][0]) <= 1e-12 * abs(out[1][0])
    assert np.abs(out[0][1] - out[1][1]).max() <= 1e-10 * np.abs(out[1][1]).max()


def test_warm_started_optimiser_steps():
    """Three projected-gradient steps with sso.warm_start: fewer PCG iterations after the first,
    same objective as cold starts."""
    md = meshes.plate(16)
    vals = {}
    for warm in (False, True):
        model = build_model(md)
        sso = jb.SSO_model(model)
        for node in md.design_nodes:
            sso.add_nodeparameter(jb.NodeParameter(int(node), 2))
        sso.initialize_parameters_values()
        sso.set_objective('strain energy')
        sso.warm_start = warm
        its, cs = [], []
        for step in range(3):
            C, g = sso.value_grad_params()
            its.append(sso.last_stats['forward']['iterations'])
            cs.append(C)
            sso.update_nodeparameter(sso.nodeparameters_values - 1e-3 * g / np.abs(g).max())
        vals[warm] = (its, cs)
    assert np.allclose(vals[True][1], vals[False][1], rtol=1e-8)
    assert vals[True][0][1] < vals[False][0][1] and vals[True][0][2] < vals[False][0][2]


def test_solver_plugin_on_reference_K_aug(golden):
    """(K_aug, f_aug) -> u_aug with the reference's solver signature: the oracle's literal
    augmented matrix (assemblemodel.py:111-163) goes in, u and the Lagrange multipliers come out."""
    from jaxsso_b200 import solver
    md = meshes.barrel_arch()
    m = to_oracle_mesh(md)
    K_aug = orc.K_aug(m)
    f_aug = orc.f_aug(m)
    u_aug = solver.b200_solve(K_aug, f_aug, rtol=1e-12)
    ref = orc.solve_literal(m)
    u_ref = orc.solve_refined(m)
    assert u_aug.shape == ref.shape
    assert np.linalg.norm(u_aug[:m.ndof] - u_ref) / np.linalg.norm(u_ref) < 1e-8
    mu_ref = (md.loads - orc.K_global(m) @ u_ref)[md.known]
    assert np.abs(u_aug[m.ndof:] - mu_ref).max() / np.abs(mu_ref).max() < 1e-7
    assert abs(0.5 * f_aug[:m.ndof] @ u_aug[:m.ndof] - golden['shell_arch_strain_energy']['value']) / 1.2e6 < 1e-8
    # CSR triple form, as the reference's pure_callback passes it (solver.py:202-209)
    A = K_aug.tocsr()
    u2 = solver.b200_solve((A.data, A.indices, A.indptr, A.shape), f_aug, rtol=1e-12)
    assert np.allclose(u2, u_aug, rtol=0, atol=1e-9 * np.abs(u_aug).max())


def test_quad_area_and_volume(mannheim_data):
    """Quad.A (element.py:471-487) and the material volume of the size example."""
    from jaxsso_b200 import _native as nat
    md = meshes.mannheim_quad(mannheim_data)
    h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
    area = h.quad_area(nat.DeviceArray.from_host(md.crds)).download()
    ref = orc.quad_area(md.crds[md.cnct_quads].reshape(-1, 12))
    assert np.abs(area - ref).max() / ref.max() < 1e-14
    assert abs(area @ md.prop_quads[:, 0] - ref @ md.prop_quads[:, 0]) / (ref @ md.prop_quads[:, 0]) < 1e-14


def test_sparse_hat_filter_matches_dense_notebook_filter(mannheim_data):
    """The examples' dense filter (Examples/Shells_Mannheim_Multihalle_Shape.ipynb cells 10-13) vs the
    sparse GPU one, forward (B @ z) and transposed (sens @ B)."""
    from jaxsso_b200.filters import HatFilter
    xs = mannheim_data['x'] - mannheim_data['x'].min()
    ys = mannheim_data['y'] - mannheim_data['y'].min()
    R = 10.0
    D_ij = (np.subtract.outer(xs, xs) ** 2 + np.subtract.outer(ys, ys) ** 2) ** 0.5
    B_ini = np.where(D_ij > R, 0, (1 / R) * (R - D_ij))
    B_dense = B_ini / B_ini.sum(axis=1)[:, None]
    f = HatFilter(np.stack([xs, ys], 1), R)
    rng = np.random.default_rng(1)
    z, g = rng.standard_normal(xs.shape[0]), rng.standard_normal(xs.shape[0])
    assert np.abs(f.apply(z) - B_dense @ z).max() < 1e-13
    assert np.abs(f.apply_T(g) - g @ B_dense).max() < 1e-13
    assert f.B.nnz < 0.2 * B_dense.size


def test_gather_rows_and_assembly_profile():
    """jsso_gather_rows (local part of a replicated vector) and the per-kernel timing hooks of
    jsso_assemble; the two assembly paths (warp tasks / chunked) give the same matrix."""
    import os
    from jaxsso_b200 import _native as nat
    rng = np.random.default_rng(3)
    src = rng.standard_normal((50, 6))
    idx = rng.permutation(50)[:31].astype(np.int32)
    out = nat.gather_rows(nat.DeviceArray.from_host(src), nat.DeviceArray.from_host(idx), 6).download()
    assert np.array_equal(out.reshape(-1, 6), src[idx])
    md = meshes.plate(24)
    D = nat.DeviceArray
    crds, pq, pb = D.from_host(md.crds), D.from_host(md.prop_quads), D.from_host(md.prop_beams)
    h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
    h.profile(True)
    h.assemble(crds, pq, pb, apply_bc=True)
    g_ms, t_ms = h.profile_read()
    assert g_ms > 0.0 and t_ms > 0.0
    v_tasks = h.values_host()
    os.environ['JSSO_ASM_CHUNKED'] = '1'
    try:
        h2 = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
    finally:
        del os.environ['JSSO_ASM_CHUNKED']
    h2.assemble(crds, pq, pb, apply_bc=True)
    assert h2.profile_read() == (0.0, 0.0)
    v_chunk = h2.values_host()
    assert np.abs(v_tasks - v_chunk).max() <= 1e-13 * np.abs(v_chunk).max()
