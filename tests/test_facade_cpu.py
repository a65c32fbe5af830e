"""CPU checks of the host mirror of the reference API (Model / SSO_model) that need no GPU:
model freezing, parameter layout, scatter / gather between ``parameter_values`` and the dense arrays
(reference: model.py:221-338, SSO_model.py:125-222), on a symbolic-only handle (device = -1)."""
import numpy as np

from jaxsso_b200 import ElementParameter, Model, NodeParameter, SSO_model
from jaxsso_b200.model import Node


def small_model():
    m = Model(device=-1)
    n = 4
    for j in range(n + 1):
        for i in range(n + 1):
            m.add_node(j * (n + 1) + i, float(i), float(j), 0.1 * i * j)
    e = 0
    for j in range(n):
        for i in range(n):
            a = j * (n + 1) + i
            m.add_quad(e, a + n + 2, a + n + 1, a, a + 1, 0.2, 1e7, 0.3)
            e += 1
    m.add_beamcol(0, 0, 1, 2e8, 8e7, 1e-5, 2e-5, 3e-5, 1e-3)
    for nd in (0, n, n * (n + 1), (n + 1) ** 2 - 1):
        m.add_support(nd, [1, 1, 1, 0, 0, 0])
    m.add_nodal_load(12, [0.0, 0.0, -5.0, 0.0, 0.0, 0.0])
    return m


def test_model_freeze_matches_reference_layout():
    m = small_model()
    m.model_ready()
    assert m.crds.shape == (25, 3) and m.ndof == 150 and m.n_quad == 16 and m.n_beamcol == 1
    assert m.cnct_quads.shape == (16, 4) and m.prop_quads.shape == (16, 5) and m.prop_beamcols.shape == (1, 6)
    assert np.array_equal(np.sort(m.known_id), np.sort(np.concatenate([6 * nd + np.arange(3) for nd in (0, 4, 20, 24)])))
    assert m.nodal_loads[6 * 12 + 2] == -5.0 and np.count_nonzero(m.nodal_loads) == 1
    assert m.handle.n_items == 16 * 16 + 4 and m.handle.n_row == 25
    nd = Node(3, 1.0, 2.0, 3.0)
    assert (nd.nodeTag, nd.X, nd.Y, nd.Z) == (3, 1.0, 2.0, 3.0)


def test_parameter_layout_scatter_gather_and_model_update():
    m = small_model()
    s = SSO_model(m)
    for nd in (6, 7, 12):
        s.add_nodeparameter(NodeParameter(nd, 2))
    s.add_nodeparameter(NodeParameter(8, 0))
    s.add_eleparameter(ElementParameter(3, 1, 0))      # quad 3 thickness
    s.add_eleparameter(ElementParameter(0, 0, 5))      # beam 0 area
    s.add_eleparameter(ElementParameter(5, 1, 1))      # quad 5 E
    s.initialize_parameters_values()
    assert s.n_node_params == 4 and s.n_ele_params == 3 and s.n_bc_params == 1 and s.n_quad_params == 2
    assert np.allclose(s.parameter_values, [0.1 * 1 * 1, 0.1 * 2 * 1, 0.1 * 2 * 2, 3.0, 0.2, 1e-3, 1e7])
    pv = s.parameter_values + np.array([1.0, 2.0, 3.0, 0.5, 0.05, 1e-4, 1e6])
    crds, pb, pq = s._arrays(pv)
    assert crds[6, 2] == pv[0] and crds[7, 2] == pv[1] and crds[12, 2] == pv[2] and crds[8, 0] == pv[3]
    assert pq[3, 0] == pv[4] and pb[0, 5] == pv[5] and pq[5, 1] == pv[6]
    assert np.array_equal(m.crds[6], [1.0, 1.0, 0.1])    # the frozen model is untouched
    # gradient gather is the transpose of the scatter
    rng = np.random.default_rng(0)
    dc, dq, db = rng.standard_normal((25, 3)), rng.standard_normal((16, 5)), rng.standard_normal((1, 6))
    g = s._gather_grad(dc, dq, db)
    assert np.allclose(g, [dc[6, 2], dc[7, 2], dc[12, 2], dc[8, 0], dq[3, 0], db[0, 5], dq[5, 1]])
    # update_model_parameter writes the node parameters back and keeps the symbolic handle
    s.update_parameter(pv)
    h_before = m.handle
    s.update_model_parameter()
    assert m.crds[6, 2] == pv[0] and m.crds[8, 0] == pv[3] and m.handle is h_before
