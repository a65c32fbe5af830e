"""Design-variable layer: mirror of JaxSSO/SSO_model.py on the B200 hot path.

Keeps the reference's names and parameter layout (SSO_model.py:33-339):
``parameter_values = [node parameters | element parameters]``, each block in the
order of the ``add_*parameter`` calls; node parameter k is ``crds[tag_k, xyz_k]``,
element parameter k is ``prop_beamcols[tag_k, prop_k]`` (``ele_type == 0``; E, G,
Iy, Iz, J, A) or ``prop_quads[tag_k, prop_k]`` (``ele_type == 1``; t, E, nu, kx, ky).

``value_grad_params`` is ONE call into the C ABI: H2D of coordinates and
properties, fused Ke+assembly, PCG solve, adjoint solve (or lam = u/2 for the
strain-energy objective), fused element sensitivity reduction, D2H of the
gradient arrays; the scatter/gather between ``parameter_values`` and the dense
(n_node,3)/(n_q,5)/(n_b,6) arrays stays on the host exactly as in the reference
(SSO_model.py:207-222).
"""
from __future__ import annotations

import numpy as np

from . import _native as nat


class NodeParameter:
    """reference: SSO_model.py:33-57.  X_Y_Z: 0 X, 1 Y, 2 Z."""

    def __init__(self, nodetag, X_Y_Z=2, upper_bound=None, lower_bound=None):
        self.tag = nodetag
        self.XYZ = X_Y_Z


class ElementParameter:
    """reference: SSO_model.py:59-89.  ele_type 0 BeamCol, 1 Quad."""

    def __init__(self, eletag, ele_type=0, prop_type=0, upper_bound=None, lower_bound=None):
        self.tag = eletag
        self.type = ele_type
        self.prop = prop_type


class SSO_model:
    def __init__(self, model):
        self.model = model
        self.nodeparameters_tags = []
        self.nodeparameters_xyzs = []
        self.eleparameters_tags = []
        self.ele_types = []
        self.eleparameters_props = []
        self.objective = None
        self.objective_args = None
        self.rtol = 1e-10
        self.warm_start = False   # start each solve from the previous design's u (optimiser loops)
        self._u_prev = None
        self.last_stats = None

    # ---- parameters (SSO_model.py:125-196) -------------------------------------------
    def add_nodeparameter(self, nodeparameter):
        self.nodeparameters_tags.append(nodeparameter.tag)
        self.nodeparameters_xyzs.append(nodeparameter.XYZ)

    def add_eleparameter(self, eleparameter):
        self.eleparameters_tags.append(eleparameter.tag)
        self.ele_types.append(eleparameter.type)
        self.eleparameters_props.append(eleparameter.prop)

    def initialize_nodeparameters_values(self):
        self.model.model_ready()
        self._np_tags = np.asarray(self.nodeparameters_tags, dtype=np.int64)
        self._np_xyzs = np.asarray(self.nodeparameters_xyzs, dtype=np.int64)
        self.nodeparameters_values = self.model.crds[self._np_tags, self._np_xyzs].astype(float)

    def initialize_eleparameters_values(self):
        self.model.model_ready()
        types = np.asarray(self.ele_types, dtype=np.int64)
        tags = np.asarray(self.eleparameters_tags, dtype=np.int64)
        props = np.asarray(self.eleparameters_props, dtype=np.int64)
        self.parameters_bc = np.flatnonzero(types == 0)
        self.parameters_quads = np.flatnonzero(types == 1)
        self._ep_tags, self._ep_props = tags, props
        vals = np.zeros(types.shape[0])
        vals[self.parameters_bc] = self.model.prop_beamcols[tags[self.parameters_bc], props[self.parameters_bc]]
        vals[self.parameters_quads] = self.model.prop_quads[tags[self.parameters_quads], props[self.parameters_quads]]
        self.eleparameters_values = vals

    def initialize_parameters_values(self):
        self.initialize_nodeparameters_values()
        self.initialize_eleparameters_values()
        self.n_node_params = len(self.nodeparameters_tags)
        self.n_ele_params = len(self.eleparameters_tags)
        self.n_bc_params = self.parameters_bc.shape[0]
        self.n_quad_params = self.parameters_quads.shape[0]
        self.parameter_values = np.concatenate((self.nodeparameters_values, self.eleparameters_values))

    def update_nodeparameter(self, params_values):
        self.nodeparameters_values = np.array(params_values, dtype=float)
        self.parameter_values[:self.n_node_params] = self.nodeparameters_values

    def update_eleparameter(self, params_values):
        self.eleparameters_values = np.array(params_values, dtype=float)
        self.parameter_values[self.n_node_params:] = self.eleparameters_values

    def update_parameter(self, params_values):
        self.parameter_values = np.array(params_values, dtype=float)

    def update_model_parameter(self):
        """Write the current node-parameter values back into the model's nodes and re-freeze it
        (reference: SSO_model.py:198-202; the symbolic state is kept, the topology did not change)."""
        for tag, xyz, val in zip(self.nodeparameters_tags, self.nodeparameters_xyzs,
                                 np.asarray(self.parameter_values[:self.n_node_params], dtype=float).tolist()):
            self.model.update_node(tag, xyz, val)
        self.model.model_ready()

    # ---- parameters -> arrays (SSO_model.py:207-222) -----------------------------------
    def node_params_crds(self, nodeparameter_values):
        crds = self.model.crds.copy()
        crds[self._np_tags, self._np_xyzs] = nodeparameter_values
        return crds

    def ele_params_props(self, eleparameter_values):
        pb = self.model.prop_beamcols.copy()
        pq = self.model.prop_quads.copy()
        b, q = self.parameters_bc, self.parameters_quads
        pb[self._ep_tags[b], self._ep_props[b]] = eleparameter_values[b]
        pq[self._ep_tags[q], self._ep_props[q]] = eleparameter_values[q]
        return pb, pq

    def _arrays(self, parameter_values):
        parameter_values = np.asarray(parameter_values, dtype=float)
        crds = (self.node_params_crds(parameter_values[:self.n_node_params])
                if self.n_node_params > 0 else self.model.crds)
        if self.n_ele_params > 0:
            pb, pq = self.ele_params_props(parameter_values[self.n_node_params:])
        else:
            pb, pq = self.model.prop_beamcols, self.model.prop_quads
        return crds, pb, pq

    def _gather_grad(self, d_crds, d_pq, d_pb):
        g = np.zeros(self.n_node_params + self.n_ele_params)
        if self.n_node_params > 0:
            g[:self.n_node_params] = d_crds[self._np_tags, self._np_xyzs]
        if self.n_ele_params > 0:
            ge = np.zeros(self.n_ele_params)
            b, q = self.parameters_bc, self.parameters_quads
            if b.size:
                ge[b] = d_pb[self._ep_tags[b], self._ep_props[b]]
            if q.size:
                ge[q] = d_pq[self._ep_tags[q], self._ep_props[q]]
            g[self.n_node_params:] = ge
        return g

    # ---- parameters -> results (SSO_model.py:224-259) -----------------------------------
    def params_u(self, parameter_values, which_solver='b200', enforce_scipy_sparse=True):
        crds, pb, pq = self._arrays(parameter_values)
        solver = self.model.select_solver(which_solver, enforce_scipy_sparse)
        return solver(crds, pb, pq, nat.make_opts(rtol=self.rtol))

    def params_c(self, parameter_values, which_solver='b200', enforce_scipy_sparse=True):
        return 0.5 * self.model.nodal_loads @ self.params_u(parameter_values, which_solver, enforce_scipy_sparse)

    def node_params_c(self, nodeparameter_values, which_solver='b200', enforce_scipy_sparse=True):
        """Strain energy as a function of the node parameters only, element parameters at their current
        values (reference: SSO_model.py:262-271)."""
        pv = np.array(self.parameter_values, dtype=float)
        pv[:self.n_node_params] = nodeparameter_values
        return self.params_c(pv, which_solver, enforce_scipy_sparse)

    # ---- objective (SSO_model.py:275-306) ------------------------------------------------
    def set_objective(self, objective='strain energy', func=None, func_args=None):
        """'strain energy': 0.5 f.u (SSO_model.py:297-301).
        'user': ``func(sso_model, u, *args)`` as in the reference (SSO_model.py:303-306).  The reference traces it with
        JAX and gets the adjoint right-hand side g = dL/du from reverse mode; here it may
          * return the scalar value alone, as in the reference: g comes from ``jax.grad`` when jax is importable,
            else from PyTorch autograd (``u`` is then a float64 ``torch.Tensor``; write the objective with torch ops);
          * or return ``(value, dvalue_du)`` with the gradient supplied by the caller (plain NumPy)."""
        if objective == 'strain energy':
            self.objective = 'strain energy'
        elif objective == 'user':
            self.objective = func
            if func_args is not None:
                self.objective_args = func_args
        else:
            raise ValueError("objective must be 'strain energy' or 'user'")

    # ---- value / gradient (SSO_model.py:313-339) -------------------------------------------
    def helper_params_to_objective(self, parameter_values, which_solver='b200', enforce_scipy_sparse=True):
        u = self.params_u(parameter_values, which_solver, enforce_scipy_sparse)
        if self.objective == 'strain energy':
            return 0.5 * self.model.nodal_loads @ u
        return self._user_objective(u)[0]

    def _user_objective(self, u):
        """(value, dL/du) of the user objective at u (NumPy), by whichever route the callable supports."""
        args = self.objective_args or ()
        try:
            out = self.objective(self, u, *args)
        except (TypeError, AttributeError, ValueError):
            out = None     # a callable written for jax / torch arrays may not accept NumPy input
        if isinstance(out, tuple) and len(out) == 2:
            return float(out[0]), np.asarray(out[1], dtype=np.float64).ravel()
        try:
            import jax
            import jax.numpy as jnp
            val, g = jax.value_and_grad(lambda uu: self.objective(self, uu, *args))(jnp.asarray(u))
            return float(val), np.asarray(g, dtype=np.float64).ravel()
        except ImportError:
            pass
        import torch
        ut = torch.tensor(np.asarray(u, dtype=np.float64), dtype=torch.float64, requires_grad=True)
        val = self.objective(self, ut, *args)
        if not torch.is_tensor(val):
            raise TypeError("a 'user' objective must return (value, dvalue_du), or a scalar computed with jax / torch "
                            "operations from u so that its gradient can be taken")
        g, = torch.autograd.grad(val, ut)
        return float(val.detach()), g.detach().numpy().astype(np.float64).ravel()

    def params_to_objective(self, which_solver='b200', enforce_scipy_sparse=True):
        return self.helper_params_to_objective(self.parameter_values, which_solver, enforce_scipy_sparse)

    def value_grad_params(self, which_solver='b200', enforce_scipy_sparse=True):
        """Value and gradient of the objective w.r.t. ``parameter_values``."""
        if self.objective is None:
            raise RuntimeError('set_objective first')
        crds, pb, pq = self._arrays(self.parameter_values)
        h = self.model.handle
        opts = nat.make_opts(rtol=self.rtol)
        if self.objective == 'strain energy':
            u0 = self._u_prev if (self.warm_start and self._u_prev is not None) else None
            if u0 is not None and u0.size != 6 * h.n_node:      # the model was rebuilt with another mesh
                u0 = self._u_prev = None
            val, u, dc, dq, db, fs, bs = h.value_and_grad_host(crds, pq, pb, self.model.nodal_loads, opts=opts,
                                                               u0=u0, allow_noconv=True)
            self._u_prev = u
            self.last_stats = {'forward': fs.as_dict(), 'backward': bs.as_dict()}
            self.model.u = u
            return val, self._gather_grad(dc, dq, db)
        return self._value_grad_user(crds, pb, pq, opts)

    def _value_grad_user(self, crds, pb, pq, opts):
        h = self.model.handle
        m = self.model
        d = {k: nat.DeviceArray.from_host(v) for k, v in
             dict(crds=crds, pq=pq, pb=pb, f=m.nodal_loads).items()}
        u_d = nat.DeviceArray((m.ndof,))
        fs = h.forward(d['crds'], d['pq'], d['pb'], d['f'], u_d, opts=opts)
        u = u_d.download()
        val, g = self._user_objective(u)
        g_d = nat.DeviceArray.from_host(np.asarray(g, dtype=float))
        dc, dq, db = nat.DeviceArray((m.crds.shape[0], 3)), nat.DeviceArray((m.n_quad, 5)), nat.DeviceArray((m.n_beamcol, 6))
        bs = h.backward(d['crds'], d['pq'], d['pb'], u_d, g_d, dc, dq if m.n_quad else None,
                        db if m.n_beamcol else None, opts=opts)
        self.last_stats = {'forward': fs.as_dict(), 'backward': bs.as_dict()}
        m.u = u
        return val, self._gather_grad(dc.download(), dq.download(), db.download())

    def grad_params(self, which_solver='b200', enforce_scipy_sparse=True):
        return self.value_grad_params(which_solver, enforce_scipy_sparse)[1]
