"""JAX-traceable surface of the hot path (level 2 of the boundary): ``fea_solve`` is a
``jax.custom_vjp`` whose forward / backward are XLA custom calls into libjsso
(csrc/jsso_xla_ffi.cc), so ``jax.grad`` / ``jax.jit`` of user code that calls
``SSO_model.helper_params_to_objective`` keeps working (reference: the custom_vjp solvers of
JaxSSO/solver.py:102-168, 176-250 and the traced use in Examples/Neural_Network_Topo_Shape.ipynb).

IMPORT-GUARDED: jax / jaxlib are not installed in the build image, so this module cannot be
exercised there.  It contains no logic beyond registration and the custom_vjp wiring.
"""
from __future__ import annotations

import ctypes
import os

try:  # pragma: no cover - jax is absent from the build image
    import jax
    import jax.numpy as jnp
    HAVE_JAX = hasattr(jax, 'ffi')
except Exception:  # ImportError or a broken install
    jax = None
    HAVE_JAX = False

_HERE = os.path.dirname(os.path.abspath(__file__))
_registered = False


def register():
    """Register the two FFI targets (needs libjsso_xla.so built against jaxlib's headers)."""
    global _registered
    if not HAVE_JAX:
        raise ImportError('jax >= 0.5 with jax.ffi is required for the traced path')
    if _registered:
        return
    lib = ctypes.CDLL(os.path.join(_HERE, 'libjsso_xla.so'))
    jax.ffi.register_ffi_target('jsso_forward', jax.ffi.pycapsule(lib.jsso_xla_forward), platform='CUDA')
    jax.ffi.register_ffi_target('jsso_backward', jax.ffi.pycapsule(lib.jsso_xla_backward), platform='CUDA')
    _registered = True


def make_fea_solve(model, rtol=1e-10):
    """Return ``fea_solve(crds, prop_q, prop_b) -> u`` for a frozen ``jaxsso_b200.Model``."""
    register()
    import numpy as np
    handle = np.int64(model.handle.h.value)
    f = jnp.asarray(model.nodal_loads)
    ndof = model.ndof

    @jax.custom_vjp
    def fea_solve(crds, prop_q, prop_b):
        return jax.ffi.ffi_call('jsso_forward', jax.ShapeDtypeStruct((ndof,), jnp.float64))(
            crds, prop_q, prop_b, f, handle=handle, rtol=np.float64(rtol))

    def fwd(crds, prop_q, prop_b):
        u = fea_solve(crds, prop_q, prop_b)
        return u, (crds, prop_q, prop_b, u)

    def bwd(res, g):
        crds, prop_q, prop_b, u = res
        out = (jax.ShapeDtypeStruct(crds.shape, jnp.float64), jax.ShapeDtypeStruct(prop_q.shape, jnp.float64),
               jax.ShapeDtypeStruct(prop_b.shape, jnp.float64))
        return tuple(jax.ffi.ffi_call('jsso_backward', out)(crds, prop_q, prop_b, u, g, handle=handle,
                                                          rtol=np.float64(rtol)))

    fea_solve.defvjp(fwd, bwd)
    return fea_solve
