"""jaxsso_b200 -- B200-native (sm_100a, FP64, hand-written CUDA) implementation of the
per-gradient-evaluation hot path of GaoyuanWu/JaxSSO, behind the reference's own
Model / SSO_model API.  See DESIGN.md and include/jsso.h."""
from . import meshes  # noqa: F401
from .model import Model  # noqa: F401
from .SSO_model import ElementParameter, NodeParameter, SSO_model  # noqa: F401
