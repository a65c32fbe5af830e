import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box via gpurun)')


def _device_count():
    try:
        from jaxsso_b200 import _native as nat
        return int(nat.lib().jsso_device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) where there is no CUDA device, so that a plain `pytest tests` works in
    the build container; `-m gpu` on a GPU box runs them all."""
    if not any('gpu' in it.keywords for it in items) or _device_count() > 0:
        return
    skip = pytest.mark.skip(reason='no CUDA device (the product path has no CPU fallback)')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    with open(os.path.join(GOLDEN_DIR, 'reference_golden.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def mannheim_data():
    return dict(np.load(os.path.join(GOLDEN_DIR, 'mannheim_quad.npz')))


def to_oracle_mesh(md):
    """MeshData (product-side generator) -> oracle Mesh."""
    from oracle import jaxsso_oracle as orc
    return orc.Mesh(md.crds, md.cnct_quads, md.prop_quads, md.cnct_beams, md.prop_beams,
                    md.known, md.loads)
