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
    config.addinivalue_line('markers', 'unverified: GPU path written after the round\'s GPU budget was spent; it '
                                       'compiles but has not run on hardware yet (JSSO_RUN_UNVERIFIED=1 runs it)')


def pytest_collection_modifyitems(config, items):
    """Tests of GPU code that has never run on hardware are opt-in, so that a first failure there cannot
    mask the verified parity tests under `-x`; they run with JSSO_RUN_UNVERIFIED=1."""
    if os.environ.get('JSSO_RUN_UNVERIFIED', '0') == '1':
        return
    skip = pytest.mark.skip(reason='not yet run on hardware; set JSSO_RUN_UNVERIFIED=1')
    for it in items:
        if 'unverified' in it.keywords:
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
