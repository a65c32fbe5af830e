"""Golden fixtures at sizes the CPU oracle needs minutes for (run once here, committed as .npz):

  plate160.npz     160 x 160 jittered MITC4 cap of SURVEY 8(d) (25 921 nodes >= 20 000: the facade's automatic
                   multigrid tier) -- u, compliance, dC/d(crds), dC/d(t, E) from oracle.value_and_grad
                   (refined reduced SuperLU solve + complex-step element derivatives);
  gridshell96.npz  design k = 3 of the 96 x 96 beam-column gridshell generator (BASELINE config 5 at reduced size);
  topo128.npz      first iterate of BASELINE config 4 at 128 x 128: non-uniform density field mu, E = mu^7 E0,
                   hat filters (radius 2.5) applied on the host exactly as Examples/shells_topo_shape.ipynb does
                   (dense formula B_ij = max(0, (R - d_ij)/R) / row sum, evaluated pairwise with a k-d tree):
                   objective, filtered dC/dz and dC/dmu.

  python tests/golden/make_large_fixtures.py [plate160|gridshell96|topo128 ...]
The GPU tests (tests/test_large_parity.py) compare the CUDA path against these files; nothing here runs on the
GPU box."""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
from scipy.spatial import cKDTree

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from jaxsso_b200 import meshes            # noqa: E402  (pure NumPy generators)
from oracle import jaxsso_oracle as orc   # noqa: E402


def oracle_mesh(md, crds=None, prop_quads=None):
    return orc.Mesh(md.crds if crds is None else crds, md.cnct_quads,
                    md.prop_quads if prop_quads is None else prop_quads, md.cnct_beams, md.prop_beams, md.known, md.loads)


def hat_filter_host(xy, R):
    """The notebooks' dense filter restated sparsely: B_ij = w_ij / sum_j w_ij, w_ij = max(0, (R - d_ij)/R)."""
    n = xy.shape[0]
    pairs = cKDTree(xy).query_pairs(R, output_type='ndarray')
    d = np.hypot(*(xy[pairs[:, 0]] - xy[pairs[:, 1]]).T)
    w = np.maximum(0.0, (R - d) / R)
    W = sp.coo_matrix((np.concatenate([w, w, np.ones(n)]),
                       (np.concatenate([pairs[:, 0], pairs[:, 1], np.arange(n)]),
                        np.concatenate([pairs[:, 1], pairs[:, 0], np.arange(n)]))), shape=(n, n)).tocsr()
    return sp.diags(1.0 / np.asarray(W.sum(1)).ravel()) @ W


def topo128_inputs(N=128):
    """Design point of the config-4 fixture (shared with tests/test_large_parity.py through the .npz)."""
    md = meshes.plate(N)
    c = md.crds[md.cnct_quads].mean(1)
    mu = 0.55 + 0.3 * np.sin(2 * np.pi * c[:, 0] / N * 3) * np.cos(2 * np.pi * c[:, 1] / N * 2)
    xi, eta = 2 * md.crds[:, 0] / N - 1, 2 * md.crds[:, 1] / N - 1
    dz = 0.02 * N * np.sin(np.pi * xi) * (1 - eta ** 2)
    return md, mu, dz


def make_plate160():
    md = meshes.plate(160)
    t0 = time.time()
    v, u, lam, dc, dq, _ = orc.value_and_grad(oracle_mesh(md))
    print('plate160 oracle %.1f s, compliance %.15e' % (time.time() - t0, v))
    np.savez_compressed(os.path.join(HERE, 'plate160.npz'), N=160, value=v, u=u, d_crds=dc, d_t=dq[:, 0], d_E=dq[:, 1])


def make_gridshell96():
    md = meshes.gridshell(96, 3)
    t0 = time.time()
    v, u, lam, dc, _, db = orc.value_and_grad(oracle_mesh(md))
    print('gridshell96 oracle %.1f s, compliance %.15e' % (time.time() - t0, v))
    np.savez_compressed(os.path.join(HERE, 'gridshell96.npz'), n=96, k=3, value=v, u=u, d_crds=dc, d_A=db[:, 5], d_Iy=db[:, 2])


def make_topo128():
    N, R, p_simp, mu_min = 128, 2.5, 7.0, 0.1
    md, mu, dz = topo128_inputs(N)
    E0 = float(md.prop_quads[0, 1])
    Bz = hat_filter_host(md.crds[:, :2], R)
    Bm = hat_filter_host(md.crds[md.cnct_quads].mean(1)[:, :2], R)
    sup = md.known[md.known % 6 == 2] // 6
    zf = Bz @ dz
    zf[sup] = 0.0
    muf = np.clip(Bm @ mu, mu_min, 1.0)
    crds = md.crds.copy()
    crds[:, 2] += zf
    pq = md.prop_quads.copy()
    pq[:, 1] = E0 * muf ** p_simp
    t0 = time.time()
    v, u, lam, dc, dq, _ = orc.value_and_grad(oracle_mesh(md, crds, pq))
    print('topo128 oracle %.1f s, objective %.15e' % (time.time() - t0, v))
    gz = dc[:, 2].copy()
    gz[sup] = 0.0
    gz = Bz.T @ gz
    gm = Bm.T @ (dq[:, 1] * p_simp * E0 * muf ** (p_simp - 1.0))
    np.savez_compressed(os.path.join(HERE, 'topo128.npz'), N=N, R=R, value=v, mu=mu, dz=dz, gz=gz, gm=gm)


if __name__ == '__main__':
    which = sys.argv[1:] or ['plate160', 'gridshell96', 'topo128']
    for w in which:
        globals()['make_' + w]()
