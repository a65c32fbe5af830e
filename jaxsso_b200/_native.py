"""ctypes binding of libjsso.so (include/jsso.h).

There is NO CPU fallback: importing this module without the built library, or
creating a handle without a CUDA device, raises.  Device memory is managed
through the library's own thin cudart wrappers, so the product path needs
neither PyTorch nor cuda-python.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('JSSO_LIB') or os.path.join(_HERE, 'libjsso.so')   # JSSO_LIB: a tuning variant

ERR_NAMES = {0: 'OK', 1: 'ARG', 2: 'CUDA', 3: 'NOCONV', 4: 'NAN', 5: 'BADJAC', 6: 'DEGENERATE_BEAM',
             7: 'NOT_SPD', 8: 'NCCL', 9: 'STATE'}
JSSO_ERR_NOCONV = 3
JSSO_DEVICE_NONE = -1   # symbolic-only handle (host pattern / task lists, no device)


class JssoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f'libjsso error {code} ({ERR_NAMES.get(code, "?")}): {msg}')
        self.code = code


class MeshDesc(C.Structure):
    _fields_ = [('n_node', C.c_int32), ('n_row', C.c_int32), ('n_quad', C.c_int32),
                ('cnct_quads', C.c_void_p), ('n_beam', C.c_int32), ('cnct_beams', C.c_void_p),
                ('n_known', C.c_int32), ('known', C.c_void_p), ('device', C.c_int32)]


class Sizes(C.Structure):
    _fields_ = [('n_node', C.c_int32), ('n_row', C.c_int32), ('n_quad', C.c_int32), ('n_beam', C.c_int32),
                ('nnzb', C.c_int64), ('n_items', C.c_int64), ('n_chunk', C.c_int32)]


class Stats(C.Structure):
    _fields_ = [('iterations', C.c_int32), ('restarts', C.c_int32), ('converged', C.c_int32),
                ('flags', C.c_int32), ('relres', C.c_double), ('relres_recur', C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SolveOpts(C.Structure):
    _fields_ = [('rtol', C.c_double), ('maxiter', C.c_int32), ('check_every', C.c_int32),
                ('use_x0', C.c_int32), ('compliance', C.c_int32), ('precond', C.c_int32),
                ('cheb_degree', C.c_int32)]


_MG_INT_FIELDS = ['n_f', 'n_c', 'nnz_p', 'nnz_ap', 'nnz_c']
_MG_PTR_FIELDS = ['agg', 'p_rowptr', 'p_col', 'p_own', 'ps_ptr', 'ps_a', 'ps_j', 'apl_ptr', 'apl_a', 'apl_p',
                  'c_rowptr', 'c_col', 'c_diag', 'cl_ptr', 'cl_p', 'cl_ap', 'pt_rowptr', 'pt_col', 'pt_src',
                  'mem_ptr', 'mem']


class MgLevelDesc(C.Structure):
    _fields_ = [(k, C.c_int32) for k in _MG_INT_FIELDS] + [(k, C.c_void_p) for k in _MG_PTR_FIELDS]



class MgHaloDesc(C.Structure):
    _fields_ = [('n_peer', C.c_int32)] + [(k, C.c_void_p) for k in ('peer_rank', 'send_ptr', 'send_idx', 'recv_ptr',
                                                                      'recv_idx', 'remote_off')] + \
               [('n_ghost', C.c_int32), ('ghost_rows', C.c_void_p)]

class MgSetupDesc(C.Structure):
    _fields_ = [('n_p_slots', C.c_int32), ('n_ap_slots', C.c_int32), ('p_slots', C.c_void_p), ('ap_slots', C.c_void_p),
                ('ac_bounds', C.c_void_p), ('p_own_lo', C.c_int32), ('p_own_hi', C.c_int32), ('pt_own_lo', C.c_int32),
                ('pt_own_hi', C.c_int32), ('scale_row_lo', C.c_int32), ('scale_row_hi', C.c_int32),
                ('factor_row_lo', C.c_int32), ('factor_row_hi', C.c_int32)]


PRECOND = {'auto': 0, 'block_jacobi': 1, 'multigrid': 2}


# every symbol include/jsso.h declares (checked by tests/test_abi.py)
SYMBOLS = ['jsso_create', 'jsso_create_from_bsr', 'jsso_set_values_host', 'jsso_destroy', 'jsso_last_error', 'jsso_get_sizes', 'jsso_pattern', 'jsso_mg_aggregate', 'jsso_mg_pattern_lists', 'jsso_assembly_tasks',
           'jsso_quad_ke', 'jsso_beam_ke', 'jsso_quad_area', 'jsso_csr_spmv', 'jsso_assemble', 'jsso_gather_rows', 'jsso_profile', 'jsso_profile_read', 'jsso_assemble_from_ke', 'jsso_get_values',
           'jsso_get_values_host', 'jsso_get_flags', 'jsso_spmv', 'jsso_pcg', 'jsso_mg_setup', 'jsso_mg_set_dist', 'jsso_mg_p2p_reserve', 'jsso_mg_p2p_export', 'jsso_mg_p2p_connect', 'jsso_mg_set_dist_setup', 'jsso_mg_dist_counters', 'jsso_adjoint',
           'jsso_forward', 'jsso_backward', 'jsso_value_and_grad_host', 'jsso_assemble_adjoint_host', 'jsso_nccl_unique_id',
           'jsso_set_halo', 'jsso_p2p_export', 'jsso_p2p_connect', 'jsso_halo_exchange', 'jsso_set_device', 'jsso_dev_alloc', 'jsso_dev_free',
           'jsso_host_alloc_pinned', 'jsso_host_free_pinned', 'jsso_memcpy_h2d', 'jsso_memcpy_d2h',
           'jsso_memset', 'jsso_stream_sync', 'jsso_device_count', 'jsso_event_create',
           'jsso_event_record', 'jsso_event_elapsed_ms', 'jsso_event_destroy', 'jsso_launch_count', 'jsso_fp64_peak', 'jsso_profiler_range']

_lib = None


def lib():
    """Load libjsso.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f'{LIB_PATH} is missing: build it with `python -m jaxsso_b200.build` '
                          '(jaxsso_b200 has no CPU fallback)')
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.jsso_create.argtypes = [C.POINTER(MeshDesc), C.POINTER(vp)]
    L.jsso_create_from_bsr.argtypes = [i32, vp, vp, i32, vp, i32, C.POINTER(vp)]
    L.jsso_set_values_host.argtypes = [vp, vp, C.c_int]
    L.jsso_destroy.argtypes = [vp]
    L.jsso_destroy.restype = None
    L.jsso_last_error.argtypes = [vp]
    L.jsso_last_error.restype = C.c_char_p
    L.jsso_get_sizes.argtypes = [vp, C.POINTER(Sizes)]
    L.jsso_pattern.argtypes = [vp, vp, vp]
    L.jsso_assembly_tasks.argtypes = [vp] * 8
    L.jsso_mg_aggregate.argtypes = [i32, vp, vp, vp, C.POINTER(i32)]
    L.jsso_mg_pattern_lists.argtypes = [i64, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, C.POINTER(i64)]
    L.jsso_quad_ke.argtypes = [vp, vp, vp, vp, vp]
    L.jsso_beam_ke.argtypes = [vp, vp, vp, vp, vp]
    L.jsso_quad_area.argtypes = [vp, vp, vp, vp]
    L.jsso_csr_spmv.argtypes = [i32, vp, vp, vp, vp, vp, vp]
    L.jsso_assemble.argtypes = [vp, vp, vp, vp, C.c_int, vp]
    L.jsso_assemble_from_ke.argtypes = [vp, vp, vp, C.c_int, vp]
    L.jsso_profile.argtypes = [vp, C.c_int]
    L.jsso_gather_rows.argtypes = [vp, vp, i32, i32, vp, vp]
    L.jsso_profile_read.argtypes = [vp, vp]
    L.jsso_get_values.argtypes = [vp, vp, vp]
    L.jsso_get_values_host.argtypes = [vp, vp]
    L.jsso_get_flags.argtypes = [vp, C.POINTER(i32)]
    L.jsso_spmv.argtypes = [vp, vp, vp, vp]
    L.jsso_pcg.argtypes = [vp, vp, vp, C.POINTER(SolveOpts), C.POINTER(Stats), vp]
    L.jsso_mg_setup.argtypes = [vp, i32, C.POINTER(MgLevelDesc)]
    L.jsso_mg_set_dist.argtypes = [vp, vp, i32, i32, i32, vp, i32, C.POINTER(MgHaloDesc)]
    L.jsso_mg_set_dist_setup.argtypes = [vp, i32, C.POINTER(MgSetupDesc)]
    L.jsso_mg_dist_counters.argtypes = [vp, vp]
    L.jsso_mg_p2p_reserve.argtypes = [vp, i32]
    L.jsso_mg_p2p_export.argtypes = [vp, vp]
    L.jsso_mg_p2p_connect.argtypes = [vp, vp, vp]
    L.jsso_adjoint.argtypes = [vp] + [vp] * 8 + [vp]
    L.jsso_forward.argtypes = [vp, vp, vp, vp, vp, vp, C.POINTER(SolveOpts), C.POINTER(Stats), vp]
    L.jsso_backward.argtypes = [vp] + [vp] * 9 + [C.POINTER(SolveOpts), C.POINTER(Stats), vp]
    L.jsso_value_and_grad_host.argtypes = [vp, vp, vp, vp, vp, C.POINTER(dbl), vp, vp, vp, vp,
                                           C.POINTER(SolveOpts), C.POINTER(Stats), C.POINTER(Stats)]
    L.jsso_assemble_adjoint_host.argtypes = [vp] * 9
    L.jsso_nccl_unique_id.argtypes = [vp]
    L.jsso_set_halo.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]
    L.jsso_halo_exchange.argtypes = [vp, vp, vp]
    L.jsso_p2p_export.argtypes = [vp, vp]
    L.jsso_p2p_connect.argtypes = [vp, vp, vp]
    L.jsso_set_device.argtypes = [C.c_int]
    L.jsso_dev_alloc.argtypes = [C.c_size_t]
    L.jsso_dev_alloc.restype = vp
    L.jsso_dev_free.argtypes = [vp]
    L.jsso_dev_free.restype = None
    L.jsso_host_alloc_pinned.argtypes = [C.c_size_t]
    L.jsso_host_alloc_pinned.restype = vp
    L.jsso_host_free_pinned.argtypes = [vp]
    L.jsso_host_free_pinned.restype = None
    L.jsso_memcpy_h2d.argtypes = [vp, vp, C.c_size_t, vp]
    L.jsso_memcpy_d2h.argtypes = [vp, vp, C.c_size_t, vp]
    L.jsso_memset.argtypes = [vp, C.c_int, C.c_size_t, vp]
    L.jsso_stream_sync.argtypes = [vp]
    L.jsso_device_count.argtypes = []
    L.jsso_event_create.argtypes = []
    L.jsso_event_create.restype = vp
    L.jsso_event_record.argtypes = [vp, vp]
    L.jsso_event_elapsed_ms.argtypes = [vp, vp, C.POINTER(C.c_float)]
    L.jsso_event_destroy.argtypes = [vp]
    L.jsso_event_destroy.restype = None
    L.jsso_launch_count.argtypes = []
    L.jsso_profiler_range.argtypes = [C.c_int]
    L.jsso_fp64_peak.argtypes = [i32, dbl, C.POINTER(dbl), C.POINTER(dbl)]
    L.jsso_launch_count.restype = i64
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class DeviceArray:
    """A flat device buffer with a NumPy dtype/shape tag."""

    def __init__(self, shape, dtype=np.float64):
        self.shape = tuple(np.atleast_1d(shape).tolist()) if not isinstance(shape, tuple) else shape
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        self.ptr = lib().jsso_dev_alloc(max(self.nbytes, 8))
        if not self.ptr:
            raise MemoryError(f'cudaMalloc({self.nbytes}) failed')

    @classmethod
    def from_host(cls, a, dtype=None, stream=None):
        a = np.ascontiguousarray(a, dtype=dtype)
        d = cls(a.shape, a.dtype)
        d.upload(a, stream)
        return d

    def upload(self, a, stream=None):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        assert a.nbytes == self.nbytes, (a.shape, self.shape)
        if self.nbytes:
            rc = lib().jsso_memcpy_h2d(self.ptr, _ptr(a), self.nbytes, stream)
            if rc:
                raise JssoError(rc, 'memcpy h2d')
            lib().jsso_stream_sync(stream)   # the source is pageable NumPy memory

    def download(self, stream=None):
        out = np.empty(self.shape, self.dtype)
        if self.nbytes:
            rc = lib().jsso_memcpy_d2h(_ptr(out), self.ptr, self.nbytes, stream)
            if rc:
                raise JssoError(rc, 'memcpy d2h')
        return out

    def zero(self, stream=None):
        lib().jsso_memset(self.ptr, 0, self.nbytes, stream)

    def free(self):
        if getattr(self, 'ptr', None):
            lib().jsso_dev_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedArray(np.ndarray):
    """NumPy array backed by page-locked host memory (cudaHostAlloc) so that the library's
    host-buffer entry points can DMA straight from / into it."""
    _keep = None


def pinned_empty(shape, dtype=np.float64):
    shape = tuple(np.atleast_1d(shape).tolist()) if not isinstance(shape, tuple) else shape
    nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
    ptr = lib().jsso_host_alloc_pinned(max(nbytes, 8))
    if not ptr:
        raise MemoryError('cudaHostAlloc failed')
    buf = (C.c_char * max(nbytes, 8)).from_address(ptr)
    a = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape).view(PinnedArray)
    a._keep = (buf, ptr)   # freed when the process exits; benchmark buffers live that long
    return a


def pinned_copy(a):
    out = pinned_empty(a.shape, a.dtype)
    out[...] = a
    return out


def _dp(a):
    """Device pointer of a DeviceArray / raw int / None."""
    if a is None:
        return None
    if isinstance(a, DeviceArray):
        return a.ptr
    return C.c_void_p(int(a))


def make_opts(rtol=1e-10, maxiter=200000, check_every=50, use_x0=False, compliance=False, precond='auto',
              cheb_degree=2):
    return SolveOpts(float(rtol), int(maxiter), int(check_every), int(bool(use_x0)), int(bool(compliance)),
                     PRECOND[precond] if isinstance(precond, str) else int(precond), int(cheb_degree))


class Handle:
    """One frozen model on one GPU (jsso_handle)."""

    def __init__(self, n_node, cnct_quads=None, cnct_beams=None, known=None, device=0, n_row=None):
        L = lib()
        cq = np.ascontiguousarray(np.zeros((0, 4)) if cnct_quads is None else cnct_quads, dtype=np.int32).reshape(-1, 4)
        cb = np.ascontiguousarray(np.zeros((0, 2)) if cnct_beams is None else cnct_beams, dtype=np.int32).reshape(-1, 2)
        kn = np.ascontiguousarray(np.zeros(0) if known is None else known, dtype=np.int32).ravel()
        self._keep = (cq, cb, kn)
        desc = MeshDesc(int(n_node), int(n_node if n_row is None else n_row), cq.shape[0], _ptr(cq),
                        cb.shape[0], _ptr(cb), kn.shape[0], _ptr(kn), int(device))
        h = C.c_void_p()
        rc = L.jsso_create(C.byref(desc), C.byref(h))
        if rc:
            raise JssoError(rc, L.jsso_last_error(None).decode())
        self.h = h
        s = Sizes()
        L.jsso_get_sizes(self.h, C.byref(s))
        self.sizes = s
        self.n_node, self.n_row, self.n_quad, self.n_beam = s.n_node, s.n_row, s.n_quad, s.n_beam
        self.nnzb, self.n_items, self.n_chunk = s.nnzb, s.n_items, s.n_chunk
        self.device = device

    @classmethod
    def from_bsr(cls, rowptr, colidx, known=None, device=0):
        """Pattern-only handle (solver-plugin compatibility mode): values come from set_values."""
        L = lib()
        self = cls.__new__(cls)
        rp = np.ascontiguousarray(rowptr, np.int32)
        ci = np.ascontiguousarray(colidx, np.int32)
        kn = np.ascontiguousarray(np.zeros(0) if known is None else known, dtype=np.int32).ravel()
        h = C.c_void_p()
        rc = L.jsso_create_from_bsr(rp.shape[0] - 1, _ptr(rp), _ptr(ci), kn.shape[0], _ptr(kn), int(device), C.byref(h))
        if rc:
            raise JssoError(rc, L.jsso_last_error(None).decode())
        self.h = h
        s = Sizes()
        L.jsso_get_sizes(self.h, C.byref(s))
        self.sizes = s
        self.n_node, self.n_row, self.n_quad, self.n_beam = s.n_node, s.n_row, s.n_quad, s.n_beam
        self.nnzb, self.n_items, self.n_chunk = s.nnzb, s.n_items, s.n_chunk
        self.device = device
        return self

    def set_values(self, blocks, apply_bc=True):
        """blocks: (nnzb, 6, 6) in [row, col] orientation, without boundary conditions."""
        v = np.ascontiguousarray(np.asarray(blocks, np.float64).transpose(0, 2, 1))   # -> column-major blocks
        assert v.shape == (self.nnzb, 6, 6)
        self._ck(lib().jsso_set_values_host(self.h, _ptr(v), int(apply_bc)))

    def _ck(self, rc, allow=()):
        if rc and rc not in allow:
            raise JssoError(rc, lib().jsso_last_error(self.h).decode())
        return rc

    def close(self):
        if getattr(self, 'h', None):
            lib().jsso_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- symbolic
    def pattern(self):
        rowptr = np.empty(self.n_row + 1, np.int32)
        colidx = np.empty(self.nnzb, np.int32)
        self._ck(lib().jsso_pattern(self.h, _ptr(rowptr), _ptr(colidx)))
        return rowptr, colidx

    def assembly_tasks(self):
        """Host copies of the warp-task lists of the two-kernel assembly (dict of arrays)."""
        cnt = np.zeros(3, np.int32)
        self._ck(lib().jsso_assembly_tasks(self.h, _ptr(cnt), None, None, None, None, None, None))
        out = {'n_task': int(cnt[0]), 'tasks_ok': bool(cnt[2]),
               'task_meta': np.empty((int(cnt[0]), 4), np.int32), 'task_els': np.empty(int(cnt[1]), np.int32),
               'item_desc': np.empty(self.n_items, np.uint16), 'blk_bc': np.empty(self.nnzb, np.uint16),
               'blk_item_ptr': np.empty(self.nnzb + 1, np.int32), 'item_code': np.empty(self.n_items, np.int32)}
        self._ck(lib().jsso_assembly_tasks(self.h, _ptr(cnt), _ptr(out['task_meta']), _ptr(out['task_els']),
                                           _ptr(out['item_desc']), _ptr(out['blk_bc']), _ptr(out['blk_item_ptr']),
                                           _ptr(out['item_code'])))
        return out

    # ---- element matrices
    def quad_ke(self, crds, prop_q, out=None, stream=None):
        out = DeviceArray((self.n_quad, 24, 24)) if out is None else out
        self._ck(lib().jsso_quad_ke(self.h, _dp(crds), _dp(prop_q), _dp(out), stream))
        return out

    def beam_ke(self, crds, prop_b, out=None, stream=None):
        out = DeviceArray((self.n_beam, 12, 12)) if out is None else out
        self._ck(lib().jsso_beam_ke(self.h, _dp(crds), _dp(prop_b), _dp(out), stream))
        return out

    def quad_area(self, crds, out=None, stream=None):
        out = DeviceArray((self.n_quad,)) if out is None else out
        self._ck(lib().jsso_quad_area(self.h, _dp(crds), _dp(out), stream))
        return out

    # ---- assembly
    def assemble(self, crds, prop_q, prop_b, apply_bc=True, stream=None):
        self._ck(lib().jsso_assemble(self.h, _dp(crds), _dp(prop_q), _dp(prop_b), int(apply_bc), stream))

    def profile(self, enable=True):
        self._ck(lib().jsso_profile(self.h, int(enable)))

    def profile_read(self):
        """(quad_geometry_kernel ms, assemble_tasks_kernel ms) of the last assemble() under profile()."""
        ms = np.zeros(2, np.float32)
        self._ck(lib().jsso_profile_read(self.h, _ptr(ms)))
        return float(ms[0]), float(ms[1])

    def assemble_from_ke(self, ke_q, ke_b, apply_bc=True, stream=None):
        self._ck(lib().jsso_assemble_from_ke(self.h, _dp(ke_q), _dp(ke_b), int(apply_bc), stream))

    def values_host(self):
        """(nnzb, 6, 6) blocks in the usual [row, col] orientation (transposed from
        the library's column-major block storage)."""
        v = np.empty((self.nnzb, 6, 6))
        self._ck(lib().jsso_get_values_host(self.h, _ptr(v)))
        return np.ascontiguousarray(v.transpose(0, 2, 1))

    def flags(self):
        f = C.c_int32()
        self._ck(lib().jsso_get_flags(self.h, C.byref(f)))
        return f.value

    # ---- linear algebra
    def spmv(self, x, y, stream=None):
        self._ck(lib().jsso_spmv(self.h, _dp(x), _dp(y), stream))

    def pcg(self, b, x, opts=None, stream=None, allow_noconv=False):
        st = Stats()
        o = opts or make_opts()
        self._ck(lib().jsso_pcg(self.h, _dp(b), _dp(x), C.byref(o), C.byref(st), stream),
                 allow=(JSSO_ERR_NOCONV,) if allow_noconv else ())
        return st

    def mg_setup(self, levels=None, max_coarse_nodes=64):
        """Build (or take) the symbolic smoothed-aggregation hierarchy and upload it."""
        from . import multigrid
        if levels is None:
            rp, ci = self.pattern()
            levels = multigrid.build_hierarchy(rp, ci, max_coarse_nodes=max_coarse_nodes)
        descs = (MgLevelDesc * max(len(levels), 1))()
        keep = []
        for d, lv in zip(descs, levels):
            for k in _MG_INT_FIELDS:
                setattr(d, k, int(lv[k]))
            for k in _MG_PTR_FIELDS:
                a = np.ascontiguousarray(lv[k], dtype=np.int32)
                keep.append(a)
                setattr(d, k, a.ctypes.data)
        self._ck(lib().jsso_mg_setup(self.h, len(levels), descs))
        self.mg_levels = [(lv['n_f'], lv['n_c']) for lv in levels]
        return levels

    def mg_set_dist(self, nccl_id, rank, n_rank, plan):
        """Distribute the multigrid solve of this whole-mesh handle by row ranges (collective; `plan` from
        jaxsso_b200.dist_multigrid.build_plan, computed identically on every rank)."""
        from . import dist_multigrid
        mine = dist_multigrid.rank_plan(plan, rank)
        n_dist = plan['n_dist']
        bounds = np.ascontiguousarray(np.stack(plan['bounds'][:n_dist + 1]), np.int32)
        descs = (MgHaloDesc * max(len(mine), 1))()      # n_dist levels (+ the all-gather plan of the first replicated one)
        keep = []
        for d, lv in zip(descs, mine):
            d.n_peer = int(lv['peer_rank'].shape[0])
            for k in ('peer_rank', 'send_ptr', 'send_idx', 'recv_ptr', 'recv_idx', 'remote_off', 'ghost_rows'):
                a = np.ascontiguousarray(lv[k], dtype=np.int32)
                keep.append(a)
                setattr(d, k, a.ctypes.data)
            d.n_ghost = int(lv['ghost_rows'].shape[0])
        idb = np.frombuffer(bytes(nccl_id), dtype=np.uint8).copy()
        self._ck(lib().jsso_mg_set_dist(self.h, _ptr(idb), int(rank), int(n_rank), int(n_dist), _ptr(bounds),
                                        len(mine), descs))

    def mg_p2p_connect(self, plan, allgather):
        """Switch the distributed multigrid solve to the peer-memory exchange (collective).  `allgather(bytes)` returns
        the list of every rank's bytes in rank order (e.g. torch.distributed.all_gather_object)."""
        from . import dist_multigrid
        common = dist_multigrid.common_max_recv(plan)
        self._ck(lib().jsso_mg_p2p_reserve(self.h, common))
        buf = np.zeros(128, np.uint8)
        self._ck(lib().jsso_mg_p2p_export(self.h, _ptr(buf)))
        blobs = allgather(buf.tobytes())
        blob = np.frombuffer(b''.join(blobs), dtype=np.uint8).copy()
        mr = np.full(plan['n_rank'], common, np.int32)
        self._ck(lib().jsso_mg_p2p_connect(self.h, _ptr(blob), _ptr(mr)))

    def mg_set_dist_setup(self, rowptr, colidx, levels, plan, rank, _drop_ghost=False):
        """Distribute the numeric multigrid setup of the distributed levels as well (after mg_set_dist): this rank
        assembles / scales only the rows it reads and computes only its share of every Galerkin product
        (jaxsso_b200.dist_multigrid.setup_plan)."""
        from . import dist_multigrid
        sp = dist_multigrid.setup_plan(rowptr, colidx, levels, plan, rank, drop_ghost=_drop_ghost)
        descs = (MgSetupDesc * max(len(sp), 1))()
        keep = []
        for d, lv in zip(descs, sp):
            for k in ('p_slots', 'ap_slots', 'ac_bounds'):
                a = np.ascontiguousarray(lv[k], dtype=np.int32)
                keep.append(a)
                setattr(d, k, a.ctypes.data)
            d.n_p_slots, d.n_ap_slots = int(lv['p_slots'].shape[0]), int(lv['ap_slots'].shape[0])
            d.p_own_lo, d.p_own_hi = (int(v) for v in lv['p_own'])
            d.pt_own_lo, d.pt_own_hi = (int(v) for v in lv['pt_own'])
            if 'scale_rows' in lv:
                d.scale_row_lo, d.scale_row_hi = (int(v) for v in lv['scale_rows'])
                d.factor_row_lo, d.factor_row_hi = (int(v) for v in lv['factor_rows'])
        self._ck(lib().jsso_mg_set_dist_setup(self.h, len(sp), descs))
        return sp

    def mg_dist_counters(self):
        out = np.zeros(3, np.int64)
        self._ck(lib().jsso_mg_dist_counters(self.h, _ptr(out)))
        self.mg_dist_p2p = bool(out[2])
        return int(out[0]), int(out[1])

    # ---- adjoint
    def adjoint(self, crds, prop_q, prop_b, u, lam, d_crds=None, d_prop_q=None, d_prop_b=None, stream=None):
        self._ck(lib().jsso_adjoint(self.h, _dp(crds), _dp(prop_q), _dp(prop_b), _dp(u), _dp(lam),
                                    _dp(d_crds), _dp(d_prop_q), _dp(d_prop_b), stream))

    # ---- drop-in pair
    def forward(self, crds, prop_q, prop_b, f, u, opts=None, stream=None):
        st = Stats()
        o = opts or make_opts()
        self._ck(lib().jsso_forward(self.h, _dp(crds), _dp(prop_q), _dp(prop_b), _dp(f), _dp(u),
                                    C.byref(o), C.byref(st), stream))
        return st

    def backward(self, crds, prop_q, prop_b, u, g, d_crds=None, d_prop_q=None, d_prop_b=None, lam=None,
                 opts=None, stream=None):
        st = Stats()
        o = opts or make_opts()
        self._ck(lib().jsso_backward(self.h, _dp(crds), _dp(prop_q), _dp(prop_b), _dp(u), _dp(g),
                                     _dp(d_crds), _dp(d_prop_q), _dp(d_prop_b), _dp(lam),
                                     C.byref(o), C.byref(st), stream))
        return st

    def value_and_grad_host(self, crds, prop_q, prop_b, f, want=('crds', 'prop_q', 'prop_b'), opts=None, u0=None,
                            allow_noconv=False, out=None):
        """End-to-end strain-energy value + gradient with HOST (NumPy) buffers.  ``u0``: initial
        guess for the solve (e.g. the previous design's u); needs opts.use_x0.  ``out``: optional
        ``(u, d_crds, d_prop_q, d_prop_b)`` arrays to fill instead of fresh ones (entries may be None) -- with
        page-locked arrays (``pinned_empty``) for inputs and ``out`` the library copies by DMA straight from / into
        them; pageable arrays are staged through its own pinned buffers (threaded memcpy), which at 1M quads costs
        tens of milliseconds per call, most of it page faults of freshly allocated results."""
        crds = np.ascontiguousarray(crds, np.float64)
        prop_q = np.ascontiguousarray(prop_q, np.float64)
        prop_b = np.ascontiguousarray(prop_b, np.float64)
        f = np.ascontiguousarray(f, np.float64)
        self._check_shapes(crds, prop_q, prop_b, f=f, u0=u0)
        ou, oc, oq, ob = out if out is not None else (None, None, None, None)
        for a, size, name in ((ou, 6 * self.n_node, 'out[0] (u)'), (oc, 3 * self.n_node, 'out[1] (d_crds)'),
                              (oq, 5 * self.n_quad, 'out[2] (d_prop_q)'), (ob, 6 * self.n_beam, 'out[3] (d_prop_b)')):
            if a is not None and (a.size != size or a.dtype != np.float64 or not a.flags['C_CONTIGUOUS']):
                raise ValueError(f'{name}: expected {size} contiguous float64 values')
        if ou is not None:
            u = ou
            if u0 is not None and u0 is not ou:
                u.ravel()[:] = np.asarray(u0, np.float64).ravel()
        else:
            u = np.empty(6 * self.n_node) if u0 is None else np.array(u0, dtype=np.float64).ravel()
        dc = (oc if oc is not None else np.empty((self.n_node, 3))) if 'crds' in want else None
        dq = (oq if oq is not None else np.empty((self.n_quad, 5))) if ('prop_q' in want and self.n_quad) else None
        db = (ob if ob is not None else np.empty((self.n_beam, 6))) if ('prop_b' in want and self.n_beam) else None
        val = C.c_double()
        fs, bs = Stats(), Stats()
        o = opts or make_opts()
        if u0 is not None:
            o.use_x0 = 1
        rc = self._ck(lib().jsso_value_and_grad_host(self.h, _ptr(crds), _ptr(prop_q), _ptr(prop_b), _ptr(f),
                                                     C.byref(val), _ptr(u), _ptr(dc), _ptr(dq), _ptr(db),
                                                     C.byref(o), C.byref(fs), C.byref(bs)),
                      allow=(JSSO_ERR_NOCONV,) if allow_noconv else ())
        if rc == JSSO_ERR_NOCONV:
            # the solver stopped at its attainable accuracy (or maxiter): u and the gradients of the best iterate are
            # returned, as the reference's direct solvers return whatever accuracy they reach; fs.relres says how far
            import warnings
            warnings.warn('jaxsso_b200: ' + lib().jsso_last_error(self.h).decode(), RuntimeWarning, stacklevel=2)
        return val.value, u, dc, dq, db, fs, bs

    def _check_shapes(self, crds, prop_q, prop_b, f=None, u0=None, u=None, lam=None):
        """The C entry points copy n_node / n_quad / n_beam sized blocks from these pointers: refuse anything else."""
        def need(a, size, name):
            if a is None:
                if size:
                    raise ValueError(f'{name} is required ({size} values)')
                return
            a = np.asarray(a)
            if a.size != size or a.dtype != np.float64 or not a.flags['C_CONTIGUOUS']:
                raise ValueError(f'{name}: expected {size} contiguous float64 values, got {a.size} of {a.dtype}')
        need(crds, 3 * self.n_node, 'crds')
        need(prop_q, 5 * self.n_quad, 'prop_q')
        need(prop_b, 6 * self.n_beam, 'prop_b')
        for a, name in ((f, 'f'), (u0, 'u0'), (u, 'u'), (lam, 'lam')):
            if a is not None:
                need(a, 6 * self.n_node, name)

    def assemble_adjoint_host(self, crds, prop_q, prop_b, u, lam, out=None):
        """Ke+assembly+adjoint with HOST buffers; `out` = (d_crds, d_prop_q, d_prop_b) to reuse."""
        self._check_shapes(crds, prop_q, prop_b, u=u, lam=lam)
        if out is None:
            out = (np.empty((self.n_node, 3)), np.empty((self.n_quad, 5)) if self.n_quad else None,
                   np.empty((self.n_beam, 6)) if self.n_beam else None)
        self._ck(lib().jsso_assemble_adjoint_host(self.h, _ptr(crds), _ptr(prop_q), _ptr(prop_b), _ptr(u),
                                                  _ptr(lam), _ptr(out[0]), _ptr(out[1]), _ptr(out[2])))
        return out

    # ---- multi-GPU
    def set_halo(self, nccl_id, rank, n_rank, peer_rank, send_ptr, send_idx, recv_start, recv_count):
        a = [np.ascontiguousarray(x, np.int32) for x in (peer_rank, send_ptr, send_idx, recv_start, recv_count)]
        idb = np.frombuffer(bytes(nccl_id), dtype=np.uint8).copy()
        self._ck(lib().jsso_set_halo(self.h, _ptr(idb), rank, n_rank, len(a[0]), *[_ptr(x) for x in a]))

    def p2p_export(self):
        buf = np.zeros(128, np.uint8)
        self._ck(lib().jsso_p2p_export(self.h, _ptr(buf)))
        return buf.tobytes()

    def p2p_connect(self, all_handles, remote_start):
        blob = np.frombuffer(b''.join(all_handles), dtype=np.uint8).copy()
        rs = np.ascontiguousarray(remote_start, np.int32)
        self._ck(lib().jsso_p2p_connect(self.h, _ptr(blob), _ptr(rs)))

    def halo_exchange(self, vec, stream=None):
        self._ck(lib().jsso_halo_exchange(self.h, _dp(vec), stream))


def gather_rows(src, idx, width, out=None, stream=None):
    """out[i, :] = src[idx[i], :] on the device (src, idx, out: DeviceArray; idx int32)."""
    n = int(np.prod(idx.shape))
    out = DeviceArray((n * width,)) if out is None else out
    rc = lib().jsso_gather_rows(src.ptr, idx.ptr, n, width, out.ptr, stream)
    if rc:
        raise JssoError(rc, 'jsso_gather_rows')
    return out


def fp64_peak(device=0, seconds=1.0):
    """Measured FP64 DFMA peak in TFLOP/s (jsso_fp64_peak)."""
    tf, sec = C.c_double(), C.c_double()
    rc = lib().jsso_fp64_peak(device, seconds, C.byref(tf), C.byref(sec))
    if rc:
        raise JssoError(rc, 'jsso_fp64_peak')
    return float(tf.value), float(sec.value)


def nccl_unique_id():
    buf = np.zeros(128, np.uint8)
    rc = lib().jsso_nccl_unique_id(_ptr(buf))
    if rc:
        raise JssoError(rc, lib().jsso_last_error(None).decode())
    return buf.tobytes()


def bsr_to_scipy(rowptr, colidx, blocks, n_col_nodes=None):
    """scipy.sparse.bsr_matrix from pattern + (nnzb,6,6) [row,col]-oriented blocks."""
    import scipy.sparse as sp
    n_row = rowptr.shape[0] - 1
    n_col = n_row if n_col_nodes is None else n_col_nodes
    return sp.bsr_matrix((blocks, colidx, rowptr), shape=(6 * n_row, 6 * n_col))
