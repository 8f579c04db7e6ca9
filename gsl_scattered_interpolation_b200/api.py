"""Host-side mirror of the C ABI, for tests and the benchmark driver.

The product is the C library (``include/gsl_scattered_cuda.h``); this module only wraps its
entry points one-to-one (same names, argument meaning and status codes) so that the parity
tests read like the reference's own driver (scattered_interp_example.c): ``alloc`` /
``init`` from x, y, z nodes / ``eval`` / ``free``.  numpy arrays are host buffers, integer
addresses (e.g. ``torch.Tensor.data_ptr()``) are device buffers.
"""
import ctypes
import os

import numpy as np

from . import _lib
from ._lib import (DEFAULT, EBADLEN, EDOM, EFAILED, EINVAL, ENOMEM, ISOSCALE, NOSTANDARDIZE,
                   SUCCESS, Info)

__all__ = ["ScatteredCuda", "GslError", "shuffle_mt19937", "device_count", "generate_uniform_host", "type_descriptor",
           "SUCCESS", "EDOM", "EINVAL", "EFAILED", "ENOMEM", "EBADLEN", "DEFAULT",
           "NOSTANDARDIZE", "ISOSCALE"]


class GslError(RuntimeError):
    def __init__(self, status, what):
        super().__init__(f"{what}: status {status}")
        self.status = status


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def device_count():
    return _lib.load().gsl_scattered_cuda_device_count()


def shuffle_mt19937(seed, n):
    """tree->shuffle after gsl_ran_shuffle with gsl_rng_mt19937 seeded `seed`."""
    order = np.zeros(n, dtype=np.uint64)
    st = _lib.load().gsl_scattered_cuda_shuffle_mt19937(seed, n, _ptr(order))
    if st:
        raise GslError(st, "shuffle_mt19937")
    return order


def type_descriptor():
    """the gsl_scattered_cuda_type of the library (cf. gsl_interp2d_bilinear)"""
    p = _lib.load().gsl_scattered_cuda_type_linear_simplex()
    return ctypes.cast(p, ctypes.POINTER(_lib.TypeDescriptor)).contents


def generate_uniform_host(seed, first, nq, xlo, xhi, ylo, yhi):
    xq = np.empty(nq, dtype=np.float64)
    yq = np.empty(nq, dtype=np.float64)
    _lib.load().gsl_scattered_cuda_generate_uniform_host(seed, first, nq, xlo, xhi, ylo, yhi,
                                                         _ptr(xq), _ptr(yq))
    return xq, yq


class ScatteredCuda:
    """gsl_scattered_cuda workspace (alloc on construction, free on close / GC)."""

    def __init__(self, n_nodes, n_fields=1):
        self._L = _lib.load()
        self.n_fields = int(n_fields)
        self._w = self._L.gsl_scattered_cuda_alloc(int(n_nodes), int(n_fields))
        if not self._w:
            raise GslError(ENOMEM, "gsl_scattered_cuda_alloc")

    # -- life cycle -------------------------------------------------------------------------
    def free(self):
        if getattr(self, "_w", None):
            self._L.gsl_scattered_cuda_free(self._w)
            self._w = None

    close = free

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.free()

    def _check(self, st, what, allow=()):
        if st != SUCCESS and st not in allow:
            raise GslError(st, what)
        return st

    # -- configuration ----------------------------------------------------------------------
    def set_devices(self, devs):
        arr = (ctypes.c_int * len(devs))(*devs)
        self._check(self._L.gsl_scattered_cuda_set_devices(self._w, arr, len(devs)), "set_devices")

    def set_param(self, name, value):
        self._check(self._L.gsl_scattered_cuda_set_param(self._w, name.encode(), float(value)),
                    f"set_param({name})")

    # -- init -------------------------------------------------------------------------------
    @staticmethod
    def _order(order):
        return None if order is None else np.ascontiguousarray(order, dtype=np.uint64)

    def build(self, xa, ya, flags=DEFAULT, order=None):
        xa, ya, order = _f64(xa), _f64(ya), self._order(order)
        self._check(self._L.gsl_scattered_cuda_build(self._w, _ptr(xa), _ptr(ya), xa.size, flags,
                                                     _ptr(order)), "build")

    def upload(self, za):
        za = _f64(za)
        self._check(self._L.gsl_scattered_cuda_upload(self._w, _ptr(za), self.n_fields), "upload")

    def init(self, xa, ya, za, flags=DEFAULT, order=None):
        xa, ya, za, order = _f64(xa), _f64(ya), _f64(za), self._order(order)
        self._check(self._L.gsl_scattered_cuda_init(self._w, _ptr(xa), _ptr(ya), _ptr(za), xa.size,
                                                    flags, _ptr(order)), "init")

    def init_from_tree(self, pts3, lnk3, node_type, shuffle, shift, scale, seed_points, data, za):
        pts3 = np.ascontiguousarray(pts3, dtype=np.int32)
        lnk3 = np.ascontiguousarray(lnk3, dtype=np.int32)
        node_type = np.ascontiguousarray(node_type, dtype=np.int32)
        shuffle = np.ascontiguousarray(shuffle, dtype=np.uint64)
        shift, scale, seed_points = _f64(shift), _f64(scale), _f64(seed_points)
        data, za = _f64(data), _f64(za)
        n = data.shape[0]
        self._check(self._L.gsl_scattered_cuda_init_from_tree(
            self._w, node_type.size, _ptr(pts3), _ptr(lnk3), _ptr(node_type), _ptr(shuffle),
            _ptr(shift), _ptr(scale), _ptr(seed_points), _ptr(data), data.shape[1] if n else 2, n,
            _ptr(za), self.n_fields), "init_from_tree")

    # -- eval (host buffers) ----------------------------------------------------------------
    def eval_e(self, xq, yq, want_values=True, want_leaf=True, out_z=None, out_leaf=None):
        """Returns (status, zq[n_fields, nq] | None, leaf[nq] | None); status is SUCCESS or EDOM."""
        xq, yq = _f64(xq), _f64(yq)
        nq = xq.size
        zq = out_z if out_z is not None else (
            np.empty((self.n_fields, nq), dtype=np.float64) if want_values else None)
        leaf = out_leaf if out_leaf is not None else (
            np.empty(nq, dtype=np.int32) if want_leaf else None)
        st = self._L.gsl_scattered_cuda_eval_e(self._w, _ptr(xq), _ptr(yq), nq, _ptr(zq),
                                               _ptr(leaf))
        self._check(st, "eval_e", allow=(EDOM,))
        return st, zq, leaf

    def eval_ptr_e(self, xq_addr, yq_addr, nq, zq_addr, leaf_addr):
        """eval_e on raw host addresses (e.g. pinned memory from host_alloc)."""
        st = self._L.gsl_scattered_cuda_eval_e(self._w, xq_addr, yq_addr, nq, zq_addr, leaf_addr)
        return self._check(st, "eval_e", allow=(EDOM,))

    def eval(self, x, y):
        return self._L.gsl_scattered_cuda_eval(self._w, float(x), float(y))

    def eval_raster_e(self, x0, dx, nx, y0, dy, ny, want_values=True, want_leaf=True):
        zq = np.empty((self.n_fields, nx, ny), dtype=np.float64) if want_values else None
        leaf = np.empty((nx, ny), dtype=np.int32) if want_leaf else None
        st = self._L.gsl_scattered_cuda_eval_raster_e(self._w, x0, dx, nx, y0, dy, ny, _ptr(zq),
                                                      _ptr(leaf))
        self._check(st, "eval_raster_e", allow=(EDOM,))
        return st, zq, leaf

    # -- derivatives (gsl_interp2d's eval_deriv_x / eval_deriv_y for the scattered method) ----
    def eval_deriv_e(self, xq, yq, want_x=True, want_y=True):
        """Returns (status, dzdx[n_fields, nq] | None, dzdy[n_fields, nq] | None)."""
        xq, yq = _f64(xq), _f64(yq)
        nq = xq.size
        dx = np.empty((self.n_fields, nq), dtype=np.float64) if want_x else None
        dy = np.empty((self.n_fields, nq), dtype=np.float64) if want_y else None
        st = self._L.gsl_scattered_cuda_eval_deriv_e(self._w, _ptr(xq), _ptr(yq), nq, _ptr(dx), _ptr(dy))
        self._check(st, "eval_deriv_e", allow=(EDOM,))
        return st, dx, dy

    def eval_deriv2_e(self, xq, yq, which="xx"):
        xq, yq = _f64(xq), _f64(yq)
        d = np.empty((self.n_fields, xq.size), dtype=np.float64)
        fn = getattr(self._L, f"gsl_scattered_cuda_eval_deriv_{which}_e")
        st = fn(self._w, _ptr(xq), _ptr(yq), xq.size, _ptr(d))
        self._check(st, f"eval_deriv_{which}_e", allow=(EDOM,))
        return st, d

    def eval_deriv_x(self, x, y):
        return self._L.gsl_scattered_cuda_eval_deriv_x(self._w, float(x), float(y))

    def eval_deriv_y(self, x, y):
        return self._L.gsl_scattered_cuda_eval_deriv_y(self._w, float(x), float(y))

    def eval_deriv_device_e(self, dev, d_xq, d_yq, nq, d_dzdx, d_dzdy, ld, d_leaf=None, stream=None):
        st = self._L.gsl_scattered_cuda_eval_deriv_device_e(self._w, dev, d_xq, d_yq, nq, d_dzdx, d_dzdy, ld,
                                                            d_leaf, stream)
        return self._check(st, "eval_deriv_device_e", allow=(EDOM,))

    def name(self):
        return self._L.gsl_scattered_cuda_name(self._w).decode()

    def min_size(self):
        return int(self._L.gsl_scattered_cuda_min_size(self._w))

    # -- eval (device buffers) --------------------------------------------------------------
    def eval_device_e(self, dev, d_xq, d_yq, nq, d_zq, ld, d_leaf, stream=None):
        st = self._L.gsl_scattered_cuda_eval_device_e(self._w, dev, d_xq, d_yq, nq, d_zq, ld,
                                                      d_leaf, stream)
        return self._check(st, "eval_device_e", allow=(EDOM,))

    def eval_raster_device_e(self, dev, x0, dx, nx, y0, dy, ny, d_zq, d_leaf, stream=None):
        st = self._L.gsl_scattered_cuda_eval_raster_device_e(self._w, dev, x0, dx, nx, y0, dy, ny,
                                                             d_zq, d_leaf, stream)
        return self._check(st, "eval_raster_device_e", allow=(EDOM,))

    # -- helpers ----------------------------------------------------------------------------
    def save(self, path):
        """Write the flattened triangulation (slab image) to `path` (gsl_scattered_cuda_save)."""
        self._check(self._L.gsl_scattered_cuda_save(self._w, os.fsencode(path)), "save")

    def load(self, path):
        """Restore a slab image written by save(): no host build (gsl_scattered_cuda_load)."""
        self._check(self._L.gsl_scattered_cuda_load(self._w, os.fsencode(path)), "load")

    def info(self):
        inf = Info()
        self._check(self._L.gsl_scattered_cuda_get_info(self._w, ctypes.byref(inf)), "get_info")
        return {name: getattr(inf, name) for name, _ in Info._fields_}

    def reset_counters(self):
        self._check(self._L.gsl_scattered_cuda_reset_counters(self._w), "reset_counters")

    def get_tree(self, n_simplexes, n_nodes):
        pts3 = np.empty(3 * n_simplexes, dtype=np.int32)
        lnk3 = np.empty(3 * n_simplexes, dtype=np.int32)
        typ = np.empty(n_simplexes, dtype=np.int32)
        shuffle = np.empty(n_nodes, dtype=np.uint64)
        shift, scale, seed = np.empty(2), np.empty(2), np.empty(6)
        self._check(self._L.gsl_scattered_cuda_get_tree(self._w, _ptr(pts3), _ptr(lnk3), _ptr(typ),
                                                        _ptr(shuffle), _ptr(shift), _ptr(scale),
                                                        _ptr(seed)), "get_tree")
        return dict(pts3=pts3, lnk3=lnk3, type=typ, shuffle=shuffle, shift=shift, scale=scale,
                    seed_points=seed)

    def n_simplexes(self):
        inf = Info()
        self._L.gsl_scattered_cuda_get_info(self._w, ctypes.byref(inf))
        return inf.n_simplexes

    def slab_bytes(self):
        return self._L.gsl_scattered_cuda_slab_bytes(self._w)

    def slab_export_device(self, dev, d_dst, stream=None):
        self._check(self._L.gsl_scattered_cuda_slab_export_device(self._w, dev, d_dst, stream),
                    "slab_export_device")

    def init_from_slab_device(self, d_src, nbytes):
        self._check(self._L.gsl_scattered_cuda_init_from_slab_device(self._w, d_src, nbytes),
                    "init_from_slab_device")

    def generate_uniform_device(self, dev, seed, first, nq, xlo, xhi, ylo, yhi, d_xq, d_yq,
                                stream=None):
        self._check(self._L.gsl_scattered_cuda_generate_uniform_device(
            self._w, dev, seed, first, nq, xlo, xhi, ylo, yhi, d_xq, d_yq, stream),
            "generate_uniform_device")

    def checksum_device(self, dev, d_leaf, d_val, n, stream=None):
        out = (ctypes.c_ulonglong * 2)()
        self._check(self._L.gsl_scattered_cuda_checksum_device(self._w, dev, d_leaf, d_val, n, out,
                                                               stream), "checksum_device")
        return int(out[0]), int(out[1])

    def build_stats(self):
        """how much of the last build() went through the batched device locate"""
        sec, pts, rounds = ctypes.c_double(0), ctypes.c_size_t(0), ctypes.c_size_t(0)
        self._check(self._L.gsl_scattered_cuda_build_stats(self._w, ctypes.byref(sec), ctypes.byref(pts),
                                                           ctypes.byref(rounds)), "build_stats")
        return {"batched_seconds": sec.value, "batched_points": pts.value, "rounds": rounds.value}

    def walk_trace(self, dev):
        """(n_ctas, 4) uint64 array {smid, start ns, end ns, queries} of the last fast-path launch
        (set_param("walk_trace", 1) first)"""
        import numpy as np
        out = np.zeros(4 * 16 * 1024, dtype=np.uint64)
        n = ctypes.c_size_t(0)
        self._check(self._L.gsl_scattered_cuda_debug_walk_trace(
            self._w, dev, out.ctypes.data_as(ctypes.c_void_p), out.size, ctypes.byref(n)), "walk_trace")
        return out[:n.value].reshape(-1, 4)

    def compare_device(self, dev, d_leaf_a, d_leaf_b, d_val_a, d_val_b, n, n_val, stream=None):
        out = (ctypes.c_ulonglong * 2)()
        self._check(self._L.gsl_scattered_cuda_compare_device(
            self._w, dev, d_leaf_a, d_leaf_b, d_val_a, d_val_b, n, n_val, out, stream),
            "compare_device")
        return int(out[0]), int(out[1])
