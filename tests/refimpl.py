"""ctypes wrappers of the checkers under oracle/ (TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this).

  RefTree  -- the real reference (oracle/_ref/libref_simplex.so, built from /root/reference
              sources in place by oracle/Makefile); travels to the GPU box as a built .so.
  port     -- the plain-C restatement oracle/simplex_oracle.c (liboracle_port.so).
"""
import ctypes
import os
from ctypes import c_double, c_int, c_long, c_ulong, c_void_p

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_simplex.so")
REF_CHECKED_SO = os.path.join(ROOT, "oracle", "_ref", "libref_simplex_checked.so")
PORT_SO = os.path.join(ROOT, "oracle", "liboracle_port.so")


def have_ref():
    return os.path.exists(REF_SO)


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _load_ref(path):
    L = ctypes.CDLL(path)
    L.ref_build.restype = c_void_p
    L.ref_build.argtypes = [c_void_p, c_int, c_int, c_int, c_ulong]
    L.ref_free.argtypes = [c_void_p]
    L.ref_build_seconds.restype = c_double
    L.ref_build_seconds.argtypes = [c_void_p]
    L.ref_n_nodes.argtypes = [c_void_p]
    L.ref_n_points.argtypes = [c_void_p]
    L.ref_export.argtypes = [c_void_p] * 8
    L.ref_eval.restype = c_long
    L.ref_eval.argtypes = [c_void_p, c_void_p, c_void_p, c_long, c_void_p, c_int, c_void_p, c_void_p]
    L.ref_eval_fork.restype = c_double
    L.ref_eval_fork.argtypes = [c_void_p, c_void_p, c_void_p, c_long, c_void_p, c_int, c_void_p,
                                c_void_p, c_int]
    L.ref_eval_raster.restype = c_long
    L.ref_eval_raster.argtypes = [c_void_p, c_double, c_double, c_long, c_double, c_double, c_long,
                                  c_void_p, c_int, c_void_p, c_void_p]
    L.ref_weather_data.argtypes = [c_void_p]
    if hasattr(L, "ref_eval_deriv"):
        L.ref_eval_deriv.restype = c_long
        L.ref_eval_deriv.argtypes = [c_void_p, c_void_p, c_void_p, c_long, c_void_p, c_int, c_void_p, c_void_p,
                                     c_void_p]
    return L


_REF = {}


def ref_lib(checked=False):
    path = REF_CHECKED_SO if checked else REF_SO
    if path not in _REF:
        _REF[path] = _load_ref(path)
        if checked:
            _REF[path].ref_check.argtypes = [c_void_p]
    return _REF[path]


def weather_data():
    w = np.zeros(150)
    ref_lib().ref_weather_data(_p(w))
    return w.reshape(50, 3)


class RefTree:
    """simplex_tree built by the reference's own simplex_tree_init."""

    def __init__(self, xa, ya, flags=0, seed=None, checked=False):
        self.L = ref_lib(checked)
        self.xa = np.ascontiguousarray(xa, dtype=np.float64)
        self.ya = np.ascontiguousarray(ya, dtype=np.float64)
        self.n = self.xa.size
        xy = np.ascontiguousarray(np.stack([self.xa, self.ya], axis=1)) if self.n else np.zeros((1, 2))
        self.h = self.L.ref_build(_p(xy), self.n, flags, 0 if seed is None else 1,
                                  0 if seed is None else seed)
        if not self.h:
            raise RuntimeError("reference simplex_tree_init failed")
        self.build_seconds = self.L.ref_build_seconds(self.h)
        self.n_simplexes = self.L.ref_n_nodes(self.h)

    def close(self):
        if self.h:
            self.L.ref_free(self.h)
            self.h = None

    def export(self):
        N = self.n_simplexes
        pts3 = np.empty(3 * N, dtype=np.int32)
        lnk3 = np.empty(3 * N, dtype=np.int32)
        typ = np.empty(N, dtype=np.int32)
        shuffle = np.empty(max(self.n, 1), dtype=np.int64)
        shift, scale, seed = np.empty(2), np.empty(2), np.empty(6)
        self.L.ref_export(self.h, _p(pts3), _p(lnk3), _p(typ), _p(shuffle), _p(shift), _p(scale),
                          _p(seed))
        return dict(pts3=pts3, lnk3=lnk3, type=typ, shuffle=shuffle[:self.n].astype(np.uint64),
                    shift=shift, scale=scale, seed_points=seed)

    def eval(self, xq, yq, za, nproc=1):
        """find_leaf + interp_point loop; za field-major (F, n). Returns (leaf, val[F,nq], n_outside)."""
        xq = np.ascontiguousarray(xq, dtype=np.float64)
        yq = np.ascontiguousarray(yq, dtype=np.float64)
        za = np.ascontiguousarray(np.atleast_2d(za), dtype=np.float64)
        F, nq = za.shape[0], xq.size
        leaf = np.empty(nq, dtype=np.int32)
        val = np.empty((F, nq), dtype=np.float64)
        if nproc > 1:
            dt = self.L.ref_eval_fork(self.h, _p(xq), _p(yq), nq, _p(za), F, _p(leaf), _p(val), nproc)
            if dt < 0:
                raise RuntimeError("ref_eval_fork failed")
            return leaf, val, int((leaf < 0).sum())
        out = self.L.ref_eval(self.h, _p(xq), _p(yq), nq, _p(za), F, _p(leaf), _p(val))
        return leaf, val, int(out)

    def eval_deriv(self, xq, yq, za):
        """find_leaf + gradient of every field (oracle-side definition on top of the reference's
        calculate_bary_coords + gsl_linalg_LU_solve).  Returns (leaf, dzdx[F,nq], dzdy[F,nq])."""
        xq = np.ascontiguousarray(xq, dtype=np.float64)
        yq = np.ascontiguousarray(yq, dtype=np.float64)
        za = np.ascontiguousarray(np.atleast_2d(za), dtype=np.float64)
        F, nq = za.shape[0], xq.size
        leaf = np.empty(nq, dtype=np.int32)
        dx = np.empty((F, nq), dtype=np.float64)
        dy = np.empty((F, nq), dtype=np.float64)
        self.L.ref_eval_deriv(self.h, _p(xq), _p(yq), nq, _p(za), F, _p(leaf), _p(dx), _p(dy))
        return leaf, dx, dy

    def time_eval(self, xq, yq, za, nproc):
        xq = np.ascontiguousarray(xq, dtype=np.float64)
        yq = np.ascontiguousarray(yq, dtype=np.float64)
        za = np.ascontiguousarray(np.atleast_2d(za), dtype=np.float64)
        return self.L.ref_eval_fork(self.h, _p(xq), _p(yq), xq.size, _p(za), za.shape[0], None, None,
                                    nproc)

    def eval_raster(self, x0, dx, nx, y0, dy, ny, za):
        za = np.ascontiguousarray(np.atleast_2d(za), dtype=np.float64)
        F = za.shape[0]
        leaf = np.empty((nx, ny), dtype=np.int32)
        val = np.empty((F, nx, ny), dtype=np.float64)
        out = self.L.ref_eval_raster(self.h, x0, dx, nx, y0, dy, ny, _p(za), F, _p(leaf), _p(val))
        return leaf, val, int(out)

    def check(self):
        return self.L.ref_check(self.h)


# ---- the plain-C restatement (oracle/simplex_oracle.c) -------------------------------------
_PORT = None


def port_lib():
    global _PORT
    if _PORT is None:
        L = ctypes.CDLL(PORT_SO)
        L.so_tree_create.restype = c_void_p
        L.so_tree_create.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, ctypes.c_size_t, c_int]
        L.so_tree_free.argtypes = [c_void_p]
        L.so_eval.restype = c_long
        L.so_eval.argtypes = [c_void_p, c_void_p, c_void_p, c_long, c_void_p, c_int, c_void_p, c_void_p]
        L.so_eval_deriv.restype = c_long
        L.so_eval_deriv.argtypes = [c_void_p, c_void_p, c_void_p, c_long, c_void_p, c_int, c_void_p, c_void_p,
                                    c_void_p]
        L.so_eval_mt.restype = c_long
        L.so_eval_mt.argtypes = [c_void_p, c_void_p, c_void_p, c_long, c_void_p, c_int, c_void_p, c_void_p, c_int]
        L.so_eval_stats.restype = c_long
        L.so_eval_stats.argtypes = [c_void_p, c_void_p, c_void_p, c_long, c_void_p, c_void_p]
        _PORT = L
    return _PORT


class PortTree:
    """The oracle port evaluated on flattened tree arrays (dict as returned by
    RefTree.export() / ScatteredCuda.get_tree())."""

    def __init__(self, tree, xa, ya):
        self.L = port_lib()
        xa = np.ascontiguousarray(xa, dtype=np.float64)
        ya = np.ascontiguousarray(ya, dtype=np.float64)
        self.n = xa.size
        data = np.ascontiguousarray(np.stack([xa, ya], axis=1)) if self.n else np.zeros((1, 2))
        a = {k: np.ascontiguousarray(v) for k, v in tree.items()}
        self.h = self.L.so_tree_create(a["type"].size, _p(a["pts3"].astype(np.int32)),
                                       _p(a["lnk3"].astype(np.int32)), _p(a["type"].astype(np.int32)),
                                       _p(a["shuffle"].astype(np.uint64)), _p(a["shift"]), _p(a["scale"]),
                                       _p(a["seed_points"]), _p(data), 2, self.n)

    def close(self):
        if self.h:
            self.L.so_tree_free(self.h)
            self.h = None

    def eval(self, xq, yq, za):
        xq = np.ascontiguousarray(xq, dtype=np.float64)
        yq = np.ascontiguousarray(yq, dtype=np.float64)
        za = np.ascontiguousarray(np.atleast_2d(za), dtype=np.float64)
        F, nq = za.shape[0], xq.size
        leaf = np.empty(nq, dtype=np.int32)
        val = np.empty((F, nq), dtype=np.float64)
        out = self.L.so_eval(self.h, _p(xq), _p(yq), nq, _p(za), F, _p(leaf), _p(val))
        return leaf, val, int(out)

    def eval_mt(self, xq, yq, za, n_threads):
        """the thread-safe multi-thread evaluation: n_threads POSIX threads on
        one shared tree"""
        xq = np.ascontiguousarray(xq, dtype=np.float64)
        yq = np.ascontiguousarray(yq, dtype=np.float64)
        za = np.ascontiguousarray(np.atleast_2d(za), dtype=np.float64)
        F, nq = za.shape[0], xq.size
        leaf = np.empty(nq, dtype=np.int32)
        val = np.empty((F, nq), dtype=np.float64)
        out = self.L.so_eval_mt(self.h, _p(xq), _p(yq), nq, _p(za), F, _p(leaf), _p(val), int(n_threads))
        return leaf, val, int(out)

    def eval_deriv(self, xq, yq, za):
        xq = np.ascontiguousarray(xq, dtype=np.float64)
        yq = np.ascontiguousarray(yq, dtype=np.float64)
        za = np.ascontiguousarray(np.atleast_2d(za), dtype=np.float64)
        F, nq = za.shape[0], xq.size
        leaf = np.empty(nq, dtype=np.int32)
        dx = np.empty((F, nq), dtype=np.float64)
        dy = np.empty((F, nq), dtype=np.float64)
        self.L.so_eval_deriv(self.h, _p(xq), _p(yq), nq, _p(za), F, _p(leaf), _p(dx), _p(dy))
        return leaf, dx, dy

    def stats(self, xq, yq):
        xq = np.ascontiguousarray(xq, dtype=np.float64)
        yq = np.ascontiguousarray(yq, dtype=np.float64)
        depth = np.empty(xq.size, dtype=np.int32)
        tests = np.empty(xq.size, dtype=np.int32)
        self.L.so_eval_stats(self.h, _p(xq), _p(yq), xq.size, _p(depth), _p(tests))
        return depth, tests


def same_bits(a, b):
    """bit-identical float arrays (NaN == NaN)"""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))).all())
