"""ctypes wrapper of tests/hostsim/libhostsim.so (kernel-logic simulator, tests only)."""
import ctypes
import os
from ctypes import c_double, c_int, c_size_t, c_void_p

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_H = None


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def lib():
    global _H
    if _H is None:
        H = ctypes.CDLL(os.path.join(ROOT, "tests", "hostsim", "libhostsim.so"))
        H.hostsim_create.restype = c_void_p
        H.hostsim_create.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_int, c_void_p,
                                     c_double, c_double, c_int, c_int]
        H.hostsim_tiles.restype = c_void_p
        H.hostsim_tiles.argtypes = [c_void_p]
        H.hostsim_tile_bits.restype = c_int
        H.hostsim_tile_bits.argtypes = [c_void_p]
        H.hostsim_tile_max_leaves.restype = c_int
        H.hostsim_tile_max_leaves.argtypes = [c_void_p]
        H.hostsim_entry.restype = c_void_p
        H.hostsim_entry.argtypes = [c_void_p]
        H.hostsim_node_is_leaf.restype = c_int
        H.hostsim_node_is_leaf.argtypes = [c_void_p, c_int]
        H.hostsim_free.argtypes = [c_void_p]
        H.hostsim_eval.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_int,
                                   c_int, c_void_p]
        H.hostsim_deriv.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]
        H.hostsim_n_leaves.restype = c_size_t
        H.hostsim_n_leaves.argtypes = [c_void_p]
        H.hostsim_grid_g.restype = c_size_t
        H.hostsim_seed_escapes.restype = c_size_t
        H.hostsim_seed_escapes.argtypes = [c_void_p]
        H.hostsim_seed_shift.restype = c_int
        H.hostsim_memo_hits.restype = ctypes.c_longlong
        H.hostsim_memo_hits.argtypes = [c_void_p]
        H.hostsim_seed_shift.argtypes = [c_void_p]
        H.hostsim_grid_g.argtypes = [c_void_p]
        H.hostsim_build_batched_equal.restype = c_int
        H.hostsim_build_batched_equal.argtypes = [c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_void_p]
        _H = H
    return _H


def build_batched_equal(xa, ya, flags=0, order=None):
    """(identical?, points located through the batched protocol, rounds)"""
    xa = np.ascontiguousarray(xa, dtype=np.float64)
    ya = np.ascontiguousarray(ya, dtype=np.float64)
    order = None if order is None else np.ascontiguousarray(order, dtype=np.uint64)
    bp, rd = ctypes.c_size_t(0), ctypes.c_size_t(0)
    r = lib().hostsim_build_batched_equal(_p(xa), _p(ya), xa.size, flags, _p(order), ctypes.byref(bp), ctypes.byref(rd))
    if r < 0:
        raise RuntimeError("build failed (%d)" % r)
    return bool(r), bp.value, rd.value


class HostSim:
    def __init__(self, xa, ya, za, flags=0, order=None, cells_per_leaf=0.0, delta=0.0, exact_entry=1,
                 tile_leaves=0):
        self.H = lib()
        xa = np.ascontiguousarray(xa, dtype=np.float64)
        ya = np.ascontiguousarray(ya, dtype=np.float64)
        za = np.ascontiguousarray(np.atleast_2d(za), dtype=np.float64)
        self.F = za.shape[0]
        order = None if order is None else np.ascontiguousarray(order, dtype=np.uint64)
        self.h = self.H.hostsim_create(_p(xa), _p(ya), _p(za), xa.size, self.F, flags, _p(order),
                                       cells_per_leaf, delta, exact_entry, tile_leaves)
        if not self.h:
            raise RuntimeError("hostsim_create failed")

    def entry(self):
        """per-cell start node of the exact replay (what k_entry_cells computes on the device)"""
        g = self.H.hostsim_grid_g(self.h)
        buf = (ctypes.c_int32 * (g * g)).from_address(self.H.hostsim_entry(self.h))
        return np.frombuffer(buf, dtype=np.int32).copy()

    def tiles(self):
        """(tile_bits, first leaf slot per Morton-prefix tile [2^bits + 1], largest tile)"""
        bits = self.H.hostsim_tile_bits(self.h)
        buf = (ctypes.c_int32 * ((1 << bits) + 1)).from_address(self.H.hostsim_tiles(self.h))
        return bits, np.frombuffer(buf, dtype=np.int32).copy(), self.H.hostsim_tile_max_leaves(self.h)

    def seed_table(self):
        """(grid side, log2 of the seed-block width, cells that escape to the int32 table)"""
        return (self.H.hostsim_grid_g(self.h), self.H.hostsim_seed_shift(self.h), self.H.hostsim_seed_escapes(self.h))

    def memo_hits(self):
        """deferred queries answered from the vertex memo so far"""
        return int(self.H.hostsim_memo_hits(self.h))

    def node_is_leaf(self, node):
        return bool(self.H.hostsim_node_is_leaf(self.h, int(node)))

    def close(self):
        if self.h:
            self.H.hostsim_free(self.h)
            self.h = None

    def eval(self, xq, yq, mode=0, max_steps=256):
        xq = np.ascontiguousarray(xq, dtype=np.float64)
        yq = np.ascontiguousarray(yq, dtype=np.float64)
        nq = xq.size
        zq = np.zeros((self.F, nq))
        leaf = np.zeros(nq, dtype=np.int32)
        st = (ctypes.c_longlong * 3)()
        self.H.hostsim_eval(self.h, _p(xq), _p(yq), nq, _p(zq), _p(leaf), mode, max_steps, st)
        return leaf, zq, dict(deferred=st[0], steps=st[1], outside=st[2])

    def deriv(self, leaf):
        leaf = np.ascontiguousarray(leaf, dtype=np.int32)
        dx = np.zeros((self.F, leaf.size))
        dy = np.zeros((self.F, leaf.size))
        self.H.hostsim_deriv(self.h, _p(leaf), leaf.size, _p(dx), _p(dy))
        return dx, dy
