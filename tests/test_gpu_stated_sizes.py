"""BASELINE.json configs[1..4] at their STATED sizes, under the driver's `pytest -m gpu`:

  configs[1]  10^5 uniform nodes, 10^7 uniform queries, 1 field
  configs[2]  10^5 clustered/anisotropic nodes, 10^7 mixed queries (in hull, outside hull, far
              outside, exactly on nodes, on edge midpoints), 1 field
  configs[3]  10^6 nodes, 8192 x 8192 raster, 16 fields  (model: scattered_interp_example.c:176-199)
  configs[4]  10^6 nodes, uniform queries through the reorder pipeline (the bench runs the 10^9)

Each test compares the fast path with the exact DAG replay on 100 % of the queries (leaf ids and
value bits, on the device) AND with the real reference (oracle/_ref: the reference's own
simplex_tree_init / find_leaf / interp_point compiled in place, forked over the host cores) on a
strided subsample of >= 10^6 queries that contains every query class.  Without oracle/_ref (it
travels to the GPU box as a built .so) the port on the product's tree stands in."""
import os

import numpy as np
import pytest

import gsl_scattered_interpolation_b200 as g
import refimpl
from gsl_scattered_interpolation_b200 import synth

pytestmark = pytest.mark.gpu

NPROC = max(1, min(64, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
TREE_KEYS = ("pts3", "lnk3", "type", "shuffle", "shift", "scale", "seed_points")


def _torch():
    import torch
    return torch


class Oracle:
    """the real reference when oracle/_ref is present, else the port on the product's tree"""

    def __init__(self, xa, ya, w=None):
        self.ref = refimpl.RefTree(xa, ya, g.DEFAULT, synth.SEED_SHUFFLE) if refimpl.have_ref() else None
        self.port = None if self.ref else refimpl.PortTree(w.get_tree(w.n_simplexes(), len(xa)), xa, ya)

    def eval(self, xq, yq, za):
        if self.ref:
            leaf, val, _ = self.ref.eval(xq, yq, za, nproc=NPROC)
        else:
            leaf, val, _ = self.port.eval(xq, yq, za)
        return leaf, val

    def close(self):
        (self.ref or self.port).close()


def _device_fast_vs_exact(w, dx, dy, nq, F):
    """fast path and exact replay of all nq device-resident queries; returns the fast results"""
    torch = _torch()
    dev = dx.device
    za = torch.empty((F, nq), dtype=torch.float64, device=dev)
    la = torch.empty(nq, dtype=torch.int32, device=dev)
    zb = torch.empty((F, nq), dtype=torch.float64, device=dev)
    lb = torch.empty(nq, dtype=torch.int32, device=dev)
    w.set_param("mode", 0)
    w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), nq, za.data_ptr(), nq, la.data_ptr())
    w.set_param("mode", 1)
    w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), nq, zb.data_ptr(), nq, lb.data_ptr())
    w.set_param("mode", 0)
    bad = w.compare_device(0, la.data_ptr(), lb.data_ptr(), za.data_ptr(), zb.data_ptr(), nq, F * nq)
    assert bad == (0, 0)
    del zb, lb
    return za, la


def test_config2_1e5_uniform_nodes_1e7_queries():
    torch = _torch()
    n, nq = 100_000, 10_000_000
    xa, ya = synth.uniform_nodes(n)
    za = synth.values(n, 1)
    w = g.ScatteredCuda(n, 1)
    w.init(xa, ya, za, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
    dev = torch.device("cuda:0")
    dx = torch.empty(nq, dtype=torch.float64, device=dev)
    dy = torch.empty(nq, dtype=torch.float64, device=dev)
    w.generate_uniform_device(0, synth.SEED_QUERIES, 0, nq, 0.0, 1.0, 0.0, 1.0, dx.data_ptr(), dy.data_ptr())
    hx, hy = synth.uniform_queries(1_000_000)                       # device generator == host generator
    assert np.array_equal(dx[:1_000_000].cpu().numpy(), hx) and np.array_equal(dy[:1_000_000].cpu().numpy(), hy)
    zf, lf = _device_fast_vs_exact(w, dx, dy, nq, 1)
    assert w.info()["n_deferred"] < 1e-3 * nq
    st, hz, hl = w.eval_e(hx, hy)                                   # host-buffer API == device API
    assert st == g.SUCCESS and np.array_equal(hl, lf[:1_000_000].cpu().numpy())
    assert refimpl.same_bits(hz, zf[:, :1_000_000].cpu().numpy())
    sub = torch.arange(0, nq, 10, device=dev)                       # 10^6 queries
    orc = Oracle(xa, ya, w)
    rl, rv = orc.eval(dx[sub].cpu().numpy(), dy[sub].cpu().numpy(), za)
    assert np.array_equal(lf[sub].cpu().numpy(), rl)
    assert refimpl.same_bits(zf[:, sub].cpu().numpy(), rv)
    orc.close()
    w.free()


def test_config3_1e5_clustered_nodes_1e7_mixed_queries():
    torch = _torch()
    n, nq = 100_000, 10_000_000
    xa, ya = synth.clustered_nodes(n)
    za = synth.values(n, 1)
    w = g.ScatteredCuda(n, 1)
    w.init(xa, ya, za, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
    tree = w.get_tree(w.n_simplexes(), n)
    lv = tree["pts3"].reshape(-1, 3)[tree["type"] == 0]
    lv = lv[(lv >= 0).all(axis=1)]
    rows = tree["shuffle"][lv].astype(np.int64)
    xq, yq, cls = synth.mixed_queries(nq, xa, ya, rows, with_class=True)
    dev = torch.device("cuda:0")
    dx, dy = torch.from_numpy(xq).to(dev), torch.from_numpy(yq).to(dev)
    w.reset_counters()
    zf, lf = _device_fast_vs_exact(w, dx, dy, nq, 1)
    leaf = lf.cpu().numpy()
    assert (leaf >= 0).all()                                        # everything is inside the cage
    # the mix really exercises the tie-breaks: on-node and on-edge queries are deferred to the exact replay
    assert w.info()["n_deferred"] > 0.15 * nq
    sub = np.arange(0, nq, 10)                                      # 10^6 queries, every class present
    counts = np.bincount(cls[sub], minlength=5)
    assert (counts > 40_000).all(), counts
    orc = Oracle(xa, ya, w)
    if orc.ref:                                                     # stated-size builder parity on the way
        rt = orc.ref.export()
        for k in TREE_KEYS:
            assert np.array_equal(np.asarray(rt[k]), np.asarray(tree[k])), k
    rl, rv = orc.eval(xq[sub], yq[sub], za)
    assert np.array_equal(leaf[sub], rl)
    assert refimpl.same_bits(zf.cpu().numpy()[:, sub], rv)
    # cage leaves (a negative vertex id) catch the outside-hull classes
    cage = (tree["pts3"].reshape(-1, 3)[leaf[sub]] < 0).any(axis=1)
    assert cage[cls[sub] == 2].mean() > 0.99 and cage[cls[sub] == 1].mean() > 0.2   # (+-50x box still covers the hull)
    orc.close()
    w.free()


@pytest.fixture(scope="module")
def million():
    """10^6 uniform nodes with 16 fields (field 1 = the linear field x + 1000 y): the CUDA workspace
    and the oracle, shared by the tests below (the reference builder takes ~15-40 s at this size)"""
    n, F = 1_000_000, 16
    xa, ya = synth.uniform_nodes(n)
    za = synth.values(n, F)
    za[1] = xa + 1000.0 * ya
    w = g.ScatteredCuda(n, F)
    w.init(xa, ya, za, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
    orc = Oracle(xa, ya, w)
    yield dict(n=n, F=F, xa=xa, ya=ya, za=za, w=w, orc=orc)
    orc.close()
    w.free()


def test_million_nodes_tree_equals_the_reference_builders(million):
    """the product's host builder reproduces the reference's DAG node for node at 10^6 nodes"""
    if not million["orc"].ref:
        pytest.skip("oracle/_ref not built")
    w = million["w"]
    tree = w.get_tree(w.n_simplexes(), million["n"])
    rt = million["orc"].ref.export()
    assert million["orc"].ref.n_simplexes == tree["type"].size
    for k in TREE_KEYS:
        assert np.array_equal(np.asarray(rt[k]), np.asarray(tree[k])), k
    assert w.info()["n_leaves"] == 2 * million["n"] + 1             # T = 2n + 1


def test_config4_1e6_nodes_8192_raster_16_fields(million):
    torch = _torch()
    w, F = million["w"], million["F"]
    xa, ya, za = million["xa"], million["ya"], million["za"]
    nx = ny = 8192
    x0, y0 = float(xa.min()), float(ya.min())
    dx, dy = (float(xa.max()) - x0) / nx, (float(ya.max()) - y0) / ny
    dev = torch.device("cuda:0")
    npx = nx * ny
    z = torch.empty((F, npx), dtype=torch.float64, device=dev)
    l = torch.empty(npx, dtype=torch.int32, device=dev)
    zx = torch.empty((F, npx), dtype=torch.float64, device=dev)
    lx = torch.empty(npx, dtype=torch.int32, device=dev)
    w.set_param("mode", 0)
    assert w.eval_raster_device_e(0, x0, dx, nx, y0, dy, ny, z.data_ptr(), l.data_ptr()) == 0
    w.set_param("mode", 1)
    assert w.eval_raster_device_e(0, x0, dx, nx, y0, dy, ny, zx.data_ptr(), lx.data_ptr()) == 0
    w.set_param("mode", 0)
    assert w.compare_device(0, l.data_ptr(), lx.data_ptr(), z.data_ptr(), zx.data_ptr(), npx, F * npx) == (0, 0)
    del zx, lx
    # >= 10^6 pixels against the reference's own loop: x = min + step*i, two roundings (numpy: mul, then add)
    sub = np.arange(0, npx, 61, dtype=np.int64)
    assert sub.size >= 1_000_000
    i, j = sub // ny, sub % ny
    xs, ys = x0 + dx * i.astype(np.float64), y0 + dy * j.astype(np.float64)
    rl, rv = million["orc"].eval(xs, ys, za)
    tsub = torch.from_numpy(sub).to(dev)
    assert np.array_equal(l[tsub].cpu().numpy(), rl)
    assert refimpl.same_bits(z[:, tsub].cpu().numpy(), rv)
    # the bbox raster touches the hull: some pixels land in cage leaves and keep only real vertices' terms
    tree_pts = w.get_tree(w.n_simplexes(), million["n"])["pts3"].reshape(-1, 3)
    assert ((tree_pts[rl] < 0).any(axis=1)).sum() > 0


def test_million_nodes_uniform_queries_direct_and_reordered(million):
    """5*10^7 uniform queries, 16-field workspace (direct walk) and a 1-field workspace on the same
    tree (reorder pipeline, the bench's path): fast == exact on 100 %, reference on 10^6, linear
    precision of field 1."""
    torch = _torch()
    w, F, n = million["w"], million["F"], million["n"]
    xa, ya, za = million["xa"], million["ya"], million["za"]
    nq = 50_000_000
    dev = torch.device("cuda:0")
    dx = torch.empty(nq, dtype=torch.float64, device=dev)
    dy = torch.empty(nq, dtype=torch.float64, device=dev)
    w.generate_uniform_device(0, synth.SEED_QUERIES, 0, nq, 0.0, 1.0, 0.0, 1.0, dx.data_ptr(), dy.data_ptr())
    nq16 = 10_000_000
    w.reset_counters()
    zf, lf = _device_fast_vs_exact(w, dx, dy, nq16, F)
    assert w.info()["n_deferred"] < 1e-3 * nq16 and w.info()["n_reordered"] == 0
    tree = w.get_tree(w.n_simplexes(), n)
    lp = lf.cpu().numpy()
    real = (tree["pts3"].reshape(-1, 3)[lp] >= 0).all(axis=1)
    assert real.mean() > 0.999
    want = (dx[:nq16] + 1000.0 * dy[:nq16]).cpu().numpy()
    got = zf[1].cpu().numpy()
    assert (np.abs(got - want) / np.abs(want))[real].max() < 1e-12
    del zf
    # one field, same tree, through the reorder pipeline (auto: 5*10^7 queries, 2*10^6 leaves)
    w1 = g.ScatteredCuda(n, 1)
    w1.set_param("reorder", 2)
    w1.init_from_tree(tree["pts3"], tree["lnk3"], tree["type"], tree["shuffle"], tree["shift"], tree["scale"],
                      tree["seed_points"], np.stack([xa, ya], axis=1), za[:1])
    z1, l1 = _device_fast_vs_exact(w1, dx, dy, nq, 1)
    assert w1.info()["n_reordered"] == nq
    assert np.array_equal(l1[:nq16].cpu().numpy(), lp)               # same leaves as the direct walk
    sub = torch.arange(0, nq, 50, device=dev)                        # 10^6 queries
    rl, rv = million["orc"].eval(dx[sub].cpu().numpy(), dy[sub].cpu().numpy(), za[:1])
    assert np.array_equal(l1[sub].cpu().numpy(), rl)
    assert refimpl.same_bits(z1[:, sub].cpu().numpy(), rv)
    w1.free()
