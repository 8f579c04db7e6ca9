"""Randomised adversarial check of the kernel logic (tests/hostsim runs csrc/gsi_eval_core.h on the
host) against the oracle port: small triangulations of every flavour the reference's own tests
use (uniform, gridded = cocircular, collinear runs, anisotropic), queries ON vertices and edge
points, a few ulps and 1e-15..1e-6 away from them, outside the hull and far outside.  Both the
fast path (grid-seeded walk + margin, deferring to the exact replay) and the exact replay with
entry nodes must return the oracle's leaf ids and value bits."""
import numpy as np
import pytest

import gsl_scattered_interpolation_b200 as g
import refimpl
from gsl_scattered_interpolation_b200 import synth
from hostsim_py import HostSim


def _nodes(rng, kind, n):
    if kind == "uniform":
        return rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    if kind == "grid":                      # cocircular quadruples everywhere (gridded_test)
        m = max(2, int(np.sqrt(n)))
        gx, gy = np.meshgrid(np.arange(m, dtype=float), np.arange(m, dtype=float))
        return gx.ravel(), gy.ravel()
    if kind == "grid_jitter":               # almost cocircular
        m = max(2, int(np.sqrt(n)))
        gx, gy = np.meshgrid(np.arange(m, dtype=float), np.arange(m, dtype=float))
        return gx.ravel() + rng.uniform(-1e-9, 1e-9, m * m), gy.ravel() + rng.uniform(-1e-9, 1e-9, m * m)
    if kind == "collinear":                 # runs of collinear nodes inside a random cloud
        k = max(3, n // 4)
        t = np.sort(rng.uniform(0, 1, k))
        return (np.concatenate([rng.uniform(0, 1, n - k), t]),
                np.concatenate([rng.uniform(0, 1, n - k), 0.25 + 0.5 * t]))
    if kind == "aniso":                     # skinny triangles, very different axis scales
        return rng.uniform(0, 1000, n), rng.normal(0, 1e-3, n)
    raise ValueError(kind)


def _queries(rng, xa, ya, tree):
    n = xa.size
    leaves = tree["pts3"].reshape(-1, 3)[tree["type"] == 0]
    real = leaves[(leaves >= 0).all(axis=1)]
    rows = tree["shuffle"].astype(np.int64)
    qx, qy = [xa.copy()], [ya.copy()]                                   # on every vertex
    for d in (1, -1):                                                   # one ulp off every vertex
        qx.append(np.nextafter(xa, d * np.inf)); qy.append(ya.copy())
        qx.append(xa.copy()); qy.append(np.nextafter(ya, d * np.inf))
    if real.size:
        pick = real[rng.integers(0, real.shape[0], 600)]
        a, b, c = rows[pick[:, 0]], rows[pick[:, 1]], rows[pick[:, 2]]
        for t in (0.5, 0.25, rng.uniform(0, 1, a.size)):
            ex, ey = xa[a] + t * (xa[b] - xa[a]), ya[a] + t * (ya[b] - ya[a])
            qx.append(ex); qy.append(ey)                                # on edges (up to rounding)
            for eps in (1e-15, 1e-12, 1e-9, 1e-6):                      # just off the edge, both sides
                nx_, ny_ = -(ya[b] - ya[a]), xa[b] - xa[a]
                qx.append(ex + eps * nx_); qy.append(ey + eps * ny_)
                qx.append(ex - eps * nx_); qy.append(ey - eps * ny_)
        w = rng.dirichlet(np.ones(3), a.size)                           # interior points
        qx.append(w[:, 0] * xa[a] + w[:, 1] * xa[b] + w[:, 2] * xa[c])
        qy.append(w[:, 0] * ya[a] + w[:, 1] * ya[b] + w[:, 2] * ya[c])
    x0, x1, y0, y1 = xa.min(), xa.max(), ya.min(), ya.max()
    wx, wy = max(x1 - x0, 1e-9), max(y1 - y0, 1e-9)
    qx.append(rng.uniform(x0 - 0.3 * wx, x1 + 0.3 * wx, 800)); qy.append(rng.uniform(y0 - 0.3 * wy, y1 + 0.3 * wy, 800))
    qx.append(rng.uniform(x0 - 40 * wx, x1 + 40 * wx, 200)); qy.append(rng.uniform(y0 - 40 * wy, y1 + 40 * wy, 200))
    return np.concatenate(qx), np.concatenate(qy)


@pytest.mark.parametrize("kind", ["uniform", "grid", "grid_jitter", "collinear", "aniso"])
@pytest.mark.parametrize("flags", [g.DEFAULT, g.NOSTANDARDIZE, g.ISOSCALE])
def test_fuzz_kernel_logic_vs_port(kind, flags):
    rng = np.random.default_rng(1000 * sum(map(ord, kind)) + 17 * int(flags) + 7)
    checked = 0
    for trial in range(10):
        n = int(rng.integers(3, 260))
        xa, ya = _nodes(rng, kind, n)
        n = xa.size
        order = g.shuffle_mt19937(int(rng.integers(1, 10 ** 6)), n) if trial % 2 == 0 else None
        za = synth.values(n, 2)
        w = g.ScatteredCuda(n, 2)
        try:
            w.build(xa, ya, flags, order)
        except g.GslError:
            w.free()                       # the reference's builder asserts on this input too
            continue
        tree = w.get_tree(w.n_simplexes(), n)
        w.free()
        xq, yq = _queries(rng, xa, ya, tree)
        port = refimpl.PortTree(tree, xa, ya)
        pl, pv, _ = port.eval(xq, yq, za)
        port.close()
        for cpl in (4.0, 0.5, 40.0):
            hs = HostSim(xa, ya, za, flags, order, cells_per_leaf=cpl)
            for mode in (0, 1):
                leaf, val, _ = hs.eval(xq, yq, mode)
                assert np.array_equal(leaf, pl), (kind, flags, trial, cpl, mode)
                assert refimpl.same_bits(val, pv), (kind, flags, trial, cpl, mode)
            hs.close()
        checked += 1
    assert checked >= 1
