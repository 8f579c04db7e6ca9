"""The per-query logic the CUDA kernels are compiled from (csrc/gsi_eval_core.h), run on the
host by tests/hostsim: exact DAG replay and the grid-seeded walk with its safety margin must
both return the reference's leaf ids and value bits (golden vectors + oracle port)."""
import numpy as np
import pytest

import gsl_scattered_interpolation_b200 as g
import refimpl
from gsl_scattered_interpolation_b200 import synth
from hostsim_py import HostSim


@pytest.mark.parametrize("mode", [0, 1])
def test_logic_weather_golden(golden, mode):
    w = golden["weather"]
    hs = HostSim(w[:, 0], w[:, 1], w[:, 2], g.DEFAULT, g.shuffle_mt19937(0, 50))
    q = golden["weatherB_q"]
    leaf, val, st = hs.eval(q[:, 0], q[:, 1], mode)
    assert np.array_equal(leaf, golden["weatherB_leaf"])
    assert refimpl.same_bits(val, golden["weatherB_val"])
    assert st["outside"] == 1
    x0, dx, nx, y0, dy, ny = golden["weatherB_raster_spec"]
    nx, ny = int(nx), int(ny)
    gx, gy = x0 + dx * np.arange(nx), y0 + dy * np.arange(ny)
    leaf, val, st = hs.eval(np.repeat(gx, ny), np.tile(gy, nx), mode)
    assert np.array_equal(leaf.reshape(nx, ny), golden["weatherB_raster_leaf"])
    assert refimpl.same_bits(val.reshape(1, nx, ny), golden["weatherB_raster_val"])
    hs.close()


@pytest.mark.parametrize("mode,cpl", [(0, 4.0), (0, 0.25), (0, 64.0), (1, 4.0)])
def test_logic_config1_golden(golden, mode, cpl):
    xa, ya = synth.uniform_nodes(1000)
    hs = HostSim(xa, ya, synth.values(1000, 1), g.DEFAULT,
                 g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000), cells_per_leaf=cpl)
    xq, yq = synth.uniform_queries(10000)
    leaf, val, st = hs.eval(xq, yq, mode)
    assert np.array_equal(leaf, golden["cfg1_leaf"])
    assert refimpl.same_bits(val, golden["cfg1_val"])
    if mode == 0:
        assert st["deferred"] < 0.01 * xq.size      # the fast path really is the common path
    hs.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_logic_config3_mixed_golden(golden, mode):
    """clustered/anisotropic nodes; queries on vertices, on edge midpoints, outside the hull
    and far outside the bbox (cage leaves), 3 fields"""
    n = 3000
    xa, ya = synth.clustered_nodes(n)
    hs = HostSim(xa, ya, synth.values(n, 3), g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
    xq, yq = synth.mixed_queries(20000, xa, ya, golden["cfg3_leaf_rows"])
    leaf, val, st = hs.eval(xq, yq, mode)
    assert np.array_equal(leaf, golden["cfg3_leaf"])
    assert refimpl.same_bits(val, golden["cfg3_val"])
    if mode == 0:
        # the 10 % of queries that coincide with data nodes are answered from the vertex memo
        # (built by the exact replay at init), the on-edge ones by the replay itself
        assert 0.08 * len(xq) < hs.memo_hits() <= st["deferred"]
        # ... and every data node, queried directly, hits the memo with the replay's own answer
        before = hs.memo_hits()
        l0, v0, s0 = hs.eval(xa, ya, 0)
        l1, v1, _ = hs.eval(xa, ya, 1)
        assert hs.memo_hits() - before == s0["deferred"] >= 0.95 * n
        assert np.array_equal(l0, l1) and refimpl.same_bits(v0, v1)
    hs.close()


def test_logic_walk_cap_and_nonfinite(golden):
    """a walk cap of 1 step and NaN/inf queries only move work to the exact replay"""
    xa, ya = synth.uniform_nodes(1000)
    za = synth.values(1000, 1)
    order = g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000)
    hs = HostSim(xa, ya, za, g.DEFAULT, order, cells_per_leaf=0.05)
    xq, yq = synth.uniform_queries(10000)
    leaf, val, st = hs.eval(xq, yq, 0, max_steps=1)
    assert np.array_equal(leaf, golden["cfg1_leaf"]) and refimpl.same_bits(val, golden["cfg1_val"])
    assert st["deferred"] > 1000
    bad = np.array([np.nan, np.inf, -np.inf, 0.5])
    l0, v0, _ = hs.eval(bad, np.array([0.5, 0.5, 0.5, np.nan]), 0)
    l1, v1, _ = hs.eval(bad, np.array([0.5, 0.5, 0.5, np.nan]), 1)
    assert np.array_equal(l0, l1) and refimpl.same_bits(v0, v1)
    port = refimpl.PortTree({k: golden["cfg1_" + k] for k in
                             ("pts3", "lnk3", "type", "shuffle", "shift", "scale", "seed_points")}, xa, ya)
    pl, pv, _ = port.eval(bad, np.array([0.5, 0.5, 0.5, np.nan]), za)
    assert np.array_equal(l1, pl) and refimpl.same_bits(v1, pv)
    hs.close()


@pytest.mark.parametrize("kind,n", [("uniform", 20000), ("clustered", 20000), ("uniform", 300)])
def test_compact_seed_table_has_no_escapes_and_coarse_blocks(kind, n):
    """slab v6+: at most 32 x 32 seed blocks (the base table fits 4 KB of shared memory), the grid of a
    large triangulation is a multiple of 8 blocks per side, and every cell has a 16-bit seed (strays take a
    neighbour's in-window seed): the walk never needs the int32 table.  The seed is only the start of the
    walk, so the results stay those of the exact replay."""
    xa, ya = synth.uniform_nodes(n) if kind == "uniform" else synth.clustered_nodes(n)
    za = synth.values(n, 1)
    hs = HostSim(xa, ya, za, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
    G, shift, escapes = hs.seed_table()
    blocks = (G + (1 << shift) - 1) >> shift
    assert shift >= 3 and blocks <= 32
    if G >= 256:
        assert G % (8 << shift) == 0
    assert escapes == 0
    xq, yq = synth.uniform_queries(20000)
    l0, v0, st = hs.eval(xq, yq, 0)
    l1, v1, _ = hs.eval(xq, yq, 1)
    assert np.array_equal(l0, l1) and refimpl.same_bits(v0, v1)
    assert st["steps"] / max(1, len(xq) - st["deferred"]) < (1.6 if kind == "uniform" else 4.0)
    hs.close()


@pytest.mark.parametrize("kind", ["collinear", "grid10"])
def test_logic_degenerate_inputs_vs_port(kind):
    """collinear / cocircular nodes (gridded_test of the example) give near-singular and
    non-Delaunay leaves, and queries exactly on vertices and edges: the exact replay must
    follow the reference's first-hit / best-worst / stale-coordinate rules
    (linear_simplex.c:363-400,641-642), and the fast path must defer what it cannot prove"""
    rng = np.random.default_rng(3)
    if kind == "collinear":
        xa = np.concatenate([rng.uniform(0, 1, 200), np.linspace(0, 1, 30)])
        ya = np.concatenate([rng.uniform(0, 1, 200), np.full(30, 0.5)])
        flags, order = g.DEFAULT, g.shuffle_mt19937(5, xa.size)
        xq = np.concatenate([xa, rng.uniform(-0.2, 1.2, 3000), np.linspace(0, 1, 500)])
        yq = np.concatenate([ya, rng.uniform(-0.2, 1.2, 3000), np.full(500, 0.5)])
    else:
        xa, ya = (np.arange(100) % 10).astype(float), (np.arange(100) // 10).astype(float)
        flags, order = g.NOSTANDARDIZE, None
        xq = np.concatenate([xa, xa + 0.5, rng.uniform(-2, 11, 3000), xa])
        yq = np.concatenate([ya, ya, rng.uniform(-2, 11, 3000), ya + 0.5])
    n = xa.size
    za = synth.values(n, 2)
    w = g.ScatteredCuda(n, 2)
    w.build(xa, ya, flags, order)
    tree = w.get_tree(w.n_simplexes(), n)
    w.free()
    hs = HostSim(xa, ya, za, flags, order)
    port = refimpl.PortTree(tree, xa, ya)
    pl, pv, _ = port.eval(xq, yq, za)
    for mode in (0, 1):
        leaf, val, _ = hs.eval(xq, yq, mode)
        assert np.array_equal(leaf, pl)
        assert refimpl.same_bits(val, pv)
    hs.close()
    port.close()


def test_linear_precision_property():
    """barycentric interpolation reproduces a linear field in every non-cage leaf
    (oracle-independent property)"""
    n = 5000
    xa, ya = synth.clustered_nodes(n, seed=77)
    za = (xa + 1000.0 * ya)[None, :]
    hs = HostSim(xa, ya, za, g.DEFAULT, g.shuffle_mt19937(1, n))
    rng = np.random.default_rng(0)
    k = rng.integers(0, n, (4000, 3))
    wts = rng.dirichlet(np.ones(3), 4000)
    xq = (xa[k] * wts).sum(axis=1)
    yq = (ya[k] * wts).sum(axis=1)                      # inside the hull by construction
    leaf, val, _ = hs.eval(xq, yq, 0)
    want = xq + 1000.0 * yq
    assert np.max(np.abs(val[0] - want) / np.abs(want)) < 1e-12
    hs.close()


@pytest.mark.parametrize("kind,cpl", [("uniform", 4.0), ("uniform", 32.0), ("clustered", 4.0), ("grid10", 16.0)])
def test_logic_entry_nodes_shortcut_is_transparent(golden, kind, cpl):
    """The exact replay started at the per-cell entry node (gsi_entry_of_cell) must return what
    the replay from the root returns -- for random queries, queries on vertices and edge
    midpoints, and queries on the grid-cell borders where the cell assignment flips -- and the
    shortcut must be real (most cells start well below the root)."""
    rng = np.random.default_rng(11)
    if kind == "uniform":
        n = 1000
        xa, ya = synth.uniform_nodes(n)
        flags, order = g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n)
    elif kind == "clustered":
        n = 3000
        xa, ya = synth.clustered_nodes(n)
        flags, order = g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n)
    else:
        n = 100
        xa, ya = (np.arange(100) % 10).astype(float), (np.arange(100) // 10).astype(float)
        flags, order = g.NOSTANDARDIZE, None
    za = synth.values(n, 2)
    a = HostSim(xa, ya, za, flags, order, cells_per_leaf=cpl, exact_entry=1)
    b = HostSim(xa, ya, za, flags, order, cells_per_leaf=cpl, exact_entry=0)
    G = a.H.hostsim_grid_g(a.h)
    x0, x1, y0, y1 = xa.min(), xa.max(), ya.min(), ya.max()
    k = rng.integers(0, n, (4000, 2))
    border = np.arange(0, G + 1) / G
    bx = x0 + (x1 - x0) * np.repeat(border, 8)
    by = y0 + (y1 - y0) * rng.uniform(0, 1, bx.size)
    xq = np.concatenate([rng.uniform(x0 - 0.2 * (x1 - x0), x1 + 0.2 * (x1 - x0), 20000), xa,
                         0.5 * (xa[k[:, 0]] + xa[k[:, 1]]), bx, by * 0 + rng.uniform(x0, x1, bx.size),
                         np.nextafter(bx, np.inf), np.nextafter(bx, -np.inf)])
    yq = np.concatenate([rng.uniform(y0 - 0.2 * (y1 - y0), y1 + 0.2 * (y1 - y0), 20000), ya,
                         0.5 * (ya[k[:, 0]] + ya[k[:, 1]]), by, y0 + (y1 - y0) * np.repeat(border, 8),
                         by, by])
    for mode in (1, 0):
        la, va, _ = a.eval(xq, yq, mode)
        lb, vb, _ = b.eval(xq, yq, mode)
        assert np.array_equal(la, lb)
        assert refimpl.same_bits(va, vb)
    ent = a.entry()
    interior = ent.reshape(G, G)[1:-1, 1:-1].ravel()
    assert (ent.reshape(G, G)[0] == 0).all() and (ent.reshape(G, G)[:, -1] == 0).all()   # border cells: root
    assert (b.entry() == 0).all()
    if kind != "grid10":
        assert (interior != 0).mean() > 0.9
        if cpl > 8:                                  # fine cells often sit inside one leaf
            assert any(a.node_is_leaf(e) for e in interior[:: max(1, interior.size // 2000)])
    a.close()
    b.close()


@pytest.mark.parametrize("n,tile_leaves", [(1000, 1000), (1000, 7), (20000, 1000), (20000, 100)])
def test_tile_table_partitions_the_leaf_order(n, tile_leaves):
    """slab section `tiles`: Morton-prefix tiles are consecutive slot ranges covering every leaf"""
    xa, ya = synth.uniform_nodes(n)
    hs = HostSim(xa, ya, synth.values(n, 1), g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n),
                 tile_leaves=tile_leaves)
    bits, tiles, biggest = hs.tiles()
    L = hs.H.hostsim_n_leaves(hs.h)
    assert 2 <= bits <= 12 and tiles.size == (1 << bits) + 1
    assert tiles[0] == 0 and tiles[-1] == L and np.all(np.diff(tiles) >= 0)
    assert biggest == np.diff(tiles).max()
    if bits < 12:
        assert L / (1 << bits) <= 1.05 * tile_leaves
    hs.close()
