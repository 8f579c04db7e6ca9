"""The oracle is pinned: the plain-C restatement (oracle/simplex_oracle.c) must reproduce
every golden vector frozen from the real reference (tests/golden/make_golden.py), and, when
oracle/_ref is present, the reference itself on fresh random inputs."""
import numpy as np
import pytest

import refimpl
from gsl_scattered_interpolation_b200 import synth


def _tree(golden, prefix):
    return {k: golden[f"{prefix}_{k}"] for k in
            ("pts3", "lnk3", "type", "shuffle", "shift", "scale", "seed_points")}


def test_port_weather_golden_table(golden):
    w = golden["weather"]
    port = refimpl.PortTree(_tree(golden, "weatherB"), w[:, 0], w[:, 1])
    q = golden["weatherB_q"]
    leaf, val, nout = port.eval(q[:, 0], q[:, 1], w[:, 2])
    assert np.array_equal(leaf, golden["weatherB_leaf"])
    assert refimpl.same_bits(val, golden["weatherB_val"])
    # SURVEY.md section 4 table (values printed with %.17g by the reference build)
    assert list(leaf[:8]) == [275, 207, 144, 52, 363, 204, 62, 161]
    assert val[0, 0] == 274.08608080148747 and val[0, 1] == 288.70999999999998
    assert val[0, 2] == 281.8310918271132 and val[0, 6] == 271.83200833901174
    assert leaf[8] == -1 and np.isnan(val[0, 8]) and nout == 1   # outside the cage
    port.close()


def test_port_weather_tree_a(golden):
    w = golden["weather"]
    port = refimpl.PortTree(_tree(golden, "weatherA"), w[:, 0], w[:, 1])
    q = golden["weatherB_q"][:8]
    leaf, val, _ = port.eval(q[:, 0], q[:, 1], w[:, 2])
    assert np.array_equal(leaf, golden["weatherA_leaf"])
    assert refimpl.same_bits(val, golden["weatherA_val"])
    assert leaf[0] == 330 and val[0, 0] == 274.05964120079028
    port.close()


def test_port_weather_raster(golden):
    w = golden["weather"]
    port = refimpl.PortTree(_tree(golden, "weatherB"), w[:, 0], w[:, 1])
    x0, dx, nx, y0, dy, ny = golden["weatherB_raster_spec"]
    nx, ny = int(nx), int(ny)
    gx = x0 + dx * np.arange(nx)
    gy = y0 + dy * np.arange(ny)
    leaf, val, _ = port.eval(np.repeat(gx, ny), np.tile(gy, nx), w[:, 2])
    assert np.array_equal(leaf.reshape(nx, ny), golden["weatherB_raster_leaf"])
    assert refimpl.same_bits(val.reshape(1, nx, ny), golden["weatherB_raster_val"])
    port.close()


def test_port_config1(golden):
    xa, ya = synth.uniform_nodes(1000)
    za = synth.values(1000, 1)
    port = refimpl.PortTree(_tree(golden, "cfg1"), xa, ya)
    xq, yq = synth.uniform_queries(10000)
    leaf, val, nout = port.eval(xq, yq, za)
    assert nout == 0
    assert np.array_equal(leaf, golden["cfg1_leaf"])
    assert refimpl.same_bits(val, golden["cfg1_val"])
    port.close()


def test_trivial_known_answers(golden):
    """scattered_interp_example.c:51-77"""
    assert golden["trivial_empty_leaf"][0] == 0 and golden["trivial_empty_val"][0, 0] == 0.0
    pts = golden["trivial_one_pts3"].reshape(-1, 3)
    lnk = golden["trivial_one_lnk3"].reshape(-1, 3)
    assert golden["trivial_one_type"][0] != 0
    assert [list(pts[lnk[0, i]]) for i in range(3)] == [[0, -2, -3], [0, -1, -3], [0, -1, -2]]
    leaf = int(golden["trivial_second_leaf"][0])
    assert list(pts[leaf]) == [0, -2, -3]


@pytest.mark.skipif(not refimpl.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("n,flags,seed,kind", [(300, 0, 7, "uniform"), (2000, 0, 99, "cluster"),
                                               (500, 1, None, "uniform"), (800, 2, 3, "cluster")])
def test_port_matches_reference(n, flags, seed, kind):
    xa, ya = synth.uniform_nodes(n, seed=n) if kind == "uniform" else synth.clustered_nodes(n, seed=n)
    za = synth.values(n, 2)
    rt = refimpl.RefTree(xa, ya, flags, seed)
    e = rt.export()
    rows = e["shuffle"][e["pts3"].reshape(-1, 3)[(e["type"] == 0)]
                        [(e["pts3"].reshape(-1, 3)[e["type"] == 0] >= 0).all(axis=1)]].astype(np.int64)
    xq, yq = synth.mixed_queries(6000, xa, ya, rows, seed=n + 1)
    rl, rv, rout = rt.eval(xq, yq, za)
    port = refimpl.PortTree(e, xa, ya)
    pl, pv, pout = port.eval(xq, yq, za)
    assert rout == pout
    assert np.array_equal(rl, pl)
    assert refimpl.same_bits(rv, pv)
    port.close()
    rt.close()


@pytest.mark.skipif(not refimpl.have_ref(), reason="oracle/_ref not built")
def test_stub_integrity_tu_is_transparent():
    """The stubbed debug TU must not change the tree: compare with the real integrity TU
    (check_delaunay after every insert) on n = 120, and run the reference's full check."""
    xa, ya = synth.uniform_nodes(120, seed=5)
    a = refimpl.RefTree(xa, ya, 0, 11)
    b = refimpl.RefTree(xa, ya, 0, 11, checked=True)
    ea, eb = a.export(), b.export()
    for k in ea:
        assert np.array_equal(ea[k], eb[k]), k
    assert b.check() == 1
    a.close()
    b.close()


def test_thread_safe_multi_thread_evaluation_matches_single_thread(golden):
    """the port evaluates one shared tree from several POSIX threads (the
    reference cannot: tree->tmp_vec1, linear_simplex.c:644) and returns the golden vectors whatever
    the thread count, including thread counts that do not divide the batch and exceed it."""
    xa, ya = synth.uniform_nodes(1000)
    za = synth.values(1000, 1)
    tree = {k: golden["cfg1_" + k] for k in ("pts3", "lnk3", "type", "shuffle", "shift", "scale", "seed_points")}
    port = refimpl.PortTree(tree, xa, ya)
    xq, yq = synth.uniform_queries(10000)
    for nt in (1, 2, 7, 16):
        leaf, val, out = port.eval_mt(xq, yq, za, nt)
        assert out == 0 and np.array_equal(leaf, golden["cfg1_leaf"]) and refimpl.same_bits(val, golden["cfg1_val"])
    leaf, val, out = port.eval_mt(np.array([0.5, 1e6, 0.25]), np.array([0.5, 0.0, 0.75]), za, 8)
    assert out == 1 and leaf[1] == -1 and np.isnan(val[0, 1]) and leaf[0] >= 0 and leaf[2] >= 0
    port.close()
