"""Parity tests proper: the CUDA path, called through the C ABI, against the golden vectors
frozen from the reference, against the oracle (port and, where present, the real reference
in oracle/_ref), and -- at full size -- through size-independent properties (fast path ==
exact DAG replay on 100 % of the queries, linear precision, host API == device API).
Bar: leaf ids bit-exact, values bit-identical (stricter than the 1e-12 relative of
BASELINE.json's north_star)."""
import numpy as np
import pytest

import gsl_scattered_interpolation_b200 as g
import refimpl
from gsl_scattered_interpolation_b200 import synth

pytestmark = pytest.mark.gpu

REL_TOL = 1e-12   # north_star tolerance; the assertions below demand bit-identity instead


def _torch():
    import torch
    return torch


def make(xa, ya, za, flags=g.DEFAULT, order=None, **params):
    za = np.atleast_2d(za)
    w = g.ScatteredCuda(len(xa), za.shape[0])
    for k, v in params.items():
        w.set_param(k, v)
    w.init(xa, ya, za, flags, order)
    return w


@pytest.mark.parametrize("mode", [0, 1])
def test_weather_golden(golden, mode):
    wd = golden["weather"]
    w = make(wd[:, 0], wd[:, 1], wd[:, 2], g.DEFAULT, g.shuffle_mt19937(0, 50), mode=mode)
    q = golden["weatherB_q"]
    st, zq, leaf = w.eval_e(q[:, 0], q[:, 1])
    assert st == g.EDOM                                   # (1e6, 0) is outside the cage
    assert np.array_equal(leaf, golden["weatherB_leaf"])
    assert refimpl.same_bits(zq, golden["weatherB_val"])
    assert w.eval(-88.0, 41.0) == 274.08608080148747
    assert np.isnan(w.eval(1e6, 0.0))
    x0, dx, nx, y0, dy, ny = golden["weatherB_raster_spec"]
    st, rz, rl = w.eval_raster_e(x0, dx, int(nx), y0, dy, int(ny))
    assert st == g.SUCCESS
    assert np.array_equal(rl, golden["weatherB_raster_leaf"])
    assert refimpl.same_bits(rz, golden["weatherB_raster_val"])
    w.free()


def test_weather_tree_a_nostandardize(golden):
    wd = golden["weather"]
    w = make(wd[:, 0], wd[:, 1], wd[:, 2], g.NOSTANDARDIZE, None)
    q = golden["weatherB_q"][:8]
    st, zq, leaf = w.eval_e(q[:, 0], q[:, 1])
    assert np.array_equal(leaf, golden["weatherA_leaf"]) and refimpl.same_bits(zq, golden["weatherA_val"])
    w.free()


def test_trivial_empty_and_single_node(golden):
    """scattered_interp_example.c:44-77: empty cage interpolates to 0; one node, then locate"""
    w = make(np.zeros(0), np.zeros(0), np.zeros((1, 0)), g.NOSTANDARDIZE, None)
    st, zq, leaf = w.eval_e(np.array([-88.0]), np.array([41.0]))
    assert st == g.SUCCESS and leaf[0] == 0 and zq[0, 0] == 0.0
    w.free()
    w = make(np.array([-88.0]), np.array([41.0]), np.array([[5.0]]), g.NOSTANDARDIZE, None)
    st, zq, leaf = w.eval_e(np.array([-89.0]), np.array([41.0]))
    assert leaf[0] == golden["trivial_second_leaf"][0]
    assert refimpl.same_bits(zq, golden["trivial_second_val"])
    st, zq, leaf = w.eval_e(np.zeros(0), np.zeros(0))      # empty batch
    assert st == g.SUCCESS and leaf.size == 0
    w.free()


@pytest.mark.parametrize("mode,cpl", [(0, 4.0), (0, 0.25), (0, 32.0), (1, 4.0)])
def test_config1_golden(golden, mode, cpl):
    xa, ya = synth.uniform_nodes(1000)
    w = make(xa, ya, synth.values(1000, 1), g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000),
             mode=mode, cells_per_leaf=cpl)
    xq, yq = synth.uniform_queries(10000)
    st, zq, leaf = w.eval_e(xq, yq)
    assert st == g.SUCCESS
    assert np.array_equal(leaf, golden["cfg1_leaf"])
    assert refimpl.same_bits(zq, golden["cfg1_val"])
    x0, dx, nx, y0, dy, ny = golden["cfg1_raster_spec"]
    st, rz, rl = w.eval_raster_e(x0, dx, int(nx), y0, dy, int(ny))
    assert np.array_equal(rl, golden["cfg1_raster_leaf"]) and refimpl.same_bits(rz, golden["cfg1_raster_val"])
    info = w.info()
    if mode == 0:
        assert info["n_deferred"] < 0.02 * 20000           # the fast path is the common path
    w.free()


def test_config1_isoscale_golden(golden):
    xa, ya = synth.uniform_nodes(1000)
    w = make(xa * 3.0, ya, synth.values(1000, 1), g.ISOSCALE, g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000))
    xq, yq = synth.uniform_queries(10000)
    st, zq, leaf = w.eval_e(xq * 3.0, yq)
    assert np.array_equal(leaf, golden["cfg1_iso_leaf"]) and refimpl.same_bits(zq, golden["cfg1_iso_val"])
    w.free()


@pytest.mark.parametrize("mode", [0, 1])
def test_config3_mixed_golden(golden, mode):
    n = 3000
    xa, ya = synth.clustered_nodes(n)
    w = make(xa, ya, synth.values(n, 3), g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n), mode=mode)
    xq, yq = synth.mixed_queries(20000, xa, ya, golden["cfg3_leaf_rows"])
    st, zq, leaf = w.eval_e(xq, yq)
    assert st == g.SUCCESS
    assert np.array_equal(leaf, golden["cfg3_leaf"])
    assert refimpl.same_bits(zq, golden["cfg3_val"])
    w.free()


def test_tiny_deferred_list_overflow_and_walk_cap(golden):
    """a 1-step walk cap defers most queries; a tiny list forces the in-place resolution"""
    xa, ya = synth.uniform_nodes(1000)
    w = make(xa, ya, synth.values(1000, 1), g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000),
             cells_per_leaf=0.05, max_steps=1, list_fraction=1e-4)
    xq, yq = synth.uniform_queries(10000)
    xq, yq = np.tile(xq, 20), np.tile(yq, 20)              # 2e5 queries, list capacity 65536
    st, zq, leaf = w.eval_e(xq, yq)
    assert np.array_equal(leaf, np.tile(golden["cfg1_leaf"], 20))
    assert refimpl.same_bits(zq, np.tile(golden["cfg1_val"], (1, 20)))
    assert w.info()["n_deferred"] > 65536
    w.free()


def test_nonfinite_queries_follow_the_reference(golden):
    xa, ya = synth.uniform_nodes(1000)
    za = synth.values(1000, 1)
    w = make(xa, ya, za, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000))
    port = refimpl.PortTree({k: golden["cfg1_" + k] for k in
                             ("pts3", "lnk3", "type", "shuffle", "shift", "scale", "seed_points")}, xa, ya)
    bx = np.array([np.nan, np.inf, -np.inf, 0.5, 1e300])
    by = np.array([0.5, 0.5, 0.5, np.nan, 0.5])
    st, zq, leaf = w.eval_e(bx, by)
    pl, pv, _ = port.eval(bx, by, za)
    assert np.array_equal(leaf, pl) and refimpl.same_bits(zq, pv)
    w.free()


def test_degenerate_grid_nodes_vs_port():
    """gridded_test(2,2)-style cocircular nodes, queries exactly on vertices / edges / centres"""
    xa, ya = (np.arange(100) % 10).astype(float), (np.arange(100) // 10).astype(float)
    za = synth.values(100, 2)
    w = make(xa, ya, za, g.NOSTANDARDIZE, None)
    tree = w.get_tree(w.n_simplexes(), 100)
    rng = np.random.default_rng(3)
    xq = np.concatenate([xa, xa + 0.5, rng.uniform(-2, 11, 3000), xa, xa + 0.5])
    yq = np.concatenate([ya, ya, rng.uniform(-2, 11, 3000), ya + 0.5, ya + 0.5])
    st, zq, leaf = w.eval_e(xq, yq)
    port = refimpl.PortTree(tree, xa, ya)
    pl, pv, _ = port.eval(xq, yq, za)
    assert np.array_equal(leaf, pl) and refimpl.same_bits(zq, pv)
    w.free()


def test_sixteen_fields_raster_vs_port():
    """config 4 shape at test size: 2*10^4 nodes, 16 fields, 384 x 256 raster"""
    n, F = 20000, 16
    xa, ya = synth.uniform_nodes(n)
    za = synth.values(n, F)
    order = g.shuffle_mt19937(synth.SEED_SHUFFLE, n)
    w = make(xa, ya, za, g.DEFAULT, order)
    nx, ny = 384, 256
    x0, y0 = xa.min(), ya.min()
    dx, dy = (xa.max() - x0) / nx, (ya.max() - y0) / ny
    st, rz, rl = w.eval_raster_e(x0, dx, nx, y0, dy, ny)
    assert st == g.SUCCESS and rz.shape == (F, nx, ny)
    port = refimpl.PortTree(w.get_tree(w.n_simplexes(), n), xa, ya)
    gx, gy = x0 + dx * np.arange(nx), y0 + dy * np.arange(ny)
    pl, pv, _ = port.eval(np.repeat(gx, ny), np.tile(gy, nx), za)
    assert np.array_equal(rl.ravel(), pl)
    assert refimpl.same_bits(rz.reshape(F, -1), pv)
    # odd field count exercises the scalar gather path
    w3 = make(xa, ya, za[:3], g.DEFAULT, order)
    st, z3, l3 = w3.eval_e(np.repeat(gx, ny)[:5000], np.tile(gy, nx)[:5000])
    assert np.array_equal(l3, pl[:5000]) and refimpl.same_bits(z3, pv[:3, :5000])
    w3.free()
    w.free()


@pytest.mark.parametrize("kernel,extra", [(0, {}), (1, {}), (0, {"walk_variant": 0}), (0, {"walk_dynamic": 0}),
                                          (0, {"walk_pf": 0}), (0, {"walk_pf": 2}), (0, {"walk_pf": 3}),
                                          (0, {"walk_pf": 4}), (0, {"walk_chunk": 32}), (0, {"walk_ctas_per_sm": 1}),
                                          (0, {"walk_variant": 0, "leaf_stride": 96}), (0, {"walk_variant": 2, "leaf_stride": 96}),
                                          (0, {"walk_variant": 2, "walk_dynamic": 0, "walk_chunk": 64}),
                                          (0, {"walk_variant": 1, "walk_chunk": 64}),
                                          (0, {"walk_variant": 3}), (0, {"walk_variant": 3, "walk_chunk": 32}),
                                          (0, {"walk_variant": 3, "walk_dynamic": 0, "walk_chunk": 96}),
                                          (0, {"walk_variant": 3, "walk_ctas_per_sm": 1, "max_steps": 2}),
                                          (0, {"walk_variant": 3, "cells_per_leaf": 0.05}),
                                          (0, {"walk_variant": 2, "grid_g": 300}), (0, {"walk_variant": 3, "grid_g": 300})])
def test_every_fast_path_kernel_on_the_golden_sets(golden, kernel, extra):
    """the walk kernels (0: k_walkp by default, k_walk with walk_variant = 0, every way of dealing chunks
    and of requesting leaf lines ahead) and k_locate (1) on config 1 (1 field) and the config-3 mix (3 fields)"""
    xa, ya = synth.uniform_nodes(1000)
    w = make(xa, ya, synth.values(1000, 1), g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000),
             kernel=kernel, **extra)
    xq, yq = synth.uniform_queries(10000)
    st, zq, leaf = w.eval_e(xq, yq)
    assert np.array_equal(leaf, golden["cfg1_leaf"]) and refimpl.same_bits(zq, golden["cfg1_val"])
    # queries only, values only
    st, zq, _ = w.eval_e(xq, yq, want_leaf=False)
    assert refimpl.same_bits(zq, golden["cfg1_val"])
    st, _, leaf = w.eval_e(xq, yq, want_values=False)
    assert np.array_equal(leaf, golden["cfg1_leaf"])
    w.free()
    n = 3000
    xa, ya = synth.clustered_nodes(n)
    w = make(xa, ya, synth.values(n, 3), g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n), kernel=kernel, **extra)
    xq, yq = synth.mixed_queries(20000, xa, ya, golden["cfg3_leaf_rows"])
    xq[:3] = [np.nan, np.inf, 1e300]                    # non-finite / outside the cage
    st, zq, leaf = w.eval_e(xq, yq)
    assert st == g.EDOM and leaf[2] == -1
    assert np.array_equal(leaf[3:], golden["cfg3_leaf"][3:])
    assert refimpl.same_bits(zq[:, 3:], golden["cfg3_val"][:, 3:])
    w.free()


@pytest.mark.parametrize("F,coop", [(1, 2), (3, 2), (1, 1), (3, 1), (3, 0)])
def test_vertex_memo_answers_queries_at_data_nodes(golden, F, coop):
    """queries that coincide with data nodes are answered from the hash table the exact replay filled at
    init: same leaf ids and value bits as the replay itself (memo off, and mode 1 which never probes it),
    through the list kernels and, with a tiny list, through the overflow bitmask"""
    n = 3000
    xa, ya = synth.clustered_nodes(n)
    za = synth.values(n, F)
    order = g.shuffle_mt19937(synth.SEED_SHUFFLE, n)
    xm, ym = synth.mixed_queries(20000, xa, ya, golden["cfg3_leaf_rows"])
    xq, yq = np.concatenate([xa, xm, xa[::-1]]), np.concatenate([ya, ym, ya[::-1]])
    ref = make(xa, ya, za, g.DEFAULT, order, mode=1)
    st, zr, lr = ref.eval_e(xq, yq)
    for extra in ({}, {"vertex_memo": 0}, {"list_fraction": 1e-4}):
        w = make(xa, ya, za, g.DEFAULT, order, exact_coop=coop, **extra)
        st, zq, leaf = w.eval_e(xq, yq)
        assert np.array_equal(leaf, lr), extra
        assert refimpl.same_bits(zq, zr), extra
        assert w.info()["n_deferred"] >= 2 * 0.95 * n
        w.free()
    ref.free()


@pytest.mark.parametrize("kernel", [0, 1, 2])
def test_raster_odd_shapes_and_chunking(golden, kernel):
    """rasters whose sides are not multiples of the 8 x 32 launch tile, evaluated in several
    device launches and several host pipeline stages, with every kernel variant"""
    xa, ya = synth.uniform_nodes(1000)
    za = synth.values(1000, 4)
    w = make(xa, ya, za, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000), kernel=kernel,
             chunk=700, host_chunk=300)
    port = refimpl.PortTree({k: golden["cfg1_" + k] for k in
                             ("pts3", "lnk3", "type", "shuffle", "shift", "scale", "seed_points")}, xa, ya)
    for nx, ny in [(37, 53), (1, 100), (100, 1), (9, 33), (64, 64)]:
        x0, y0, dx, dy = -0.05, 0.02, 1.1 / nx, 0.97 / ny
        st, rz, rl = w.eval_raster_e(x0, dx, nx, y0, dy, ny)
        gx, gy = x0 + dx * np.arange(nx), y0 + dy * np.arange(ny)
        pl, pv, _ = port.eval(np.repeat(gx, ny), np.tile(gy, nx), za)
        assert np.array_equal(rl.ravel(), pl), (nx, ny)
        assert refimpl.same_bits(rz.reshape(4, -1), pv), (nx, ny)
    w.free()


def test_init_from_reference_tree_and_slab_replication(golden):
    """true drop-in: adopt the arrays of a tree built by the reference's simplex_tree_init;
    then replicate the device slab into a second workspace"""
    torch = _torch()
    xa, ya = synth.uniform_nodes(1000)
    za = synth.values(1000, 1)
    w = g.ScatteredCuda(1000, 1)
    w.init_from_tree(golden["cfg1_pts3"], golden["cfg1_lnk3"], golden["cfg1_type"], golden["cfg1_shuffle"],
                     golden["cfg1_shift"], golden["cfg1_scale"], golden["cfg1_seed_points"],
                     np.stack([xa, ya], axis=1), za)
    xq, yq = synth.uniform_queries(10000)
    st, zq, leaf = w.eval_e(xq, yq)
    assert np.array_equal(leaf, golden["cfg1_leaf"]) and refimpl.same_bits(zq, golden["cfg1_val"])
    nbytes = w.slab_bytes()
    buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda:0")
    w.slab_export_device(0, buf.data_ptr())
    w2 = g.ScatteredCuda(1000, 1)
    w2.init_from_slab_device(buf.data_ptr(), nbytes)
    st, zq2, leaf2 = w2.eval_e(xq, yq)
    assert np.array_equal(leaf2, leaf) and refimpl.same_bits(zq2, zq)
    w2.free()
    w.free()


def test_pinned_host_buffers_and_small_pipeline_chunks(golden):
    """eval_e with pinned buffers and a pipeline chunk much smaller than the batch"""
    import ctypes
    from gsl_scattered_interpolation_b200 import _lib
    L = _lib.load()
    xa, ya = synth.uniform_nodes(1000)
    w = make(xa, ya, synth.values(1000, 1), g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000),
             host_chunk=777)
    nq = 10000
    px = L.gsl_scattered_cuda_host_alloc(nq * 8); py = L.gsl_scattered_cuda_host_alloc(nq * 8)
    pz = L.gsl_scattered_cuda_host_alloc(nq * 8); pl = L.gsl_scattered_cuda_host_alloc(nq * 4)
    assert px and py and pz and pl
    xq, yq = synth.uniform_queries(nq)
    ctypes.memmove(px, xq.ctypes.data, nq * 8); ctypes.memmove(py, yq.ctypes.data, nq * 8)
    assert w.eval_ptr_e(px, py, nq, pz, pl) == g.SUCCESS
    zq = np.ctypeslib.as_array(ctypes.cast(pz, ctypes.POINTER(ctypes.c_double)), (nq,)).copy()
    leaf = np.ctypeslib.as_array(ctypes.cast(pl, ctypes.POINTER(ctypes.c_int)), (nq,)).copy()
    assert np.array_equal(leaf, golden["cfg1_leaf"]) and refimpl.same_bits(zq[None, :], golden["cfg1_val"])
    for p in (px, py, pz, pl):
        L.gsl_scattered_cuda_host_free(p)
    w.free()


def test_two_devices_if_present(golden):
    if g.device_count() < 2:
        pytest.skip("single GPU box")
    xa, ya = synth.uniform_nodes(1000)
    w = g.ScatteredCuda(1000, 1)
    w.set_devices([0, 1])
    w.init(xa, ya, synth.values(1000, 1), g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000))
    xq, yq = synth.uniform_queries(10000)
    st, zq, leaf = w.eval_e(xq, yq)
    assert w.info()["n_devices"] == 2
    assert np.array_equal(leaf, golden["cfg1_leaf"]) and refimpl.same_bits(zq, golden["cfg1_val"])
    # the reorder pipeline (per-device shared-memory attributes, per-replica scratch) and the
    # raster split by columns on both replicas
    w.set_param("reorder", 1)
    st, zq, leaf = w.eval_e(xq, yq)
    assert np.array_equal(leaf, golden["cfg1_leaf"]) and refimpl.same_bits(zq, golden["cfg1_val"])
    assert w.info()["n_reordered"] == 10000
    w.set_param("reorder", 0)
    x0, dx, nx, y0, dy, ny = golden["cfg1_raster_spec"]
    st, rz, rl = w.eval_raster_e(x0, dx, int(nx), y0, dy, int(ny))
    assert np.array_equal(rl, golden["cfg1_raster_leaf"]) and refimpl.same_bits(rz, golden["cfg1_raster_val"])
    w.free()


# ---- reorder pipeline (k_bin -> k_tile_locate -> k_unpermute) -------------------------------

@pytest.mark.parametrize("stride", [96, 128])       # 96: the tile arrives by one TMA bulk copy; 128: by 128-bit loads
@pytest.mark.parametrize("chunk,tile_leaves,threads", [(8192, 1000, 768), (8192, 3, 768), (4096, 40, 768), (96, 10, 768),
                                                       (1024, 100000, 768), (8192, 1000, 64), (96, 3, 32)])
def test_reorder_pipeline_on_the_golden_sets(golden, chunk, tile_leaves, threads, stride):
    """Forced reorder (param reorder=1) on config 1 and on the config-3 mix (3 fields, on-edge /
    on-vertex / outside-hull queries that are deferred to the exact replay): every chunk size, tile
    size (one tile for everything, tiles of a handful of leaves) and CTA shape of the tile kernel must give the reference's leaf ids and value bits."""
    xa, ya = synth.uniform_nodes(1000)
    w = make(xa, ya, synth.values(1000, 1), g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000),
             reorder=1, reorder_chunk=chunk, tile_leaves=tile_leaves, tile_threads=threads, leaf_stride=stride)
    xq, yq = synth.uniform_queries(10000)
    st, zq, leaf = w.eval_e(xq, yq)
    assert st == g.SUCCESS
    assert np.array_equal(leaf, golden["cfg1_leaf"])
    assert refimpl.same_bits(zq, golden["cfg1_val"])
    assert w.info()["n_reordered"] == 10000
    # leaf ids only / values only
    st, zq2, _ = w.eval_e(xq, yq, want_leaf=False)
    assert refimpl.same_bits(zq2, golden["cfg1_val"])
    st, _, leaf2 = w.eval_e(xq, yq, want_values=False)
    assert np.array_equal(leaf2, golden["cfg1_leaf"])
    w.free()
    n = 3000
    xa, ya = synth.clustered_nodes(n)
    w = make(xa, ya, synth.values(n, 3), g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n),
             reorder=1, reorder_chunk=chunk, tile_leaves=tile_leaves, tile_threads=threads, leaf_stride=stride)
    xq, yq = synth.mixed_queries(20000, xa, ya, golden["cfg3_leaf_rows"])
    st, zq, leaf = w.eval_e(xq, yq)
    assert st == g.SUCCESS
    assert np.array_equal(leaf, golden["cfg3_leaf"])
    assert refimpl.same_bits(zq, golden["cfg3_val"])
    assert w.info()["n_deferred"] > 0
    w.free()


def test_reorder_pipeline_edge_cases(golden):
    """non-finite queries, every query in one tile (run length = chunk), tiny deferred list
    (overflow bitmask), windows that do not divide the batch, outside-cage status"""
    xa, ya = synth.uniform_nodes(1000)
    za = synth.values(1000, 1)
    order = g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000)
    port = refimpl.PortTree({k: golden["cfg1_" + k] for k in
                             ("pts3", "lnk3", "type", "shuffle", "shift", "scale", "seed_points")}, xa, ya)
    w = make(xa, ya, za, g.DEFAULT, order, reorder=1, reorder_window=7001, tile_leaves=20)
    bx = np.array([np.nan, np.inf, -np.inf, 0.5, 1e300, 0.25])
    by = np.array([0.5, 0.5, 0.5, np.nan, 0.5, 0.75])
    st, zq, leaf = w.eval_e(bx, by)
    pl, pv, _ = port.eval(bx, by, za)
    assert np.array_equal(leaf, pl) and refimpl.same_bits(zq, pv)
    # one hot spot: 30000 copies of one point plus the golden set, window 7001 (4 full + 1 ragged)
    xq, yq = synth.uniform_queries(10000)
    hx = np.concatenate([np.full(30000, xq[17]), xq])
    hy = np.concatenate([np.full(30000, yq[17]), yq])
    st, zq, leaf = w.eval_e(hx, hy)
    assert st == g.SUCCESS
    assert np.all(leaf[:30000] == golden["cfg1_leaf"][17])
    assert np.array_equal(leaf[30000:], golden["cfg1_leaf"])
    assert refimpl.same_bits(zq[:, 30000:], golden["cfg1_val"])
    assert np.all(zq[0, :30000].view(np.uint64) == golden["cfg1_val"].ravel().view(np.uint64)[17])
    # outside the cage: leaf -1, NaN, EDOM
    st, zq, leaf = w.eval_e(np.array([0.5, 1e6]), np.array([0.5, 0.0]))
    assert st == g.EDOM and leaf[1] == -1 and np.isnan(zq[0, 1]) and leaf[0] >= 0
    w.free()
    w = make(xa, ya, za, g.DEFAULT, order, reorder=1, cells_per_leaf=0.05, max_steps=1, list_fraction=1e-4)
    xq2, yq2 = np.tile(xq, 20), np.tile(yq, 20)
    st, zq, leaf = w.eval_e(xq2, yq2)
    assert np.array_equal(leaf, np.tile(golden["cfg1_leaf"], 20))
    assert refimpl.same_bits(zq, np.tile(golden["cfg1_val"], (1, 20)))
    assert w.info()["n_deferred"] > 65536
    w.free()


@pytest.mark.parametrize("n,nq,window", [(100000, 10_000_000, 3_000_001), (100000, 10_000_000, 1 << 27),
                                         (1_000_000, 20_000_000, 1 << 27)])
def test_reorder_pipeline_scale_vs_exact_and_direct(n, nq, window):
    """10^5 nodes, 10^7 device-generated queries: reorder pipeline == direct fast path == exact
    DAG replay on 100 % of the queries (leaf ids and value bits), plus the port on a subsample."""
    torch = _torch()
    xa, ya = synth.uniform_nodes(n)
    za = synth.values(n, 1)
    w = make(xa, ya, za, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n), reorder=1, reorder_window=window)
    dev = torch.device("cuda:0")
    dx = torch.empty(nq, dtype=torch.float64, device=dev)
    dy = torch.empty(nq, dtype=torch.float64, device=dev)
    w.generate_uniform_device(0, synth.SEED_QUERIES, 0, nq, 0.0, 1.0, 0.0, 1.0, dx.data_ptr(), dy.data_ptr())
    res = {}
    for name, params in (("reorder", {"reorder": 1}), ("direct", {"reorder": 0}), ("exact", {"mode": 1})):
        for k, v in params.items():
            w.set_param(k, v)
        z = torch.empty(nq, dtype=torch.float64, device=dev)
        l = torch.empty(nq, dtype=torch.int32, device=dev)
        assert w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), nq, z.data_ptr(), nq, l.data_ptr()) == 0
        res[name] = (z, l)
    assert w.info()["n_reordered"] == nq
    for other in ("direct", "exact"):
        bad_leaf, bad_val = w.compare_device(0, res["reorder"][1].data_ptr(), res[other][1].data_ptr(),
                                             res["reorder"][0].data_ptr(), res[other][0].data_ptr(), nq, nq)
        assert bad_leaf == 0 and bad_val == 0, other
    sub = torch.arange(0, nq, 997, device=dev)
    port = refimpl.PortTree(w.get_tree(w.n_simplexes(), n), xa, ya)
    pl, pv, _ = port.eval(dx[sub].cpu().numpy(), dy[sub].cpu().numpy(), za)
    assert np.array_equal(res["reorder"][1][sub].cpu().numpy(), pl)
    assert refimpl.same_bits(res["reorder"][0][sub].cpu().numpy().reshape(1, -1), pv)
    w.free()


def test_slab_save_load_round_trip(golden, tmp_path):
    """the flattened triangulation written to a file and restored into a
    fresh workspace (no host build) evaluates bit-identically; a flipped byte is detected."""
    xa, ya = synth.uniform_nodes(1000)
    w = make(xa, ya, synth.values(1000, 1), g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000))
    path = tmp_path / "cfg1.gsltri"
    w.save(path)
    assert path.stat().st_size == w.info()["slab_bytes"]
    w.free()
    w2 = g.ScatteredCuda(1000, 1)
    w2.load(path)
    xq, yq = synth.uniform_queries(10000)
    for mode in (0, 1):
        w2.set_param("mode", mode)
        st, zq, leaf = w2.eval_e(xq, yq)
        assert st == g.SUCCESS
        assert np.array_equal(leaf, golden["cfg1_leaf"]) and refimpl.same_bits(zq, golden["cfg1_val"])
    x0, dx, nx, y0, dy, ny = golden["cfg1_raster_spec"]
    st, rz, rl = w2.eval_raster_e(x0, dx, int(nx), y0, dy, int(ny))
    assert np.array_equal(rl, golden["cfg1_raster_leaf"]) and refimpl.same_bits(rz, golden["cfg1_raster_val"])
    info = w2.info()
    assert info["n_leaves"] == 2001 and info["n_simplexes"] == 8967
    w2.free()
    raw = bytearray(path.read_bytes())
    raw[len(raw) // 2] ^= 0x40
    bad = tmp_path / "bad.gsltri"
    bad.write_bytes(bytes(raw))
    w3 = g.ScatteredCuda(1000, 1)
    with pytest.raises(g.GslError) as e:
        w3.load(bad)
    assert e.value.status == g.EINVAL
    w3.free()
    w4 = g.ScatteredCuda(1000, 2)                       # n_fields differs from the image
    with pytest.raises(g.GslError) as e:
        w4.load(path)
    assert e.value.status == g.EBADLEN
    w4.free()


def test_exact_replay_entry_nodes_are_transparent_on_device(golden):
    """k_entry_cells: the exact replay started at the per-cell entry node equals the replay from
    the root on 100 % of 10^7 device-generated queries (10^5 nodes) and on the config-3 mix
    (on-vertex / on-edge / outside-hull queries), and the golden vectors hold in both modes."""
    torch = _torch()
    n, nq = 100000, 10_000_000
    xa, ya = synth.uniform_nodes(n)
    w = make(xa, ya, synth.values(n, 1), g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n), mode=1)
    dev = torch.device("cuda:0")
    dx = torch.empty(nq, dtype=torch.float64, device=dev)
    dy = torch.empty(nq, dtype=torch.float64, device=dev)
    # queries over an inflated box: cage leaves and the clamped border cells are exercised too
    w.generate_uniform_device(0, synth.SEED_QUERIES, 0, nq, -0.1, 1.1, -0.1, 1.1, dx.data_ptr(), dy.data_ptr())
    out = []
    for entry in (1, 0):
        w.set_param("exact_entry", entry)
        z = torch.empty(nq, dtype=torch.float64, device=dev)
        l = torch.empty(nq, dtype=torch.int32, device=dev)
        assert w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), nq, z.data_ptr(), nq, l.data_ptr()) == 0
        out.append((z, l))
    bad_leaf, bad_val = w.compare_device(0, out[0][1].data_ptr(), out[1][1].data_ptr(),
                                         out[0][0].data_ptr(), out[1][0].data_ptr(), nq, nq)
    assert bad_leaf == 0 and bad_val == 0
    w.free()
    n = 3000
    xa, ya = synth.clustered_nodes(n)
    xq, yq = synth.mixed_queries(20000, xa, ya, golden["cfg3_leaf_rows"])
    for entry in (1, 0):
        w = make(xa, ya, synth.values(n, 3), g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n),
                 mode=1, exact_entry=entry)
        st, zq, leaf = w.eval_e(xq, yq)
        assert np.array_equal(leaf, golden["cfg3_leaf"]) and refimpl.same_bits(zq, golden["cfg3_val"])
        w.free()


@pytest.mark.parametrize("kernel", [2, 0])
@pytest.mark.parametrize("F,nx,ny", [(1, 1024, 2048), (16, 512, 1024), (4, 300, 1028), (2, 257, 516)])
def test_raster_kernel_vs_exact_replay_on_device(F, nx, ny, kernel):
    """the raster kernels (k_locate by default, k_walk on request) against the exact DAG replay on
    every pixel, for a raster over a box that sticks out of the data hull (cage leaves, cage
    vertices in the blend) and a fine one inside it."""
    torch = _torch()
    n = 100000
    xa, ya = synth.uniform_nodes(n)
    w = make(xa, ya, synth.values(n, F), g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n), kernel=kernel)
    dev = torch.device("cuda:0")
    for (x0, x1, y0, y1) in [(-0.02, 1.02, -0.02, 1.02), (0.4, 0.45, 0.3, 0.35)]:
        dx, dy = (x1 - x0) / nx, (y1 - y0) / ny
        res = []
        for mode in (0, 1):
            w.set_param("mode", mode)
            z = torch.empty((F, nx * ny), dtype=torch.float64, device=dev)
            l = torch.empty(nx * ny, dtype=torch.int32, device=dev)
            assert w.eval_raster_device_e(0, x0, dx, nx, y0, dy, ny, z.data_ptr(), l.data_ptr()) == 0
            res.append((z, l))
        w.set_param("mode", 0)
        bad_leaf, bad_val = w.compare_device(0, res[0][1].data_ptr(), res[1][1].data_ptr(),
                                             res[0][0].data_ptr(), res[1][0].data_ptr(), nx * ny, F * nx * ny)
        assert bad_leaf == 0 and bad_val == 0, (x0, x1)
    w.free()


def test_eval_is_reentrant_across_host_threads(golden):
    """eval* may be called concurrently from several host threads on one
    workspace -- every call in flight gets its own evaluation context (streams, deferred list,
    counters, staging).  Four threads hammer eval_e / eval_raster_e / eval_deriv_e on one
    workspace; every result must be the golden one and the outside-cage status must stay with the
    call that caused it."""
    import threading
    xa, ya = synth.uniform_nodes(1000)
    w = make(xa, ya, synth.values(1000, 1), g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000), host_chunk=1500)
    xq, yq = synth.uniform_queries(10000)
    x0, dx, nx, y0, dy, ny = golden["cfg1_raster_spec"]
    # the NpzFile reads lazily and is not thread-safe: take the arrays out before the threads start
    golden = {k: np.array(golden[k]) for k in ("cfg1_leaf", "cfg1_val", "cfg1_raster_leaf", "cfg1_raster_val")}
    bad_x, bad_y = xq.copy(), yq.copy()
    bad_x[123] = 1e6                                     # one query outside the cage: this call alone reports EDOM
    errors = []

    def worker(kind):
        try:
            for _ in range(6):
                if kind == 0:
                    st, zq, leaf = w.eval_e(xq, yq)
                    assert st == g.SUCCESS and np.array_equal(leaf, golden["cfg1_leaf"])
                    assert refimpl.same_bits(zq, golden["cfg1_val"])
                elif kind == 1:
                    st, rz, rl = w.eval_raster_e(x0, dx, int(nx), y0, dy, int(ny))
                    assert st == g.SUCCESS and np.array_equal(rl, golden["cfg1_raster_leaf"])
                    assert refimpl.same_bits(rz, golden["cfg1_raster_val"])
                elif kind == 2:
                    st, zq, leaf = w.eval_e(bad_x, bad_y)
                    assert st == g.EDOM and leaf[123] == -1 and np.isnan(zq[0, 123])
                    keep = np.arange(10000) != 123
                    assert np.array_equal(leaf[keep], golden["cfg1_leaf"][keep])
                else:
                    st, ddx, ddy = w.eval_deriv_e(xq, yq)
                    assert st == g.SUCCESS and np.isfinite(ddx).all()
        except Exception as e:                            # surfaced in the main thread below
            errors.append((kind, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in (0, 1, 2, 3, 0)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    info = w.info()
    assert info["n_contexts"] >= 2                       # the calls really overlapped
    assert info["n_outside"] == 6                        # counters of all contexts are folded into get_info
    w.free()


def test_pageable_host_arrays_are_staged_through_pinned_buffers(golden):
    """plain numpy (malloc) arrays above the pageable threshold go through the library's pinned
    staging in several chunks, 3 fields, ragged last chunk; same bits as the small direct call"""
    n = 3000
    xa, ya = synth.clustered_nodes(n)
    za = synth.values(n, 3)
    w = make(xa, ya, za, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n), host_chunk=70000, copy_threads=3,
             pageable_min=100000)
    xq, yq = synth.mixed_queries(20000, xa, ya, golden["cfg3_leaf_rows"])
    reps = 16                                            # 320000 queries: 4 full chunks + a ragged one
    bx, by = np.tile(xq, reps), np.tile(yq, reps)
    st, zq, leaf = w.eval_e(bx, by)
    assert st == g.SUCCESS
    assert np.array_equal(leaf, np.tile(golden["cfg3_leaf"], reps))
    assert refimpl.same_bits(zq, np.tile(golden["cfg3_val"], (1, reps)))
    st, _, leaf = w.eval_e(bx, by, want_values=False)    # leaf ids only
    assert np.array_equal(leaf, np.tile(golden["cfg3_leaf"], reps))
    st, zq, _ = w.eval_e(bx, by, want_leaf=False)        # values only
    assert refimpl.same_bits(zq, np.tile(golden["cfg3_val"], (1, reps)))
    w.free()


def test_division_by_hoisted_reciprocal():
    """csrc/gsi_core.h: gsi_div_rcp / gsi_div_by_rcp are nvcc's own binary64 division split where it
    stops depending on the divisor only.  On 10^9 operand pairs -- data-like magnitudes, raw random
    bit patterns (NaN, inf, denormals), mixed, and the whole exponent range -- every quotient the
    hoisted form vouches for must be bit-identical to __ddiv_rn; the rest is deferred by design."""
    import ctypes
    from gsl_scattered_interpolation_b200 import _lib
    L = _lib.load()
    out = (ctypes.c_ulonglong * 2)()
    n = 1_000_000_000
    assert L.gsl_scattered_cuda_selftest_division(0, 20260004, n, out) == g.SUCCESS
    vouched, mismatches = int(out[0]), int(out[1])
    assert mismatches == 0
    assert vouched > 0.3 * n                    # the data-like family is never deferred, the wild ones mostly are


@pytest.mark.parametrize("kind,n", [("uniform", 300_000), ("cluster", 200_000)])
def test_builder_with_device_located_points_gives_the_host_tree(kind, n):
    """the builder's point locations in batches on the GPU (k_build_locate on
    snapshots of the growing tree, the host finishing each descent) give the tree of the host-only
    build node for node; the evaluation of both is identical too"""
    xa, ya = synth.uniform_nodes(n, seed=5) if kind == "uniform" else synth.clustered_nodes(n, seed=5)
    order = g.shuffle_mt19937(11, n)
    trees, stats = [], []
    for dev in (1, 0):
        w = g.ScatteredCuda(n, 1)
        w.set_param("build_device", dev)
        w.build(xa, ya, g.DEFAULT, order)
        trees.append(w.get_tree(w.n_simplexes(), n))
        stats.append(w.build_stats())
        w.free()
    assert stats[0]["batched_points"] == n - 4096 and stats[0]["rounds"] >= 6
    assert stats[1]["batched_points"] == 0
    for k in ("pts3", "lnk3", "type", "shuffle", "shift", "scale", "seed_points"):
        assert np.array_equal(trees[0][k], trees[1][k]), k


def test_walk_trace_diagnostic(golden):
    """per-CTA trace of the last fast-path launch: every query handed out exactly once"""
    n, nq = 20000, 1_000_000
    xa, ya = synth.uniform_nodes(n)
    w = make(xa, ya, synth.values(n, 1), g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n), walk_trace=1)
    xq, yq = synth.uniform_queries(nq)
    w.eval_e(xq, yq)
    tr = w.walk_trace(0)
    tr = tr[tr[:, 2] > 0]                               # CTAs of the launch
    assert tr.shape[0] > 0 and int(tr[:, 3].sum()) == nq
    assert (tr[:, 2] >= tr[:, 1]).all() and (tr[:, 0] < 200).all()
    w.free()
