"""Host builder (csrc/gsi_builder.cpp) must reproduce the reference's history DAG node for
node: vertex ids, links, types, shuffle, shift/scale and cage seed points bit-identical."""
import hashlib

import numpy as np
import pytest

import gsl_scattered_interpolation_b200 as g
import refimpl
from gsl_scattered_interpolation_b200 import synth


def digest(t):
    h = hashlib.sha256()
    for k in ("pts3", "lnk3", "type"):
        h.update(np.ascontiguousarray(t[k], dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(t["shuffle"], dtype=np.uint64).tobytes())
    for k in ("shift", "scale", "seed_points"):
        h.update(np.ascontiguousarray(t[k], dtype=np.float64).tobytes())
    return h.hexdigest()


def build(xa, ya, flags, order):
    w = g.ScatteredCuda(len(xa), 1)
    w.build(xa, ya, flags, order)
    t = w.get_tree(w.n_simplexes(), len(xa))
    w.free()
    return t


def test_shuffle_matches_gsl_ran_shuffle(golden):
    assert np.array_equal(g.shuffle_mt19937(0, 50), golden["weatherB_shuffle"])
    assert np.array_equal(g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000), golden["cfg1_shuffle"])


def test_builder_weather_trees(golden):
    w = golden["weather"]
    tb = build(w[:, 0], w[:, 1], g.DEFAULT, g.shuffle_mt19937(0, 50))
    ta = build(w[:, 0], w[:, 1], g.NOSTANDARDIZE, None)
    for k in ("pts3", "lnk3", "type", "shuffle", "shift", "scale", "seed_points"):
        assert np.array_equal(tb[k], golden["weatherB_" + k]), k
        assert np.array_equal(ta[k], golden["weatherA_" + k]), k
    assert tb["type"].size == 369 and ta["type"].size == 429


def test_builder_trivial(golden):
    t = build(np.array([-88.0]), np.array([41.0]), g.NOSTANDARDIZE, None)
    assert np.array_equal(t["pts3"], golden["trivial_one_pts3"])
    assert np.array_equal(t["lnk3"], golden["trivial_one_lnk3"])
    assert np.array_equal(t["type"], golden["trivial_one_type"])
    t0 = build(np.zeros(0), np.zeros(0), g.NOSTANDARDIZE, None)
    assert t0["type"].size == 1 and list(t0["pts3"]) == [-1, -2, -3]


def test_builder_config1_and_config3_digests(golden):
    xa, ya = synth.uniform_nodes(1000)
    t = build(xa, ya, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000))
    assert t["type"].size == int(golden["cfg1_n_simplexes"])
    assert digest(t) == str(golden["cfg1_digest"])
    ti = build(xa * 3.0, ya, g.ISOSCALE, g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000))
    assert digest(ti) == str(golden["cfg1_iso_digest"])
    xa, ya = synth.clustered_nodes(3000)
    t = build(xa, ya, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, 3000))
    assert t["type"].size == int(golden["cfg3_n_simplexes"])
    assert digest(t) == str(golden["cfg3_digest"])


def test_builder_errors():
    w = g.ScatteredCuda(10, 1)
    with pytest.raises(g.GslError) as e:
        w.build(np.zeros(11), np.zeros(11))          # more nodes than allocated
    assert e.value.status == g.EBADLEN
    with pytest.raises(g.GslError) as e:
        w.build(np.arange(5.0), np.arange(5.0), order=np.array([0, 1, 2, 3, 9]))
    assert e.value.status == g.EFAILED
    w.free()


@pytest.mark.skipif(not refimpl.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("n,flags,seed,kind", [(50, 0, 1, "uniform"), (5000, 0, 12345, "uniform"),
                                               (5000, 1, None, "uniform"), (4000, 2, 8, "cluster"),
                                               (25, 1, None, "grid"), (20000, 0, 4, "cluster")])
def test_builder_matches_reference(n, flags, seed, kind):
    if kind == "uniform":
        xa, ya = synth.uniform_nodes(n, seed=1000 + n)
    elif kind == "cluster":
        xa, ya = synth.clustered_nodes(n, seed=1000 + n)
    else:                                              # gridded_test(2,2) of the example
        xa, ya = (np.arange(25) % 5).astype(float), (np.arange(25) // 5).astype(float)
    rt = refimpl.RefTree(xa, ya, flags, seed)
    ref = rt.export()
    mine = build(xa, ya, flags, None if seed is None else g.shuffle_mt19937(seed, n))
    assert mine["type"].size == rt.n_simplexes
    for k in ref:
        assert np.array_equal(mine[k], ref[k]), k
    rt.close()


# ---- the batched-locate protocol of the builder: gsi_build_locator with the
# descents run on the host over a copy of every snapshot must give the plain build's tree node for
# node (the CUDA locator, k_build_locate, is compared the same way in tests/test_gpu_parity.py)
@pytest.mark.parametrize("kind,n,flags,seed", [("uniform", 20000, 0, 5), ("cluster", 30000, 0, 9),
                                               ("grid", 10000, 0, 3), ("grid", 8100, 1, None),
                                               ("collinear", 6000, 2, 4), ("uniform", 4097, 0, 1)])
def test_batched_locate_protocol_gives_the_same_tree(kind, n, flags, seed):
    import hostsim_py
    if kind == "uniform":
        xa, ya = synth.uniform_nodes(n, seed=77 + n)
    elif kind == "cluster":
        xa, ya = synth.clustered_nodes(n, seed=77 + n)
    elif kind == "grid":                               # cocircular quadruples everywhere: ties and fallbacks
        s = int(round(n ** 0.5))
        rng = np.random.default_rng(n)
        xa, ya = (np.arange(s * s) % s).astype(float), (np.arange(s * s) // s).astype(float)
        xa, ya = xa + 1e-6 * rng.standard_normal(s * s), ya + 1e-6 * rng.standard_normal(s * s)
    else:                                              # long collinear runs
        rng = np.random.default_rng(n)
        xa = np.repeat(np.arange(n // 6, dtype=float), 6)
        ya = rng.integers(0, 40, size=xa.size).astype(float)
        xa, ya = np.unique(np.stack([xa, ya], axis=1), axis=0).T.copy()
        xa = xa + 1e-7 * rng.standard_normal(xa.size)   # exactly collinear triples are rejected by the builder
    order = None if seed is None else g.shuffle_mt19937(seed, xa.size)
    try:
        same, batched, rounds = hostsim_py.build_batched_equal(xa, ya, flags, order)
    except RuntimeError:
        pytest.skip("the plain build rejects this point set (as the reference's builder does)")
    assert same
    assert batched == xa.size - 4096 and rounds >= 1
