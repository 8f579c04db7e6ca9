"""Generate tests/golden/golden_v1.npz from the REAL reference (oracle/_ref, built from
/root/reference sources by oracle/Makefile).  Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden.py

Everything stored is an output of the reference's own simplex_tree_init / find_leaf /
interp_point (or the 50-station fixture of its example program); synthetic inputs are
regenerated deterministically by gsl_scattered_interpolation_b200/synth.py at test time.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refimpl  # noqa: E402
from gsl_scattered_interpolation_b200 import synth  # noqa: E402


def tree_digest(t):
    h = hashlib.sha256()
    for k in ("pts3", "lnk3", "type"):
        h.update(np.ascontiguousarray(t[k], dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(t["shuffle"], dtype=np.uint64).tobytes())
    for k in ("shift", "scale", "seed_points"):
        h.update(np.ascontiguousarray(t[k], dtype=np.float64).tobytes())
    return h.hexdigest()


def main():
    out = {}
    w = refimpl.weather_data()
    out["weather"] = w
    xa, ya, za = w[:, 0].copy(), w[:, 1].copy(), w[:, 2].copy()

    # tree B of the example: standardised, gsl_rng_default (mt19937, seed 0); SURVEY section 4
    tb = refimpl.RefTree(xa, ya, 0, seed=0)
    eb = tb.export()
    for k, v in eb.items():
        out["weatherB_" + k] = v
    q = np.array([(-88, 41), (-88.0957, 42.0915), (-87.5, 42), (-86.303, 40.9479), (-89.6763, 43.2),
                  (-87.9067, 41.9816), (-95, 41), (-88, 60), (1e6, 0), (np.nan, 1.0)], dtype=float)
    leaf, val, nout = tb.eval(q[:, 0], q[:, 1], za)
    out["weatherB_q"], out["weatherB_leaf"], out["weatherB_val"] = q, leaf, val
    # the example's 100 x 100 raster (scattered_interp_example.c:166-199)
    mn, mx = (-89.6763, 40.9479), (-86.303, 43.20)
    n_grid = 100
    xstep, ystep = (mx[0] - mn[0]) / n_grid, (mx[1] - mn[1]) / n_grid
    rl, rv, _ = tb.eval_raster(mn[0], xstep, n_grid, mn[1], ystep, n_grid, za)
    out["weatherB_raster_spec"] = np.array([mn[0], xstep, n_grid, mn[1], ystep, n_grid])
    out["weatherB_raster_leaf"], out["weatherB_raster_val"] = rl, rv

    # tree A of the example: NOSTANDARDIZE, insertion in file order
    ta = refimpl.RefTree(xa, ya, 1, seed=None)
    ea = ta.export()
    for k, v in ea.items():
        out["weatherA_" + k] = v
    leaf, val, _ = ta.eval(q[:8, 0], q[:8, 1], za)
    out["weatherA_leaf"], out["weatherA_val"] = leaf, val

    # trivial_test known answers (scattered_interp_example.c:25-81)
    t0 = refimpl.RefTree(np.zeros(0), np.zeros(0), 1, seed=None)
    leaf, val, _ = t0.eval(np.array([-88.0]), np.array([41.0]), np.zeros((1, 1)))
    out["trivial_empty_leaf"], out["trivial_empty_val"] = leaf, val
    t1 = refimpl.RefTree(np.array([-88.0]), np.array([41.0]), 1, seed=None)
    e1 = t1.export()
    out["trivial_one_pts3"], out["trivial_one_lnk3"], out["trivial_one_type"] = e1["pts3"], e1["lnk3"], e1["type"]
    leaf, val, _ = t1.eval(np.array([-89.0]), np.array([41.0]), np.array([[5.0]]))
    out["trivial_second_leaf"], out["trivial_second_val"] = leaf, val

    # config 1: 1000 uniform nodes, 10^4 uniform queries, mt19937(12345) insertion order
    xa, ya = synth.uniform_nodes(1000)
    za = synth.values(1000, 1)
    t = refimpl.RefTree(xa, ya, 0, seed=synth.SEED_SHUFFLE)
    e = t.export()
    out["cfg1_digest"] = np.array(tree_digest(e))
    out["cfg1_n_simplexes"] = np.array(t.n_simplexes)
    for k in ("pts3", "lnk3", "type", "shuffle", "shift", "scale", "seed_points"):
        out["cfg1_" + k] = e[k]
    xq, yq = synth.uniform_queries(10000)
    leaf, val, _ = t.eval(xq, yq, za)
    out["cfg1_leaf"], out["cfg1_val"] = leaf, val
    # 100 x 100 raster over the data bbox, the example's own query shape
    x0, x1, y0, y1 = xa.min(), xa.max(), ya.min(), ya.max()
    dx, dy = (x1 - x0) / 100, (y1 - y0) / 100
    rl, rv, _ = t.eval_raster(x0, dx, 100, y0, dy, 100, za)
    out["cfg1_raster_spec"] = np.array([x0, dx, 100, y0, dy, 100])
    out["cfg1_raster_leaf"], out["cfg1_raster_val"] = rl, rv
    # ISOSCALE variant
    ti = refimpl.RefTree(xa * 3.0, ya, 2, seed=synth.SEED_SHUFFLE)
    out["cfg1_iso_digest"] = np.array(tree_digest(ti.export()))
    leaf, val, _ = ti.eval(xq * 3.0, yq, za)
    out["cfg1_iso_leaf"], out["cfg1_iso_val"] = leaf, val

    # config-3 style: clustered anisotropic nodes, mixed queries incl. on-vertex, on-edge,
    # outside hull, far outside; 3 fields
    n = 3000
    xa, ya = synth.clustered_nodes(n)
    za = synth.values(n, 3)
    t = refimpl.RefTree(xa, ya, 0, seed=synth.SEED_SHUFFLE)
    e = t.export()
    out["cfg3_digest"] = np.array(tree_digest(e))
    out["cfg3_n_simplexes"] = np.array(t.n_simplexes)
    leaves = np.where(e["type"] == 0)[0]
    lv = e["pts3"].reshape(-1, 3)[leaves]
    lv = lv[(lv >= 0).all(axis=1)]
    rows = e["shuffle"][lv].astype(np.int64)  # data rows of real-leaf vertices
    out["cfg3_leaf_rows"] = rows
    xq, yq = synth.mixed_queries(20000, xa, ya, rows)
    leaf, val, nout = t.eval(xq, yq, za)
    assert nout == 0
    out["cfg3_leaf"], out["cfg3_val"] = leaf, val

    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")
    print("cfg1 nodes", int(out["cfg1_n_simplexes"]), "cfg3 nodes", int(out["cfg3_n_simplexes"]))


if __name__ == "__main__":
    main()
