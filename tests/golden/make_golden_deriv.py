"""Generate tests/golden/golden_v2_deriv.npz: derivative evaluators.

The reference has no derivative entry point; the oracle-side definition is
oracle/ref_driver.c:ref_eval_deriv, a few lines on top of the REAL reference's
calculate_bary_coords + gsl_linalg_LU_solve + interp_point's blend loop (see the comment there).
This script freezes what that definition returns, computed by oracle/_ref (the reference
compiled in place), on the weather fixture, config 1 and the config-3 mix.  Build container only:

    make -C oracle ref && python tests/golden/make_golden_deriv.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refimpl  # noqa: E402
from gsl_scattered_interpolation_b200 import synth  # noqa: E402


def main():
    g1 = np.load(os.path.join(HERE, "golden_v1.npz"))
    out = {}
    w = refimpl.weather_data()
    xa, ya, za = w[:, 0].copy(), w[:, 1].copy(), w[:, 2].copy()
    tb = refimpl.RefTree(xa, ya, 0, seed=0)
    q = g1["weatherB_q"]
    leaf, dx, dy = tb.eval_deriv(q[:, 0], q[:, 1], za)
    assert np.array_equal(leaf, g1["weatherB_leaf"])
    out["weatherB_dzdx"], out["weatherB_dzdy"] = dx, dy

    xa, ya = synth.uniform_nodes(1000)
    za = synth.values(1000, 1)
    t = refimpl.RefTree(xa, ya, 0, seed=synth.SEED_SHUFFLE)
    xq, yq = synth.uniform_queries(10000)
    leaf, dx, dy = t.eval_deriv(xq, yq, za)
    assert np.array_equal(leaf, g1["cfg1_leaf"])
    out["cfg1_dzdx"], out["cfg1_dzdy"] = dx, dy

    n = 3000
    xa, ya = synth.clustered_nodes(n)
    za = synth.values(n, 3)
    t = refimpl.RefTree(xa, ya, 0, seed=synth.SEED_SHUFFLE)
    xq, yq = synth.mixed_queries(20000, xa, ya, g1["cfg3_leaf_rows"])
    leaf, dx, dy = t.eval_deriv(xq, yq, za)
    assert np.array_equal(leaf, g1["cfg3_leaf"])
    out["cfg3_dzdx"], out["cfg3_dzdy"] = dx, dy

    path = os.path.join(HERE, "golden_v2_deriv.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
