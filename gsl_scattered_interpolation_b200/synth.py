"""Deterministic synthetic workloads, numpy side.

u(seed, i) = (splitmix64(seed + (i+1)*0x9E3779B97F4A7C15) >> 11) * 2^-53 -- the same pure
integer -> double map as gsi_uniform() in csrc/gsi_core.h, so host- and device-generated
inputs are bit-identical.
"""
import numpy as np

SEED_NODES, SEED_VALUES, SEED_QUERIES, SEED_SHUFFLE = 20260001, 20260002, 20260003, 12345
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def uniform(seed, idx):
    idx = np.asarray(idx, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (idx + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform_nodes(n, seed=SEED_NODES):
    i = np.arange(n, dtype=np.uint64)
    return uniform(seed, 2 * i), uniform(seed, 2 * i + 1)


def values(n, n_fields=1, seed=SEED_VALUES):
    """field-major (n_fields, n): z_f(i) = u(seed, n_fields*i + f)"""
    i = np.arange(n, dtype=np.uint64)
    return np.stack([uniform(seed, np.uint64(n_fields) * i + np.uint64(f)) for f in range(n_fields)])


def uniform_queries(nq, first=0, seed=SEED_QUERIES, box=(0.0, 1.0, 0.0, 1.0)):
    k = np.arange(first, first + nq, dtype=np.uint64)
    xlo, xhi, ylo, yhi = box
    return xlo + (xhi - xlo) * uniform(seed, 2 * k), ylo + (yhi - ylo) * uniform(seed, 2 * k + 1)


def clustered_nodes(n, n_clusters=64, seed=SEED_NODES + 100):
    """config 3: Gaussian clusters, anisotropic bounding box (x range ~1000, y range ~1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cx = rng.uniform(0, 1000, n_clusters)
    cy = rng.uniform(0, 1, n_clusters)
    sx = rng.uniform(1, 30, n_clusters)
    sy = rng.uniform(1e-3, 3e-2, n_clusters)
    c = rng.integers(0, n_clusters, n)
    return cx[c] + sx[c] * rng.standard_normal(n), cy[c] + sy[c] * rng.standard_normal(n)


def mixed_queries(nq, xa, ya, leaf_vertices=None, seed=SEED_QUERIES + 100, with_class=False):
    """config 3 query mix: 60 % uniform in the bbox, 15 % in the bbox inflated by 25 % per
    side, 5 % far outside (+-50x extent, still inside the cage), 10 % exact node copies,
    10 % edge midpoints 0.5*(a+b) of leaf edges (or of random node pairs)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x0, x1, y0, y1 = xa.min(), xa.max(), ya.min(), ya.max()
    w, h = x1 - x0, y1 - y0
    n1, n2, n3, n4 = int(0.6 * nq), int(0.15 * nq), int(0.05 * nq), int(0.1 * nq)
    n5 = nq - n1 - n2 - n3 - n4
    xs = [rng.uniform(x0, x1, n1), rng.uniform(x0 - 0.25 * w, x1 + 0.25 * w, n2),
          rng.uniform(x0 - 50 * w, x1 + 50 * w, n3)]
    ys = [rng.uniform(y0, y1, n1), rng.uniform(y0 - 0.25 * h, y1 + 0.25 * h, n2),
          rng.uniform(y0 - 50 * h, y1 + 50 * h, n3)]
    k = rng.integers(0, xa.size, n4)
    xs.append(xa[k]); ys.append(ya[k])
    if leaf_vertices is not None and len(leaf_vertices):
        lv = leaf_vertices[rng.integers(0, len(leaf_vertices), n5)]
        e = rng.integers(0, 3, n5)
        a = lv[np.arange(n5), e]; b = lv[np.arange(n5), (e + 1) % 3]
    else:
        a = rng.integers(0, xa.size, n5); b = rng.integers(0, xa.size, n5)
    xs.append(0.5 * (xa[a] + xa[b])); ys.append(0.5 * (ya[a] + ya[b]))
    xq, yq = np.concatenate(xs), np.concatenate(ys)
    p = rng.permutation(nq)
    if with_class:
        # 0 = in bbox, 1 = inflated bbox, 2 = far outside, 3 = on a node, 4 = on an edge midpoint
        cls = np.repeat(np.arange(5, dtype=np.int8), [n1, n2, n3, n4, n5])
        return np.ascontiguousarray(xq[p]), np.ascontiguousarray(yq[p]), cls[p]
    return np.ascontiguousarray(xq[p]), np.ascontiguousarray(yq[p])
