#!/usr/bin/env python
"""Secondary measurements: BASELINE.json configs[1..3] (the parity-test configurations) timed on
one B200 through the C ABI's device entry points, each verified in the same run (fast path ==
exact DAG replay on 100 % of the queries on the device; CPU oracle on a subsample).

    python scripts/bench_configs.py [2] [3] [4]      -> one JSON line per config on stdout
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import gsl_scattered_interpolation_b200 as g  # noqa: E402
import refimpl  # noqa: E402
from gsl_scattered_interpolation_b200 import synth  # noqa: E402

DEV = torch.device("cuda:0")
PEAK = 6446.0
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def split_ms(w, launches):
    """per-launch device time of the fast-path kernel and of the exact-replay kernel"""
    inf = w.info()
    return {"locate_ms": 1e3 * inf["locate_seconds"] / launches, "exact_ms": 1e3 * inf["exact_seconds"] / launches}


def timed(fn, steps=5, warmup=3):
    st = torch.cuda.Stream(device=DEV)
    for _ in range(warmup):
        fn(st.cuda_stream)
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        fn(st.cuda_stream)
    e1.record(st)
    st.synchronize()
    return e0.elapsed_time(e1) / steps


def build(xa, ya, za, **params):
    za = np.atleast_2d(za)
    n = len(xa)
    w = g.ScatteredCuda(n, za.shape[0])
    w.set_param("profile", 1)
    for kv in filter(None, os.environ.get("GSI_PARAMS", "").split(",")):
        k, v = kv.split("=")
        params[k] = float(v)
    for k, v in params.items():
        w.set_param(k, v)
    t0 = time.perf_counter()
    w.build(xa, ya, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
    t1 = time.perf_counter()
    w.upload(za)
    t2 = time.perf_counter()
    return w, t1 - t0, t2 - t1


def verify_arrays(w, dx, dy, nq, F, dz, dl):
    vz = torch.empty((F, nq), dtype=torch.float64, device=DEV)
    vl = torch.empty(nq, dtype=torch.int32, device=DEV)
    w.set_param("mode", 1)
    w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), nq, vz.data_ptr(), nq, vl.data_ptr())
    w.set_param("mode", 0)
    return w.compare_device(0, dl.data_ptr(), vl.data_ptr(), dz.data_ptr(), vz.data_ptr(), nq, F * nq)


def oracle_subsample(w, xa, ya, za, xs, ys, leaf_s, val_s):
    n = len(xa)
    port = refimpl.PortTree(w.get_tree(w.n_simplexes(), n), xa, ya)
    pl, pv, _ = port.eval(xs, ys, za)
    return int((pl != leaf_s).sum()), (0 if refimpl.same_bits(pv, val_s) else 1)


def config2():
    n, nq, F = 100_000, 10_000_000, 1
    xa, ya = synth.uniform_nodes(n)
    za = synth.values(n, F)
    w, tb, tu = build(xa, ya, za)
    dx = torch.empty(nq, dtype=torch.float64, device=DEV); dy = torch.empty_like(dx)
    dz = torch.empty((F, nq), dtype=torch.float64, device=DEV); dl = torch.empty(nq, dtype=torch.int32, device=DEV)
    w.generate_uniform_device(0, synth.SEED_QUERIES, 0, nq, 0.0, 1.0, 0.0, 1.0, dx.data_ptr(), dy.data_ptr())
    w.reset_counters()
    ms = timed(lambda s: w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), nq, dz.data_ptr(), nq, dl.data_ptr(), s))
    sp = split_ms(w, 8)
    bad = verify_arrays(w, dx, dy, nq, F, dz, dl)
    sub = slice(0, 200_000)
    ob = oracle_subsample(w, xa, ya, za, dx[sub].cpu().numpy(), dy[sub].cpu().numpy(),
                          dl[sub].cpu().numpy(), dz[:, sub].cpu().numpy())
    # host-buffer path
    xq, yq = synth.uniform_queries(nq)
    t0 = time.perf_counter(); w.eval_e(xq, yq); t_host = time.perf_counter() - t0
    inf = w.info()
    out = {"config": "configs[1]: 1e5 uniform nodes, 1e7 uniform queries, 1 field", "queries_per_s": nq / (ms * 1e-3),
           "ms": ms, "achieved_GBps": nq * 28 / (ms * 1e-3) / 1e9, "frac_of_hbm": nq * 28 / (ms * 1e-3) / 1e9 / PEAK,
           "host_buffer_pageable_queries_per_s": nq / t_host, "host_build_s": tb, "upload_s": tu,
           "fast_vs_exact_mismatches": bad, "oracle_port_mismatches_on_2e5": ob,
           "deferred_total": inf["n_deferred"], **sp, "params": os.environ.get("GSI_PARAMS", "")}
    w.free()
    return out


def config3():
    n, nq, F = 100_000, 10_000_000, 1
    xa, ya = synth.clustered_nodes(n)
    za = synth.values(n, F)
    w, tb, tu = build(xa, ya, za)
    tree = w.get_tree(w.n_simplexes(), n)
    lv = tree["pts3"].reshape(-1, 3)[tree["type"] == 0]
    lv = lv[(lv >= 0).all(axis=1)]
    rows = tree["shuffle"][lv].astype(np.int64)
    xq, yq = synth.mixed_queries(nq, xa, ya, rows)
    dx = torch.from_numpy(xq).to(DEV); dy = torch.from_numpy(yq).to(DEV)
    dz = torch.empty((F, nq), dtype=torch.float64, device=DEV); dl = torch.empty(nq, dtype=torch.int32, device=DEV)
    w.reset_counters()
    ms = timed(lambda s: w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), nq, dz.data_ptr(), nq, dl.data_ptr(), s), steps=3, warmup=2)
    inf = w.info()
    sp = split_ms(w, 5)
    w.set_param("stats", 1); w.reset_counters()
    w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), nq, dz.data_ptr(), nq, dl.data_ptr())
    i2 = w.info(); w.set_param("stats", 0)
    sp["walk_steps_per_accepted_query"] = i2["n_steps"] / max(1, nq - i2["n_deferred"])
    bad = verify_arrays(w, dx, dy, nq, F, dz, dl)
    sub = slice(0, 200_000)
    ob = oracle_subsample(w, xa, ya, za, xq[sub], yq[sub], dl[sub].cpu().numpy(), dz[:, sub].cpu().numpy())
    leaf = dl.cpu().numpy()
    cage = float(((tree["pts3"].reshape(-1, 3)[leaf] < 0).any(axis=1)).mean())
    out = {"config": "configs[2]: 1e5 clustered/anisotropic nodes, 1e7 mixed queries (60% in bbox, 15% inflated bbox, "
                     "5% far outside, 10% on nodes, 10% on edge midpoints), 1 field",
           "queries_per_s": nq / (ms * 1e-3), "ms": ms, "achieved_GBps": nq * 28 / (ms * 1e-3) / 1e9,
           "frac_of_hbm": nq * 28 / (ms * 1e-3) / 1e9 / PEAK, "deferred_fraction": inf["n_deferred"] / (5 * nq),
           "fraction_in_cage_leaves": cage, "fast_vs_exact_mismatches": bad, "oracle_port_mismatches_on_2e5": ob,
           "host_build_s": tb, "upload_s": tu, **sp, "params": os.environ.get("GSI_PARAMS", "")}
    w.free()
    return out


def config4():
    n, F, nx, ny = 1_000_000, 16, 8192, 8192
    xa, ya = synth.uniform_nodes(n)
    za = synth.values(n, F)
    w, tb, tu = build(xa, ya, za)
    x0, y0 = float(xa.min()), float(ya.min())
    ddx, ddy = (float(xa.max()) - x0) / nx, (float(ya.max()) - y0) / ny
    npx = nx * ny
    dz = torch.empty((F, npx), dtype=torch.float64, device=DEV); dl = torch.empty(npx, dtype=torch.int32, device=DEV)
    w.reset_counters()
    ms = timed(lambda s: w.eval_raster_device_e(0, x0, ddx, nx, y0, ddy, ny, dz.data_ptr(), dl.data_ptr(), s), steps=3, warmup=2)
    inf = w.info()
    sp = split_ms(w, 5)
    vz = torch.empty((F, npx), dtype=torch.float64, device=DEV); vl = torch.empty(npx, dtype=torch.int32, device=DEV)
    w.set_param("mode", 1)
    w.eval_raster_device_e(0, x0, ddx, nx, y0, ddy, ny, vz.data_ptr(), vl.data_ptr())
    w.set_param("mode", 0)
    bad = w.compare_device(0, dl.data_ptr(), vl.data_ptr(), dz.data_ptr(), vz.data_ptr(), npx, F * npx)
    del vz, vl
    # CPU oracle on a strided sub-raster (every 64th pixel of every 64th column)
    ii, jj = np.arange(0, nx, 64), np.arange(0, ny, 64)
    gx, gy = x0 + ddx * ii, y0 + ddy * jj
    pidx = (ii[:, None] * ny + jj[None, :]).ravel()
    tsub = torch.from_numpy(pidx).to(DEV)
    ob = oracle_subsample(w, xa, ya, za, np.repeat(gx, jj.size), np.tile(gy, ii.size),
                          dl[tsub].cpu().numpy(), dz[:, tsub].cpu().numpy())
    alg = npx * (4 + 8 * F)
    out = {"config": "configs[3]: 1e6 uniform nodes, 8192x8192 raster, 16 fields (132 B/pixel algorithmic)",
           "pixels_per_s": npx / (ms * 1e-3), "ms": ms, "achieved_GBps": alg / (ms * 1e-3) / 1e9,
           "frac_of_hbm": alg / (ms * 1e-3) / 1e9 / PEAK, "deferred_total": inf["n_deferred"],
           "fast_vs_exact_mismatches": bad, "oracle_port_mismatches_on_subraster": ob, "host_build_s": tb,
           "upload_s": tu, **sp, "params": os.environ.get("GSI_PARAMS", "")}
    w.free()
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["2", "3", "4"]
    for c in which:
        r = {"2": config2, "3": config3, "4": config4}[c]()
        print(json.dumps(r), flush=True)
