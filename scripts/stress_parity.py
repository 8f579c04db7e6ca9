"""Stress parity at the target size with adversarial data: N clustered/anisotropic nodes (config-3
generator), Q mixed queries (on nodes, on edge midpoints, outside the hull, far outside).  All
evaluation paths must agree bit for bit on every query: direct fast path, reorder pipeline, exact
replay from entry nodes, exact replay from the root; plus the CPU oracle port on a subsample.
usage: stress_parity.py [nodes] [queries]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import gsl_scattered_interpolation_b200 as g
import refimpl
from gsl_scattered_interpolation_b200 import synth

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 20_000_000
xa, ya = synth.clustered_nodes(n)
za = synth.values(n, 1)
w = g.ScatteredCuda(n, 1)
t0 = time.perf_counter()
w.init(xa, ya, za, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
t_init = time.perf_counter() - t0
tree = w.get_tree(w.n_simplexes(), n)
lv = tree["pts3"].reshape(-1, 3)[tree["type"] == 0]
lv = lv[(lv >= 0).all(axis=1)]
rows = tree["shuffle"][lv].astype(np.int64)
xq, yq = synth.mixed_queries(nq, xa, ya, rows)
dev = torch.device("cuda:0")
dx, dy = torch.from_numpy(xq).to(dev), torch.from_numpy(yq).to(dev)
res = {}
for name, params in (("direct", {"mode": 0, "reorder": 0, "exact_entry": 1}),
                     ("reorder", {"mode": 0, "reorder": 1, "exact_entry": 1}),
                     ("exact_entry", {"mode": 1, "reorder": 0, "exact_entry": 1}),
                     ("exact_root", {"mode": 1, "reorder": 0, "exact_entry": 0})):
    for k, v in params.items():
        w.set_param(k, v)
    z = torch.empty(nq, dtype=torch.float64, device=dev); l = torch.empty(nq, dtype=torch.int32, device=dev)
    w.reset_counters()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), nq, z.data_ptr(), nq, l.data_ptr())
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res[name] = (z, l, dt, w.info()["n_deferred"], st)
out = {"nodes": n, "queries": nq, "init_s": t_init, "n_simplexes": int(w.n_simplexes())}
for name in ("reorder", "exact_entry", "exact_root"):
    bad = w.compare_device(0, res["direct"][1].data_ptr(), res[name][1].data_ptr(), res["direct"][0].data_ptr(),
                           res[name][0].data_ptr(), nq, nq)
    out[f"direct_vs_{name}"] = list(bad)
for name, r in res.items():
    out[f"{name}_s"] = r[2]; out[f"{name}_deferred"] = int(r[3])
sub = np.arange(0, nq, max(1, nq // 200_000))
port = refimpl.PortTree(tree, xa, ya)
pl, pv, _ = port.eval(xq[sub], yq[sub], za)
gl = res["direct"][1].cpu().numpy()[sub]; gz = res["direct"][0].cpu().numpy()[sub]
out["port_mismatches_on_subsample"] = [int((gl != pl).sum()), int((gz.view(np.uint64) != pv[0].view(np.uint64)).sum()), int(sub.size)]
print(json.dumps(out))
w.free()
