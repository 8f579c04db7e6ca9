"""ncu driver: BASELINE.json configs[2] (10^5 clustered nodes, 10^7 mixed queries) -- the exact replay of the
deferred 20 % (on nodes / on edge midpoints) is what its step time is made of.
    python scripts/profile_cfg3.py [passes] [name=value ...]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import gsl_scattered_interpolation_b200 as g
from gsl_scattered_interpolation_b200 import synth

passes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n, nq = 100_000, 10_000_000
xa, ya = synth.clustered_nodes(n)
za = synth.values(n, 1)
w = g.ScatteredCuda(n, 1)
w.set_param("profile", 1)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    w.set_param(k, float(v))
w.init(xa, ya, za, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
tree = w.get_tree(w.n_simplexes(), n)
lv = tree["pts3"].reshape(-1, 3)[tree["type"] == 0]
lv = lv[(lv >= 0).all(axis=1)]
xq, yq = synth.mixed_queries(nq, xa, ya, tree["shuffle"][lv].astype(np.int64))
dev = torch.device("cuda:0")
dx, dy = torch.from_numpy(xq).to(dev), torch.from_numpy(yq).to(dev)
dz = torch.empty((1, nq), dtype=torch.float64, device=dev)
dl = torch.empty(nq, dtype=torch.int32, device=dev)
w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), nq, dz.data_ptr(), nq, dl.data_ptr())
w.reset_counters()
for _ in range(passes):
    w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), nq, dz.data_ptr(), nq, dl.data_ptr())
torch.cuda.synchronize()
inf = w.info()
vz = torch.empty((1, nq), dtype=torch.float64, device=dev)
vl = torch.empty(nq, dtype=torch.int32, device=dev)
w.set_param("mode", 1)
w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), nq, vz.data_ptr(), nq, vl.data_ptr())
bad = w.compare_device(0, dl.data_ptr(), vl.data_ptr(), dz.data_ptr(), vz.data_ptr(), nq, nq)
print("config 3: locate %.3f ms exact %.3f ms per 1e7 queries (%.3e q/s), deferred %.0f, mismatches vs one-thread replay of all: %s"
      % (1e3 * inf["locate_seconds"] / passes, 1e3 * inf["exact_seconds"] / passes,
         nq * passes / (inf["locate_seconds"] + inf["exact_seconds"]), inf["n_deferred"] / passes, bad))
w.free()
