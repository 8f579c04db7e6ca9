"""Which deferred queries of BASELINE.json configs[2] cost what: the exact replay timed on 2*10^6 queries of one
class at a time (0 = in the bounding box, 1 = inflated bounding box, 2 = far outside, 3 = copies of data nodes,
4 = midpoints of leaf edges), vertex memo on and off."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import gsl_scattered_interpolation_b200 as g
from gsl_scattered_interpolation_b200 import synth

n, nq = 100_000, 20_000_000
xa, ya = synth.clustered_nodes(n)
za = synth.values(n, 1)
dev = torch.device("cuda:0")
for memo in (1, 0):
    w = g.ScatteredCuda(n, 1)
    w.set_param("profile", 1)
    w.set_param("stats", 1)
    w.set_param("vertex_memo", memo)
    w.init(xa, ya, za, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
    tree = w.get_tree(w.n_simplexes(), n)
    lv = tree["pts3"].reshape(-1, 3)[tree["type"] == 0]
    lv = lv[(lv >= 0).all(axis=1)]
    xq, yq, cls = synth.mixed_queries(nq, xa, ya, tree["shuffle"][lv].astype(np.int64), with_class=True)
    for c in ((0, 1, 2, 3, 4) if memo else (3, 4)):
        sel = np.flatnonzero(cls == c)
        x, y = np.ascontiguousarray(xq[sel]), np.ascontiguousarray(yq[sel])
        m = x.size
        dx, dy = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
        dz = torch.empty((1, m), dtype=torch.float64, device=dev)
        dl = torch.empty(m, dtype=torch.int32, device=dev)
        for _ in range(2):
            w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), m, dz.data_ptr(), m, dl.data_ptr())
        w.reset_counters()
        for _ in range(3):
            w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), m, dz.data_ptr(), m, dl.data_ptr())
        torch.cuda.synchronize()
        inf = w.info()
        print("memo %d class %d: %d queries, deferred %.0f, walk steps per accepted query %.2f, locate %.3f ms (%.2e q/s), exact %.3f ms = %.2f ns per deferred query"
              % (memo, c, m, inf["n_deferred"] / 3, inf["n_steps"] / max(1.0, 3.0 * m - inf["n_deferred"]),
                 1e3 * inf["locate_seconds"] / 3, 3 * m / max(1e-12, inf["locate_seconds"]), 1e3 * inf["exact_seconds"] / 3,
                 1e9 * inf["exact_seconds"] / max(1, inf["n_deferred"])))
    w.free()
