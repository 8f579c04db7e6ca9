#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scripts/placement_probe.py 10 1e9 > gpurun_out/r2i_probe_vmm.log 2>&1; cat gpurun_out/r2i_probe_vmm.log | tail -11
timeout 400 python scripts/placement_probe.py 6 1e9 vmm=0 > gpurun_out/r2i_probe_novmm.log 2>&1; tail -7 gpurun_out/r2i_probe_novmm.log
for i in 1 2 3 4 5; do timeout 200 python scripts/placement_probe.py 1 1e9 | tail -1; done > gpurun_out/r2i_probe_procs.log 2>&1; cat gpurun_out/r2i_probe_procs.log
# config 3 with the four-lane replay vs one thread
timeout 300 python - <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np, torch
import gsl_scattered_interpolation_b200 as g
from gsl_scattered_interpolation_b200 import synth
n, nq = 100_000, 10_000_000
xa, ya = synth.clustered_nodes(n)
za = synth.values(n, 1)
for coop in (1, 0):
    w = g.ScatteredCuda(n, 1); w.set_param("profile", 1); w.set_param("exact_coop", coop)
    w.init(xa, ya, za, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
    tree = w.get_tree(w.n_simplexes(), n)
    lv = tree["pts3"].reshape(-1, 3)[tree["type"] == 0]; lv = lv[(lv >= 0).all(axis=1)]
    xq, yq = synth.mixed_queries(nq, xa, ya, tree["shuffle"][lv].astype(np.int64))
    dev = torch.device("cuda:0")
    dx, dy = torch.from_numpy(xq).to(dev), torch.from_numpy(yq).to(dev)
    dz = torch.empty((1, nq), dtype=torch.float64, device=dev); dl = torch.empty(nq, dtype=torch.int32, device=dev)
    for _ in range(2): w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), nq, dz.data_ptr(), nq, dl.data_ptr())
    w.reset_counters()
    for _ in range(5): w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), nq, dz.data_ptr(), nq, dl.data_ptr())
    inf = w.info()
    vz = torch.empty((1, nq), dtype=torch.float64, device=dev); vl = torch.empty(nq, dtype=torch.int32, device=dev)
    w.set_param("mode", 1); w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), nq, vz.data_ptr(), nq, vl.data_ptr()); w.set_param("mode", 0)
    bad = w.compare_device(0, dl.data_ptr(), vl.data_ptr(), dz.data_ptr(), vz.data_ptr(), nq, nq)
    print("config 3 exact_coop=%d: locate %.3f ms exact %.3f ms per 1e7, deferred %.0f, mismatches vs one-thread replay of all: %s" % (coop, 1e3*inf["locate_seconds"]/5, 1e3*inf["exact_seconds"]/5, inf["n_deferred"]/5, bad))
    w.free()
PY
