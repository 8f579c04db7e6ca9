"""Is the bimodal k_walk time (39 vs 53-62 ms per 10^9 queries, seen per process in multi-GPU runs) a
property of where the slab lands?  One process: build once, save; then adopt the image again and
again (fresh allocation each time) and time the same evaluation.  usage: placement_probe.py [reps] [queries] [name=value...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gsl_scattered_interpolation_b200 as g
from gsl_scattered_interpolation_b200 import synth

pos = [a for a in sys.argv[1:] if "=" not in a]
kv = [a.split("=") for a in sys.argv[1:] if "=" in a]
reps = int(pos[0]) if pos else 10
Q = int(float(pos[1])) if len(pos) > 1 else 1_000_000_000
n = 1_000_000
path = "/tmp/placement_probe.gsltri"
dev = torch.device("cuda:0")
dx = torch.empty(Q, dtype=torch.float64, device=dev); dy = torch.empty_like(dx)
dz = torch.empty((1, Q), dtype=torch.float64, device=dev); dl = torch.empty(Q, dtype=torch.int32, device=dev)
if not os.path.exists(path):
    xa, ya = synth.uniform_nodes(n)
    w = g.ScatteredCuda(n, 1)
    w.init(xa, ya, synth.values(n, 1), g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
    w.save(path)
    w.free()
filler = []
for r in range(reps):
    w = g.ScatteredCuda(n, 1)
    for k, v in kv:
        w.set_param(k, float(v))
    w.set_param("profile", 1)
    w.load(path)
    if r == 0:
        w.generate_uniform_device(0, synth.SEED_QUERIES, 0, Q, 0.0, 1.0, 0.0, 1.0, dx.data_ptr(), dy.data_ptr())
    w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), Q, dz.data_ptr(), Q, dl.data_ptr())
    w.reset_counters()
    for _ in range(2):
        w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), Q, dz.data_ptr(), Q, dl.data_ptr())
    inf = w.info()
    print("rep %2d  locate %.2f ms per %g queries" % (r, 1e3 * inf["locate_seconds"] / 2, Q), flush=True)
    w.free()
    # perturb the heap so the next slab lands elsewhere
    filler.append(torch.empty((r + 1) * 37 * 1024 * 1024, dtype=torch.uint8, device=dev))
