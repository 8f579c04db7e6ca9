"""Small driver for ncu: n nodes, Q device-generated queries, a few eval_device_e passes.
usage: profile_locate.py [nodes] [queries] [fields] [passes] [name=value ...]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gsl_scattered_interpolation_b200 as g
from gsl_scattered_interpolation_b200 import synth

pos = [a for a in sys.argv[1:] if "=" not in a]
kv = [a.split("=") for a in sys.argv[1:] if "=" in a]
n = int(float(pos[0])) if len(pos) > 0 else 1_000_000
Q = int(float(pos[1])) if len(pos) > 1 else 50_000_000
F = int(pos[2]) if len(pos) > 2 else 1
passes = int(pos[3]) if len(pos) > 3 else 3
xa, ya = synth.uniform_nodes(n)
za = synth.values(n, F)
w = g.ScatteredCuda(n, F)
for k, v in kv:
    w.set_param(k, float(v))
w.set_param("profile", 1)
w.init(xa, ya, za, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
dev = torch.device("cuda:0")
dx = torch.empty(Q, dtype=torch.float64, device=dev); dy = torch.empty_like(dx)
dz = torch.empty((F, Q), dtype=torch.float64, device=dev); dl = torch.empty(Q, dtype=torch.int32, device=dev)
w.generate_uniform_device(0, synth.SEED_QUERIES, 0, Q, 0.0, 1.0, 0.0, 1.0, dx.data_ptr(), dy.data_ptr())
for _ in range(passes):
    w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), Q, dz.data_ptr(), Q, dl.data_ptr())
torch.cuda.synchronize()
inf = w.info()
print("ok", inf["n_deferred"], inf["n_reordered"], "locate %.3f ms exact %.3f ms per pass of %d queries" % (
    1e3 * inf["locate_seconds"] / passes, 1e3 * inf["exact_seconds"] / passes, Q))
w.free()
