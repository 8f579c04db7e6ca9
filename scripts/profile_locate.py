"""Small driver for ncu: 10^6 nodes, Q device-generated queries, a few eval_device_e passes."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gsl_scattered_interpolation_b200 as g
from gsl_scattered_interpolation_b200 import synth

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
Q = int(float(sys.argv[2])) if len(sys.argv) > 2 else 50_000_000
F = int(sys.argv[3]) if len(sys.argv) > 3 else 1
passes = int(sys.argv[4]) if len(sys.argv) > 4 else 3
xa, ya = synth.uniform_nodes(n)
za = synth.values(n, F)
w = g.ScatteredCuda(n, F)
w.init(xa, ya, za, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
dev = torch.device("cuda:0")
dx = torch.empty(Q, dtype=torch.float64, device=dev); dy = torch.empty_like(dx)
dz = torch.empty((F, Q), dtype=torch.float64, device=dev); dl = torch.empty(Q, dtype=torch.int32, device=dev)
w.generate_uniform_device(0, synth.SEED_QUERIES, 0, Q, 0.0, 1.0, 0.0, 1.0, dx.data_ptr(), dy.data_ptr())
for _ in range(passes):
    w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), Q, dz.data_ptr(), Q, dl.data_ptr())
torch.cuda.synchronize()
print("ok", w.info()["n_deferred"])
w.free()
