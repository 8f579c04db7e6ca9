"""ncu driver: config-4 shaped raster (10^6 nodes, 16 fields) at reduced size NX x NY."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gsl_scattered_interpolation_b200 as g
from gsl_scattered_interpolation_b200 import synth
n = 1_000_000; F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
ny = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
kernel = int(sys.argv[4]) if len(sys.argv) > 4 else 2
xa, ya = synth.uniform_nodes(n); za = synth.values(n, F)
w = g.ScatteredCuda(n, F); w.set_param("kernel", kernel)
w.init(xa, ya, za, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
x0, y0 = float(xa.min()), float(ya.min())
dx, dy = (float(xa.max()) - x0) / 8192, (float(ya.max()) - y0) / 8192
dz = torch.empty((F, nx * ny), dtype=torch.float64, device="cuda:0"); dl = torch.empty(nx * ny, dtype=torch.int32, device="cuda:0")
for _ in range(3):
    w.eval_raster_device_e(0, x0, dx, nx, y0, dy, ny, dz.data_ptr(), dl.data_ptr())
torch.cuda.synchronize(); print("ok"); w.free()
