"""Raster throughput vs field count and raster size on one B200 (10^6 nodes), fast path verified
against the exact replay on the device for every pixel."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gsl_scattered_interpolation_b200 as g
from gsl_scattered_interpolation_b200 import synth

n = 1_000_000
xa, ya = synth.uniform_nodes(n)
order = g.shuffle_mt19937(synth.SEED_SHUFFLE, n)
dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6446.0
for F in (1, 4, 16):
    w = g.ScatteredCuda(n, F)
    w.set_param("profile", 1)
    w.init(xa, ya, synth.values(n, F), g.DEFAULT, order)
    x0, y0 = float(xa.min()), float(ya.min())
    for nx, ny in ((8192, 8192), (2048, 2048), (16384, 16384) if F == 1 else (1024, 1024)):
        dx, dy = (float(xa.max()) - x0) / nx, (float(ya.max()) - y0) / ny
        z = torch.empty((F, nx * ny), dtype=torch.float64, device=dev); l = torch.empty(nx * ny, dtype=torch.int32, device=dev)
        st = torch.cuda.Stream(device=dev)
        for _ in range(3):
            w.eval_raster_device_e(0, x0, dx, nx, y0, dy, ny, z.data_ptr(), l.data_ptr(), st.cuda_stream)
        st.synchronize(); w.reset_counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5):
            w.eval_raster_device_e(0, x0, dx, nx, y0, dy, ny, z.data_ptr(), l.data_ptr(), st.cuda_stream)
        e1.record(st); st.synchronize()
        ms = e0.elapsed_time(e1) / 5
        inf = w.info()
        w.set_param("mode", 1)
        z2 = torch.empty_like(z); l2 = torch.empty_like(l)
        w.eval_raster_device_e(0, x0, dx, nx, y0, dy, ny, z2.data_ptr(), l2.data_ptr())
        w.set_param("mode", 0)
        bad = w.compare_device(0, l.data_ptr(), l2.data_ptr(), z.data_ptr(), z2.data_ptr(), nx * ny, F * nx * ny)
        gbs = nx * ny * (4 + 8 * F) / (ms * 1e-3) / 1e9
        print(json.dumps({"fields": F, "raster": [nx, ny], "ms": ms, "pixels_per_s": nx * ny / (ms * 1e-3), "achieved_GBps": gbs,
                          "frac_of_hbm": gbs / peak, "locate_ms": 1e3 * inf["locate_seconds"] / 5, "exact_ms": 1e3 * inf["exact_seconds"] / 5,
                          "mismatches_vs_exact": list(bad)}), flush=True)
        del z, l, z2, l2
    w.free()
