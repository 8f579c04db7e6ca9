#!/usr/bin/env python
"""bench.py -- interp queries/sec of the batched scattered-interpolation eval path on B200.

One "step" = one pass of the hot path (point location + barycentric blend) over one batch of
synthetic queries.  Default workload (BASELINE.json configs[4], the configuration the metric
is quoted on): 10^6 uniform-random nodes, one scalar field, 10^9 uniform-random queries PER
GPU (weak scaling; the triangulation is replicated on every GPU, queries are sharded with no
collective on the data path).

  python bench.py [--gpus N --steps K --warmup W]          our arm (CUDA, through the C ABI)
  python bench.py --impl reference [...]                   the reference's own CPU eval loop
                                                           (oracle/_ref, forked on all host cores)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

ALG_BYTES_PER_QUERY = 28          # 16 B (x, y) in + 4 B leaf id out + 8 B value out, F = 1
L2_NOTE = "inputs (16 B/query) and outputs (12 B/query) are far larger than the 126 MB L2"


def workload_string(nodes, fields):
    """identical in both arms (ours / --impl reference): the driver compares config.workload"""
    return (f"BASELINE.json configs[4]: {nodes} uniform-random nodes, {fields} field(s), uniform-random queries "
            f"in the unit square; triangulation replicated per GPU, queries sharded (no collective on the data path)")


def source_sha16():
    """hash of the kernel / library sources: ties a committed ncu capture to the code it measured
    (the GPU box has no .git, so a commit hash cannot be checked there)"""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "gsl_scattered_interpolation_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".h", ".cpp")):
            h.update(name.encode())
            h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()[:16]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nodes", type=int, default=1_000_000)
    ap.add_argument("--queries", type=int, default=1_000_000_000, help="queries per GPU per step")
    ap.add_argument("--e2e-queries", type=int, default=100_000_000,
                    help="queries per GPU per step for the host-buffer (e2e) leg")
    ap.add_argument("--fields", type=int, default=1)
    ap.add_argument("--cells-per-leaf", type=float, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify-queries", type=int, default=1_100_000_000,
                    help="cross-check the whole step against the exact replay when it has at most this many queries")
    ap.add_argument("--cpu-sample-seconds", type=float, default=15.0)
    ap.add_argument("--stats", action="store_true", help="extra untimed pass counting walk steps")
    ap.add_argument("--param", action="append", default=[], help="library knob name=value (repeatable)")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[1..3] legs (N = 1 only)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (N > 1 only)")
    ap.add_argument("--trace-out", default=None,
                    help="diagnostic: write every rank's per-CTA trace of the last timed launch to PREFIX_rank<r>.npy")
    ap.add_argument("--pageable-queries", type=int, default=20_000_000,
                    help="queries per GPU per step for the pageable host-buffer leg")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM / memory clocks, power and throttle reasons of one GPU sampled DURING the timed region.

    The sampler is started once, before the warm-up steps, and keeps sampling; window() returns the
    summary of the samples taken between two time stamps.  (Starting an nvidia-smi process right
    before the timed region, as an earlier version did, stalls the first kernel launch of the region
    by tens of milliseconds on multi-GPU boxes while NVML enumerates the devices: the GPU idles and
    the step looks 35-60 % slower on whichever ranks lose that race -- profiles/r2n2b_*.)
    In-process NVML (pynvml) when available, else one long-lived `nvidia-smi -lms` child."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks.mem,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, period=0.025):
        self.index, self.period = index, period
        self.samples = []                       # (t, sm, sm_max, power_w, mem, [reasons])
        self.stop_flag = threading.Event()
        self.proc = None
        self.kind = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.kind = "nvml"
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
        except Exception:
            try:
                self.proc = subprocess.Popen(
                    ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                     str(int(1e3 * period)), "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                    text=True)
                self.kind = "nvidia-smi"
                self.t = threading.Thread(target=self._read_smi, daemon=True)
                self.t.start()
            except Exception:
                self.kind = None

    def _poll_nvml(self):
        nv = self.nv
        bits = [("hw_slowdown", getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4))]
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mem = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_MEM))
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1e3
                r = int(reasons_fn(self.h))
                self.samples.append((time.perf_counter(), sm, self.sm_max, pw, mem, [n for n, b in bits if r & b]))
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def _read_smi(self):
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                self.samples.append((time.perf_counter(), float(c[1]), float(c[2]), float(c[3]), float(c[4]),
                                     [n for n, v in zip(self.NAMES, c[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue

    def window(self, t0, t1):
        if self.kind is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"]}
        rows = [x for x in list(self.samples) if t0 <= x[0] <= t1]
        if not rows:                            # a region shorter than the period: the nearest sample
            rows = sorted(list(self.samples), key=lambda x: abs(x[0] - 0.5 * (t0 + t1)))[:1]
        reasons = sorted({n for x in rows for n in x[5]})
        return {"sm_mhz": float(np.median([x[1] for x in rows])) if rows else None,
                "sm_max_mhz": float(max(x[2] for x in rows)) if rows else None,
                "power_w_max": float(max(x[3] for x in rows)) if rows else None,
                "mem_mhz": float(np.median([x[4] for x in rows])) if rows else None,
                "samples": len(rows), "reasons": reasons, "source": self.kind}

    def stop(self):
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()


def bind_near_gpu(local):
    """Run this rank's host threads (and so first-touch its pinned buffers) on the CPUs NVML reports
    as closest to its GPU, when the container's cpuset contains any of them.  Only affects the
    host-buffer (e2e) leg.  Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, 64)
        near = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        have = os.sched_getaffinity(0)
        use = sorted(near & have)
        if use and len(use) < len(have):
            os.sched_setaffinity(0, use)
            return f"bound to {len(use)} of {len(have)} cpus near gpu {local}"
        return f"not bound ({len(near & have)} of {len(have)} visible cpus are near gpu {local})"
    except Exception as e:                  # pragma: no cover - depends on the box
        return f"not bound ({type(e).__name__})"


def workload_nodes(n, fields):
    from gsl_scattered_interpolation_b200 import synth
    xa, ya = synth.uniform_nodes(n)
    za = synth.values(n, fields)
    return xa, ya, za


def cpu_reference_leg(args, steps, warmup, sample_seconds, check=None):
    """The reference's own eval loop (find_leaf + interp_point, linear_simplex.c:331-402,
    678-711) from oracle/_ref, forked over all host cores, on bounded samples of the workload."""
    import refimpl
    from gsl_scattered_interpolation_b200 import synth
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    xa, ya, za = workload_nodes(args.nodes, args.fields)
    if refimpl.have_ref():
        kind = "reference"
        tree = refimpl.RefTree(xa, ya, 0, synth.SEED_SHUFFLE)
        build_s = tree.build_seconds

        def run(xq, yq, nproc):
            return tree.time_eval(xq, yq, za, nproc)
    else:                                   # oracle/_ref missing: fall back to the C port
        kind = "port"
        import gsl_scattered_interpolation_b200 as g
        w = g.ScatteredCuda(args.nodes, args.fields)
        w.build(xa, ya, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, args.nodes))
        build_s = w.info()["build_seconds"]
        port = refimpl.PortTree(w.get_tree(w.n_simplexes(), args.nodes), xa, ya)
        w.free()
        cores = 1

        def run(xq, yq, nproc):
            t0 = time.perf_counter()
            port.eval(xq, yq, za)
            return time.perf_counter() - t0
    # calibrate on one core, then size each step to ~sample_seconds/(steps) of wall time
    xq, yq = synth.uniform_queries(20000)
    dt = run(xq, yq, 1)
    rate1 = 20000 / dt
    per_step = max(20000, int(rate1 * cores * max(1.0, sample_seconds / max(1, steps))))
    times = []
    for s in range(warmup + steps):
        xq, yq = synth.uniform_queries(per_step, first=20000 + s * per_step)
        dt = run(xq, yq, cores)
        if s >= warmup:
            times.append(dt)
    t = float(np.mean(times))
    parity = None
    if check is not None:
        # the GPU's results on a strided subsample of ITS queries against the CPU arm, bit for bit
        xs, ys, gl, gz = check
        if kind == "reference":
            rl, rv, _ = tree.eval(xs, ys, za, nproc=cores)
        else:
            rl, rv, _ = port.eval(xs, ys, za)
        parity = {"queries": int(xs.size), "leaf": int(np.count_nonzero(rl != gl)),
                  "value_bits": int(np.count_nonzero(np.ascontiguousarray(rv).view(np.uint64)
                                                     != np.ascontiguousarray(gz).view(np.uint64)))}
    # the thread-safe restatement on `cores` POSIX threads sharing one tree
    # (the reference itself cannot be threaded: tree->tmp_vec1, linear_simplex.c:644)
    port_mt = None
    try:
        pt = refimpl.PortTree(tree.export(), xa, ya) if kind == "reference" else port
        nmt = max(20000, per_step // 4)
        xq, yq = synth.uniform_queries(nmt, first=7 * per_step)
        t0 = time.perf_counter()
        pt.eval_mt(xq, yq, za, cores if kind == "reference" else (os.cpu_count() or 1))
        port_mt = {"value": nmt / (time.perf_counter() - t0), "unit": "queries/s",
                   "threads": cores if kind == "reference" else (os.cpu_count() or 1), "kind": "port (oracle/simplex_oracle.c, pthreads)",
                   "sample": f"{nmt} uniform queries, one pass"}
    except Exception as e:                  # pragma: no cover
        port_mt = {"error": f"{type(e).__name__}: {e}"}
    return {"parity": parity, "port_mt": port_mt, "value": per_step / t, "unit": "queries/s", "cores": cores, "kind": kind,
            "sample": f"{per_step} uniform queries per step x {steps} steps on the {args.nodes}-node tree, "
                      f"find_leaf + {args.fields} x interp_point, forked over {cores} cores",
            "single_core_qps": rate1, "build_seconds": build_s, "ms_per_step": 1e3 * t,
            "per_step": per_step}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_leg(args, args.steps, args.warmup, sample_seconds=2.0 * args.steps)
    line = {
        "impl": "reference", "metric": "interp queries/sec", "value": cb["value"], "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_string(args.nodes, args.fields), "nodes": args.nodes, "fields": args.fields,
                   "l2": L2_NOTE},
        "sample_queries_per_step": cb["per_step"],
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "reference_build_seconds": cb["build_seconds"],
        "e2e": {"value": cb["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def timed_on_stream(torch, fn, dev, steps, warmup):
    """ms per call of fn(stream_ptr), CUDA events on the launching stream"""
    st = torch.cuda.Stream(device=dev)
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


def config_legs(torch, g, synth, dev, peak, tree_1e6, nodes_1e6):
    """BASELINE.json configs[1..3] (the parity-test configurations) at their stated sizes: a short
    device-timed leg each, verified in the same run (fast path == exact DAG replay on 100 % of the
    queries; the stated-size comparison with oracle/_ref is in tests/test_gpu_stated_sizes.py).
    Not part of the driver-timed region; rank 0 at N = 1 only."""
    out = {}

    def finish(w, steps, name, units, alg, ms, bad, extra):
        inf = w.info()
        out[name] = {"value": units / (ms * 1e-3), "ms_per_step": ms, "frac": alg / (ms * 1e-3) / 1e9 / peak,
                     "achieved_GBps": alg / (ms * 1e-3) / 1e9,
                     "kernel_ms": 1e3 * inf["locate_seconds"] / steps, "exact_replay_ms": 1e3 * inf["exact_seconds"] / steps,
                     "deferred_per_step": inf["n_deferred"] / steps, "reordered": inf["n_reordered"] > 0,
                     "fast_vs_exact_mismatches": {"leaf": bad[0], "value_bits": bad[1]}, **extra}

    def arrays_leg(name, xa, ya, za, xq_t, yq_t, nq, label):
        n = len(xa)
        w = g.ScatteredCuda(n, 1)
        w.set_param("profile", 1)
        w.init(xa, ya, za, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
        dz = torch.empty((1, nq), dtype=torch.float64, device=dev)
        dl = torch.empty(nq, dtype=torch.int32, device=dev)
        fn = lambda sp: w.eval_device_e(0 if dev.index is None else dev.index, xq_t.data_ptr(), yq_t.data_ptr(), nq,
                                        dz.data_ptr(), nq, dl.data_ptr(), sp)
        timed_on_stream(torch, fn, dev, 1, 2)
        w.reset_counters()
        steps = 5
        ms = timed_on_stream(torch, fn, dev, steps, 0)
        inf_steps = steps
        vz = torch.empty((1, nq), dtype=torch.float64, device=dev)
        vl = torch.empty(nq, dtype=torch.int32, device=dev)
        w.set_param("profile", 0)
        w.set_param("mode", 1)
        w.eval_device_e(dev.index or 0, xq_t.data_ptr(), yq_t.data_ptr(), nq, vz.data_ptr(), nq, vl.data_ptr())
        w.set_param("mode", 0)
        bad = w.compare_device(dev.index or 0, dl.data_ptr(), vl.data_ptr(), dz.data_ptr(), vz.data_ptr(), nq, nq)
        finish(w, inf_steps, name, nq, nq * 28, ms, bad, {"unit": "queries/s", "workload": label})
        w.free()

    # configs[1]: 10^5 uniform nodes, 10^7 uniform queries
    n, nq = 100_000, 10_000_000
    xa, ya = synth.uniform_nodes(n)
    xq, yq = synth.uniform_queries(nq)
    arrays_leg("configs[1]", xa, ya, synth.values(n, 1), torch.from_numpy(xq).to(dev), torch.from_numpy(yq).to(dev), nq,
               "1e5 uniform nodes, 1e7 uniform queries, 1 field")
    # configs[2]: 10^5 clustered nodes, 10^7 mixed queries (needs the leaves for the edge midpoints)
    xa, ya = synth.clustered_nodes(n)
    wt = g.ScatteredCuda(n, 1)
    wt.build(xa, ya, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
    tree = wt.get_tree(wt.n_simplexes(), n)
    wt.free()
    lv = tree["pts3"].reshape(-1, 3)[tree["type"] == 0]
    lv = lv[(lv >= 0).all(axis=1)]
    xq, yq = synth.mixed_queries(nq, xa, ya, tree["shuffle"][lv].astype(np.int64))
    arrays_leg("configs[2]", xa, ya, synth.values(n, 1), torch.from_numpy(xq).to(dev), torch.from_numpy(yq).to(dev), nq,
               "1e5 clustered/anisotropic nodes, 1e7 mixed queries (60% in bbox, 15% inflated bbox, 5% far outside, "
               "10% on nodes, 10% on edge midpoints), 1 field")
    del xq, yq
    # configs[3]: 10^6 nodes (the bench's own triangulation), 8192 x 8192 raster, 16 fields
    if tree_1e6 is not None:
        xa, ya = nodes_1e6
        n, F, nx, ny = len(xa), 16, 8192, 8192
        za = synth.values(n, F)
        w = g.ScatteredCuda(n, F)
        w.set_param("profile", 1)
        w.init_from_tree(tree_1e6["pts3"], tree_1e6["lnk3"], tree_1e6["type"], tree_1e6["shuffle"], tree_1e6["shift"],
                         tree_1e6["scale"], tree_1e6["seed_points"], np.stack([xa, ya], axis=1), za)
        x0, y0 = float(xa.min()), float(ya.min())
        ddx, ddy = (float(xa.max()) - x0) / nx, (float(ya.max()) - y0) / ny
        npx = nx * ny
        dz = torch.empty((F, npx), dtype=torch.float64, device=dev)
        dl = torch.empty(npx, dtype=torch.int32, device=dev)
        d = dev.index or 0
        fn = lambda sp: w.eval_raster_device_e(d, x0, ddx, nx, y0, ddy, ny, dz.data_ptr(), dl.data_ptr(), sp)
        timed_on_stream(torch, fn, dev, 1, 2)
        w.reset_counters()
        steps = 5
        ms = timed_on_stream(torch, fn, dev, steps, 0)
        vz = torch.empty((F, npx), dtype=torch.float64, device=dev)
        vl = torch.empty(npx, dtype=torch.int32, device=dev)
        w.set_param("profile", 0)
        w.set_param("mode", 1)
        w.eval_raster_device_e(d, x0, ddx, nx, y0, ddy, ny, vz.data_ptr(), vl.data_ptr())
        w.set_param("mode", 0)
        bad = w.compare_device(d, dl.data_ptr(), vl.data_ptr(), dz.data_ptr(), vz.data_ptr(), npx, F * npx)
        finish(w, steps, "configs[3]", npx, npx * (4 + 8 * F), ms, bad,
               {"unit": "pixels/s", "workload": "1e6 uniform nodes, 8192x8192 raster, 16 fields (132 B/pixel algorithmic)"})
        w.free()
    return out


def main_ours(args):
    import torch
    import torch.distributed as dist
    import gsl_scattered_interpolation_b200 as g
    from gsl_scattered_interpolation_b200 import dist_util, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if g.device_count() <= 0:
        raise SystemExit("bench.py needs a CUDA device: libgsl_scattered_cuda has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_near_gpu(local) if world > 1 and not os.environ.get("GSI_NO_BIND") else "not bound (single rank)"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n, F, Q = args.nodes, args.fields, args.queries
    # query / result arrays of this rank, allocated before anything else touches the device heap
    stream = torch.cuda.Stream(device=dev)
    sp = stream.cuda_stream
    dx = torch.empty(Q, dtype=torch.float64, device=dev)
    dy = torch.empty(Q, dtype=torch.float64, device=dev)
    dz = torch.empty((F, Q), dtype=torch.float64, device=dev)
    dl = torch.empty(Q, dtype=torch.int32, device=dev)
    w = g.ScatteredCuda(n, F)
    w.set_devices([local])
    if args.cells_per_leaf:
        w.set_param("cells_per_leaf", args.cells_per_leaf)
    w.set_param("profile", 1)
    if args.trace_out:
        w.set_param("walk_trace", 1)
    for kv in args.param:
        k, v = kv.split("=")
        w.set_param(k, float(v))
    build_info = {}
    # --- triangulation: built once on the host by rank 0 (timed separately), uploaded, and
    #     replicated to the other GPUs as one slab over NCCL/NVLink.  Every rank ends up with its slab
    #     in an allocation of its own made the same way (rank 0's staging copy for the broadcast is
    #     released before anything is timed).
    xa = ya = None
    if rank == 0:
        xa, ya, za = workload_nodes(n, F)
        t0 = time.perf_counter()
        w.build(xa, ya, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
        t1 = time.perf_counter()
        w.upload(za)
        t2 = time.perf_counter()
        inf = w.info()
        build_info = {"host_build_seconds": t1 - t0, "flatten_upload_seed_seconds": t2 - t1,
                      "n_simplexes": inf["n_simplexes"], "n_leaves": inf["n_leaves"],
                      "grid": inf["grid_g"], "slab_bytes": inf["slab_bytes"]}
    if world > 1:
        slab = torch.empty(w.slab_bytes() if rank == 0 else 0, dtype=torch.uint8, device=dev)
        if rank == 0:
            w.slab_export_device(local, slab.data_ptr())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        slab = dist_util.broadcast_blob(slab, 0)
        torch.cuda.synchronize()
        build_info["slab_broadcast_seconds"] = time.perf_counter() - t0
        if rank != 0:
            w.init_from_slab_device(slab.data_ptr(), slab.numel())
        del slab
        if not os.environ.get("GSI_BENCH_KEEP_CACHE"):
            torch.cuda.empty_cache()
    # the part of the slab the fast path reads: leaf records, vertex values,
    # seed grid -- not the history DAG / entry table of the exact replay
    inf0 = w.info()
    struct_bytes = int(inf0["n_leaves"]) * 96 + n * 8 * F + int(inf0["grid_g"]) ** 2 * 4

    # --- queries of this rank: k in [rank*Q, (rank+1)*Q) of the global stream, generated on
    #     the device by an untimed kernel and resident in HBM when the timed region starts
    k0, k1 = dist_util.shard_range(rank, world, Q)
    w.generate_uniform_device(local, synth.SEED_QUERIES, k0, k1 - k0, 0.0, 1.0, 0.0, 1.0,
                              dx.data_ptr(), dy.data_ptr(), sp)
    stream.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)                  # started long before anything is timed

    def timed_region(nq, steps, warmup, sample_clocks):
        """`steps` passes over the first nq queries of this rank's arrays: (max-over-ranks ms of the
        region, per-rank ms per step, clocks); every pass is one eval_device_e call"""
        def step():
            w.eval_device_e(local, dx.data_ptr(), dy.data_ptr(), nq, dz.data_ptr(), Q, dl.data_ptr(), sp)
        for _ in range(warmup):
            step()
        barrier()
        w.reset_counters()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        stream.synchronize()
        t1 = time.perf_counter()
        barrier()
        ms_mine = e0.elapsed_time(e1)
        clocks = sampler.window(t0, t1) if sample_clocks else None   # every rank samples its own GPU
        return dist_util.max_over_ranks(ms_mine, dev), dist_util.gather_over_ranks(ms_mine / steps, dev), clocks

    ms_total, ms_by_rank, clocks = timed_region(Q, args.steps, args.warmup, True)
    if args.trace_out:
        tr = w.walk_trace(local)
        np.save(f"{args.trace_out}_rank{rank}.npy", tr[tr[:, 2] > 0])
    sm_by_rank = dist_util.gather_over_ranks(clocks["sm_mhz"] or 0.0, dev)
    mem_by_rank = dist_util.gather_over_ranks(clocks.get("mem_mhz") or 0.0, dev)
    pw_by_rank = dist_util.gather_over_ranks(clocks.get("power_w_max") or 0.0, dev)
    clocks["by_rank"] = {"sm_mhz": sm_by_rank, "mem_mhz": mem_by_rank, "power_w_max": pw_by_rank}
    inf = w.info()
    value = world * Q * args.steps / (ms_total * 1e-3)
    locate_s = inf["locate_seconds"] / max(1, args.steps)          # dominant kernel(s), per launch
    exact_s = inf["exact_seconds"] / max(1, args.steps)
    launches = int(inf["n_kernel_launches"])
    reordered = inf.get("n_reordered", 0) > 0

    # --- verification inside the run: checksum, and fast path == exact replay on a prefix
    chk = w.checksum_device(local, dl.data_ptr(), dz.data_ptr(), Q, sp)
    # strided subsample of this rank's queries and results for the CPU cross-check (rank 0, N = 1)
    sample = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        idx = torch.arange(0, Q, max(1, Q // 1_000_000), device=dev)[:1_000_000]
        sample = (dx[idx].cpu().numpy(), dy[idx].cpu().numpy(), dl[idx].cpu().numpy(),
                  dz[:, idx].cpu().numpy())
        del idx
    # exact DAG replay of EVERY query of the step (all 10^9 by default) against the fast path
    nv = Q if Q <= args.verify_queries else min(Q, 20_000_000)
    vz = torch.empty((F, nv), dtype=torch.float64, device=dev)
    vl = torch.empty(nv, dtype=torch.int32, device=dev)
    w.set_param("mode", 1)
    w.eval_device_e(local, dx.data_ptr(), dy.data_ptr(), nv, vz.data_ptr(), nv, vl.data_ptr(), sp)
    w.set_param("mode", 0)
    bad = w.compare_device(local, dl.data_ptr(), vl.data_ptr(), dz.data_ptr(), vz.data_ptr(), nv, nv, sp)
    del vz, vl

    stats = None
    if args.stats:
        w.set_param("stats", 1)
        w.reset_counters()
        w.eval_device_e(local, dx.data_ptr(), dy.data_ptr(), Q, dz.data_ptr(), Q, dl.data_ptr(), sp)
        torch.cuda.synchronize()
        i2 = w.info()
        stats = {"walk_steps_per_query": i2["n_steps"] / max(1, Q - i2["n_deferred"]),
                 "deferred_fraction": i2["n_deferred"] / Q}
        w.set_param("stats", 0)

    # --- strong scaling (BASELINE.json configs[4] as written: a FIXED stream of Q queries sharded
    #     over the N GPUs): rank r re-generates k in [r*Q/N, (r+1)*Q/N) into the front of its arrays
    strong = None
    if world > 1 and not args.no_strong:
        s0, s1 = dist_util.split_range(rank, world, Q)
        w.generate_uniform_device(local, synth.SEED_QUERIES, s0, s1 - s0, 0.0, 1.0, 0.0, 1.0,
                                  dx.data_ptr(), dy.data_ptr(), sp)
        stream.synchronize()
        ms_s, ms_s_by_rank, _ = timed_region(s1 - s0, args.steps, max(1, args.warmup), False)
        strong = {"scaling": "strong", "queries_total_per_step": Q, "value": Q * args.steps / (ms_s * 1e-3),
                  "unit": "queries/s", "ms_per_step": ms_s / args.steps, "ms_per_step_by_rank": ms_s_by_rank}

    sampler.stop()

    # --- e2e: the same metric through the host-buffer C-ABI call (H2D of the inputs and D2H of leaf
    #     ids + values inside the timed region), from pinned host memory and from pageable (malloc'd)
    #     arrays such as the reference's gsl_matrix / gsl_vector
    del dx, dy, dz, dl
    torch.cuda.empty_cache()
    Qe = min(args.e2e_queries, Q)
    L = g._lib.load()
    hx, hy = L.gsl_scattered_cuda_host_alloc(Qe * 8), L.gsl_scattered_cuda_host_alloc(Qe * 8)
    hz, hl = L.gsl_scattered_cuda_host_alloc(Qe * 8 * F), L.gsl_scattered_cuda_host_alloc(Qe * 4)
    e2e = None
    if hx and hy and hz and hl:
        L.gsl_scattered_cuda_generate_uniform_host(synth.SEED_QUERIES, rank * Q, Qe, 0.0, 1.0, 0.0, 1.0, hx, hy)
        w.eval_ptr_e(hx, hy, Qe, hz, hl)                       # warm-up (allocates staging)
        e2e_steps = max(2, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            w.eval_ptr_e(hx, hy, Qe, hz, hl)
        torch.cuda.synchronize()
        dt = dist_util.max_over_ranks(time.perf_counter() - t0, dev)
        e2e = {"value": world * Qe * e2e_steps / dt, "unit": "queries/s",
               "h2d_bytes_per_step": Qe * 16, "d2h_bytes_per_step": Qe * (4 + 8 * F),
               "queries_per_gpu_per_step": Qe, "host_memory": "pinned (gsl_scattered_cuda_host_alloc)",
               "cpu_affinity": affinity}
    for p in (hx, hy, hz, hl):
        if p:
            L.gsl_scattered_cuda_host_free(p)
    e2e_pageable = None
    Qp = min(args.pageable_queries, Q)
    if Qp > 0:
        px, py = g.generate_uniform_host(synth.SEED_QUERIES, rank * Q, Qp, 0.0, 1.0, 0.0, 1.0)   # numpy = malloc'd
        pz = np.empty((F, Qp), dtype=np.float64)
        pl = np.empty(Qp, dtype=np.int32)
        # several ranks share the host cores: keep the copy threads of all ranks within them
        copy_threads = max(2, min(16, (3 * (os.cpu_count() or 16)) // (4 * world)))
        w.set_param("copy_threads", copy_threads)
        w.eval_e(px, py, out_z=pz, out_leaf=pl)                 # warm-up (allocates the pinned staging)
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            w.eval_e(px, py, out_z=pz, out_leaf=pl)
        dt = dist_util.max_over_ranks(time.perf_counter() - t0, dev)
        e2e_pageable = {"value": world * Qp * 2 / dt, "unit": "queries/s", "queries_per_gpu_per_step": Qp,
                        "host_memory": "pageable (numpy / malloc), staged through the library's pinned buffers by "
                                       f"{copy_threads} host threads", "h2d_bytes_per_step": Qp * 16, "d2h_bytes_per_step": Qp * (4 + 8 * F)}
        del px, py, pz, pl
    tree_1e6 = None
    if rank == 0 and world == 1 and not args.no_configs and n == 1_000_000:
        tree_1e6 = w.get_tree(w.n_simplexes(), n)
    w.free()

    configs = None
    if rank == 0 and world == 1 and not args.no_configs:
        peak0, _ = peaks()
        try:
            configs = config_legs(torch, g, synth, dev, peak0, tree_1e6, (xa, ya))
        except Exception as e:                                   # a failed side leg must not lose the main line
            configs = {"error": f"{type(e).__name__}: {e}"}
    del tree_1e6

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_leg(args, steps=3, warmup=1, sample_seconds=args.cpu_sample_seconds, check=sample)

    if rank == 0:
        peak, peak_src = peaks()
        variant = {0: "k_walk", 1: "k_walkp", 2: "k_walkc", 3: "k_walkf"}
        for kv in args.param:
            if kv.startswith("walk_variant="):
                walk_variant = int(float(kv.split("=")[1]))
                break
        else:
            walk_variant = 2
        for kv in args.param:
            if kv.startswith("leaf_stride=") and int(float(kv.split("=")[1])) != 128 and walk_variant == 2:
                walk_variant = 0
        kernel_name = "k_bin + k_tile_walk + k_unpermute (reorder pipeline)" if reordered else variant[walk_variant]
        # DRAM traffic of the dominant kernel(s) from the committed ncu --set full capture (per query,
        # scaled to this launch).  Only emitted when the capture was taken on exactly these sources and
        # on this configuration (scripts/ncu_capture.sh writes the source hash into the file).
        traffic, traffic_src, ncu_counters = None, None, None
        try:
            with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
                tj = json.load(f)
            key = "pipeline" if reordered else "k_walk"
            if tj.get("source_sha16") == source_sha16() and n == tj.get("nodes") and F == tj.get("fields") and key in tj:
                traffic = tj[key]["dram_bytes_per_query"] * Q
                ncu_counters = tj[key].get("counters")
                traffic_src = (f"{tj[key]['dram_bytes_per_query']:.1f} B/query x {Q} queries; ncu dram__bytes_read+write.sum "
                               f"of one {tj['queries_per_launch']}-query launch, sources {tj['source_sha16']} "
                               f"({tj.get('source', 'profiles/')})")
            else:
                traffic_src = "no ncu capture of these sources / this configuration (profiles/r2_traffic.json)"
        except Exception:
            pass
        alg = Q * (16 + 4 + 8 * F) + struct_bytes
        achieved = alg / locate_s / 1e9 if locate_s > 0 else None
        line = {
            "metric": "interp queries/sec", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "ms_per_step_by_rank": ms_by_rank,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_string(n, F), "nodes": n, "fields": F, "l2": L2_NOTE},
            "setup": {"queries_per_gpu_per_step": Q, "source_sha16": source_sha16(), **build_info},
            "clocks": clocks,
            "e2e": e2e,
            "e2e_pageable": e2e_pageable,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "traffic_source": traffic_src,
                         # the same launch on its REAL DRAM bytes (ncu) instead of the algorithmic ones: how close
                         # the kernel is to the HBM wall on the traffic it actually causes
                         "traffic_GBps": (traffic / locate_s / 1e9) if (traffic and locate_s > 0) else None,
                         "traffic_frac_of_peak": (traffic / locate_s / 1e9 / peak) if (traffic and locate_s > 0) else None,
                         "ncu": ncu_counters,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                         "kernel_ms_per_launch": 1e3 * locate_s, "exact_replay_ms_per_step": 1e3 * exact_s},
            "cpu_baseline": ({k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")} if cb else None),
            "verify": {"checksum_leaf": chk[0], "checksum_val": chk[1],
                       "fast_vs_exact_mismatches": {"queries": nv, "leaf": bad[0], "value_bits": bad[1]},
                       "cpu_reference_mismatches_on_strided_subsample": (cb or {}).get("parity"),
                       "deferred_queries_per_step": inf["n_deferred"] / max(1, args.steps),
                       "outside_cage": inf["n_outside"]},
        }
        if strong:
            line["scaling_strong"] = strong
        if configs is not None:
            line["configs"] = configs
        if cb:
            line["cpu_reference_detail"] = {"single_core_qps": cb["single_core_qps"],
                                            "reference_build_seconds": cb["build_seconds"],
                                            "thread_safe_port_multi_thread": cb["port_mt"]}
        if stats:
            line["walk_stats"] = stats
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


class OnlyJsonOnStdout:
    """stdout must carry exactly one JSON line; native libraries (NCCL's version banner under
    torchrun) write to file descriptor 1 behind Python's back.  Everything written to fd 1 while
    this is active goes to stderr; emit() writes to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        os.write(self.real, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        os.close(self.real)
        return False


OUT = None


def emit(line):
    text = json.dumps(line)
    if OUT is not None:
        OUT.emit(text)
    else:
        print(text, flush=True)


if __name__ == "__main__":
    a = parse()
    with OnlyJsonOnStdout() as OUT:
        if a.impl == "reference":
            main_reference(a)
        else:
            main_ours(a)
    sys.exit(0)

