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
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks.mem,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power, mem = [], [], set(), [], []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3])); mem.append(float(c[4]))
            except ValueError:
                continue
            for name, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "power_w_max": float(max(power)) if power else None,
                "mem_mhz": float(np.median(mem)) if mem else None,
                "samples": len(sm), "reasons": sorted(reasons)}


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
    return {"parity": parity, "value": per_step / t, "unit": "queries/s", "cores": cores, "kind": kind,
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
        "config": {"workload": f"{args.nodes} uniform nodes, {args.fields} field(s), uniform random queries "
                               f"(BASELINE.json configs[4]); CPU arm evaluates a bounded sample of "
                               f"{cb['per_step']} queries per step",
                   "nodes": args.nodes, "fields": args.fields},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "reference_build_seconds": cb["build_seconds"],
        "e2e": {"value": cb["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


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
    for kv in args.param:
        k, v = kv.split("=")
        w.set_param(k, float(v))
    build_info = {}
    # --- triangulation: built once on the host by rank 0 (timed separately), uploaded, and
    #     replicated to the other GPUs as one slab over NCCL/NVLink
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

    def step():
        w.eval_device_e(local, dx.data_ptr(), dy.data_ptr(), Q, dz.data_ptr(), Q, dl.data_ptr(), sp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    w.reset_counters()
    sampler = ClockSampler(local)
    sampler.start()                      # every rank samples its own GPU
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_mine = e0.elapsed_time(e1)
    ms_total = dist_util.max_over_ranks(ms_mine, dev)
    ms_by_rank = dist_util.gather_over_ranks(ms_mine / args.steps, dev)
    clocks = sampler.stop()
    sm_by_rank = dist_util.gather_over_ranks(clocks["sm_mhz"] or 0.0, dev)
    mem_by_rank = dist_util.gather_over_ranks(clocks.get("mem_mhz") or 0.0, dev)
    pw_by_rank = dist_util.gather_over_ranks(clocks.get("power_w_max") or 0.0, dev)
    clocks["by_rank"] = {"sm_mhz": sm_by_rank, "mem_mhz": mem_by_rank, "power_w_max": pw_by_rank}
    inf = w.info()
    value = world * Q * args.steps / (ms_total * 1e-3)
    locate_s = inf["locate_seconds"] / max(1, args.steps)          # dominant kernel, per launch
    exact_s = inf["exact_seconds"] / max(1, args.steps)
    launches = int(inf["n_kernel_launches"])

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
        step()
        torch.cuda.synchronize()
        i2 = w.info()
        stats = {"walk_steps_per_query": i2["n_steps"] / max(1, Q - i2["n_deferred"]),
                 "deferred_fraction": i2["n_deferred"] / Q}
        w.set_param("stats", 0)

    # --- e2e: the same metric through the host-buffer C-ABI call (pinned host memory; H2D of
    #     the inputs and D2H of leaf ids + values inside the timed region)
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
    w.free()

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_leg(args, steps=3, warmup=1, sample_seconds=args.cpu_sample_seconds, check=sample)

    if rank == 0:
        peak, peak_src = peaks()
        reordered = inf.get("n_reordered", 0) > 0
        kernel_name = "k_bin + k_tile_locate + k_unpermute (reorder pipeline)" if reordered else "k_walk"
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture (per query,
        # scaled to this launch); only valid for the configuration the capture was taken on
        traffic, traffic_src, ncu_counters = None, None, None
        try:
            with open(os.path.join(ROOT, "profiles", "r1c_k_walk_traffic.json")) as f:
                tj = json.load(f)
            if not reordered and n == 1_000_000 and F == 1:
                traffic = tj["dram_bytes_per_query"] * Q
                ncu_counters = {k: tj[k] for k in ("dram_throughput_pct_of_peak", "lts_sector_hit_rate_pct",
                                                   "branch_targets_threads_uniform_pct", "sm_warps_active_pct",
                                                   "l1_sectors_per_query") if k in tj}
                traffic_src = (f"{tj['dram_bytes_per_query']:.1f} B/query x {Q} queries; ncu dram__bytes_read+write.sum "
                               f"of one {tj['queries_per_launch']}-query launch ({tj['source']})")
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
            "config": {"workload": f"BASELINE.json configs[4]: {n} uniform-random nodes, {F} field(s), "
                                   f"{Q} uniform-random queries per GPU per step, triangulation replicated "
                                   f"per GPU, queries sharded (no collective on the data path)",
                       "nodes": n, "fields": F, "queries_per_gpu": Q,
                       "l2": "inputs (16 B/query) and outputs (12 B/query) are far larger than the 126 MB L2",
                       **build_info},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "traffic_source": traffic_src, "ncu": ncu_counters,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                         "kernel_ms_per_launch": 1e3 * locate_s, "exact_replay_ms_per_step": 1e3 * exact_s},
            "cpu_baseline": ({k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")} if cb else None),
            "verify": {"checksum_leaf": chk[0], "checksum_val": chk[1],
                       "fast_vs_exact_mismatches": {"queries": nv, "leaf": bad[0], "value_bits": bad[1]},
                       "cpu_reference_mismatches_on_strided_subsample": (cb or {}).get("parity"),
                       "deferred_queries_per_step": inf["n_deferred"] / max(1, args.steps),
                       "outside_cage": inf["n_outside"]},
        }
        if cb:
            line["cpu_reference_detail"] = {"single_core_qps": cb["single_core_qps"],
                                            "reference_build_seconds": cb["build_seconds"]}
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

