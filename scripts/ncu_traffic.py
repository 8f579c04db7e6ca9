#!/usr/bin/env python
"""Turn `ncu --page raw --csv` dumps of the hot kernels into profiles/r2_traffic.json: DRAM bytes
per query (dram__bytes_read.sum + dram__bytes_write.sum of ONE launch / its queries) and the
counters north_star names, stamped with the hash of the sources they were measured on (bench.py
only quotes the file when the hash matches the library it runs).

    python scripts/ncu_traffic.py OUT.json QUERIES_PER_LAUNCH NODES FIELDS pipeline=PIPE.csv k_walk=WALK.csv [source=...]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import source_sha16  # noqa: E402

COUNTERS = {
    "dram_throughput_pct_of_peak": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts_sector_hit_rate_pct": "lts__t_sector_hit_rate.pct",
    "l1_sector_hit_rate_pct": "l1tex__t_sector_hit_rate.pct",
    "branch_targets_threads_uniform_pct": "smsp__sass_average_branch_targets_threads_uniform.pct",
    "sm_warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "fp64_pipe_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "lsu_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "registers_per_thread": "launch__registers_per_thread",
}


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return value * scale.get(unit, 1)


def to_ms(value, unit):
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "s": 1e3, "second": 1e3, "nsecond": 1e-6}
    return value * scale.get(unit, 1)


def read(path, queries):
    rows = list(csv.reader(open(path)))
    # skip ncu's "==PROF==" preamble lines if the dump came from --log-file
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[start], rows[start + 1]
    out = []
    for r in rows[start + 2:]:
        if len(r) != len(hdr):
            continue
        g = lambda name: (num(r[hdr.index(name)]), units[hdr.index(name)]) if name in hdr else (None, "")
        rd, rdu = g("dram__bytes_read.sum")
        wr, wru = g("dram__bytes_write.sum")
        t, tu = g("gpu__time_duration.sum")
        inst, _ = g("smsp__inst_executed.sum")
        tinst, _ = g("smsp__thread_inst_executed.sum")
        k = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].split("::")[-1].split("<")[0],
             "ms": to_ms(t, tu) if t is not None else None,
             "dram_read_bytes_per_query": to_bytes(rd, rdu) / queries if rd is not None else None,
             "dram_write_bytes_per_query": to_bytes(wr, wru) / queries if wr is not None else None,
             "warp_instructions_per_query": inst / queries if inst else None,
             "thread_instructions_per_query": tinst / queries if tinst else None,
             "counters": {k2: g(m)[0] for k2, m in COUNTERS.items() if g(m)[0] is not None}}
        k["dram_bytes_per_query"] = (k["dram_read_bytes_per_query"] or 0) + (k["dram_write_bytes_per_query"] or 0)
        out.append(k)
    return out


def main():
    out_path, queries, nodes, fields = sys.argv[1], int(float(sys.argv[2])), int(float(sys.argv[3])), int(sys.argv[4])
    kv = dict(a.split("=", 1) for a in sys.argv[5:])
    doc = {"source_sha16": source_sha16(), "queries_per_launch": queries, "nodes": nodes, "fields": fields,
           "source": kv.get("source", "scripts/ncu_capture.sh")}
    if "pipeline" in kv:
        ks = read(kv["pipeline"], queries)
        doc["pipeline"] = {"kernels": ks, "dram_bytes_per_query": sum(k["dram_bytes_per_query"] for k in ks),
                           "ms": sum(k["ms"] or 0 for k in ks),
                           "counters": {k["kernel"]: k["counters"] for k in ks}}
    if "k_walk" in kv:
        ks = read(kv["k_walk"], queries)
        doc["k_walk"] = {"kernels": ks[:1], "dram_bytes_per_query": ks[0]["dram_bytes_per_query"], "ms": ks[0]["ms"],
                         "counters": ks[0]["counters"]}
    if "k_locate" in kv:
        ks = read(kv["k_locate"], queries)
        doc["k_locate_raster"] = {"kernels": ks[:1], "counters": ks[0]["counters"], "ms": ks[0]["ms"]}
    json.dump(doc, open(out_path, "w"), indent=1)
    print(json.dumps({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk != "kernels"})
                      for k, v in doc.items()}, indent=1))


if __name__ == "__main__":
    main()
