#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`, no GPU needed) into a small markdown/JSON pair.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_name [--alg-bytes B | --alg-flops F]

Writes <out>.md (table of the metrics the roofline argument uses, per profiled launch) and <out>.json.
"""
from __future__ import annotations

import csv
import io
import json
import subprocess
import sys

KEYS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def to_bytes(val: str, unit: str) -> float:
    v = float(val.replace(",", ""))
    u = unit.lower()
    for name, mul in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0)):
        if u.startswith(name):
            return v * mul
    return v


def to_us(val: str, unit: str) -> float:
    v = float(val.replace(",", ""))
    u = unit.lower()
    return {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(u.replace("second", "s").replace("usecond", "us"), v) \
        if u in ("ns", "us", "ms", "s") else (v / 1e3 if u.startswith("n") else v * 1e3 if u.startswith("m") else v)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    alg_bytes = alg_flops = None
    if "--alg-bytes" in sys.argv:
        alg_bytes = float(sys.argv[sys.argv.index("--alg-bytes") + 1])
    if "--alg-flops" in sys.argv:
        alg_flops = float(sys.argv[sys.argv.index("--alg-flops") + 1])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = {"value": r[i], "unit": units[i]}
        if "gpu__time_duration.sum" in d:
            us = to_us(d["gpu__time_duration.sum"]["value"], d["gpu__time_duration.sum"]["unit"])
            tr = to_bytes(d["dram__bytes_read.sum"]["value"], d["dram__bytes_read.sum"]["unit"]) + \
                to_bytes(d["dram__bytes_write.sum"]["value"], d["dram__bytes_write.sum"]["unit"])
            d["_derived"] = {"duration_us": us, "dram_traffic_bytes": tr, "dram_GBps_under_ncu": tr / us / 1e3}
            if alg_bytes:
                d["_derived"]["algorithmic_bytes"] = alg_bytes
                d["_derived"]["traffic_over_algorithmic"] = tr / alg_bytes
            if alg_flops:
                d["_derived"]["algorithmic_flops"] = alg_flops
                d["_derived"]["TFLOPs_under_ncu"] = alg_flops / us / 1e6
        launches.append(d)
    with open(out + ".json", "w") as f:
        json.dump({"source": rep, "launches": launches}, f, indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu summary of `{rep}`\n\n`ncu --set full --clock-control none`; durations under ncu are cold-cache and serialised "
                "(compare shares and traffic, not absolutes).\n\n")
        for n, d in enumerate(launches):
            f.write(f"## launch {n}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]['value']} | {d[k]['unit']} |\n")
            for k, v in d.get("_derived", {}).items():
                f.write(f"| derived: {k} | {v:.6g} | |\n")
            f.write("\n")
    print(json.dumps(launches[0].get("_derived", {})) if launches else "no launches")


if __name__ == "__main__":
    main()
