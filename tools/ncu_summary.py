#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/.

  python tools/ncu_summary.py launches gpurun_out/x_launches.csv > profiles/rNN_launches.txt
  python tools/ncu_summary.py report   gpurun_out/x.ncu-rep      > profiles/rNN_kernel.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active",
        "gpu__dram_throughput", "sm__pipe_tensor_cycles_active", "sm__warps_active.avg.pct_of_peak",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "sm__throughput.avg.pct_of_peak", "sm__inst_executed_pipe_tensor", "smsp__cycles_active.avg",
        "l1tex__data_bank_conflicts", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active", "launch__shared_mem_per_block",
        "sm__inst_executed_pipe_uniform", "smsp__average_warp")


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        name = r[ki].split("(")[0][-70:]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.3f} ms total (ncu: cold cache, serialised; compare SHARES)")
    print(f"{'kernel':70s} {'n':>5s} {'total_ms':>10s} {'avg_us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:70s} {v[0]:5d} {v[1] / 1e3:10.3f} {v[1] / v[0]:10.2f} {v[1] / tot:7.3f}")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H = rows[0]
    units = rows[1]
    name_i = H.index("Kernel Name")
    for r in rows[2:]:
        print(f"== {r[name_i][:100]}  (ID {r[0]})")
        for i, h in enumerate(H):
            if any(h.startswith(k) for k in KEYS):
                print(f"   {h:75s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
