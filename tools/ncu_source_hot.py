#!/usr/bin/env python
"""Per-source-line aggregation of an ncu report's source page (needs -lineinfo + --import-source on):
samples, executed warp instructions and the dominant stall reasons per CUDA source line.
  python tools/ncu_source_hot.py gpurun_out/x.ncu-rep [top_n]"""
import collections
import csv
import io
import subprocess
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Line No"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
H = rows[0]
li, si = H.index("Line No"), H.index("# Samples")
ii = H.index("Instructions Executed")
src_i = H.index("Source")
def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


stall_cols = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), ""])
for r in rows[1:]:
    if len(r) <= ii:
        continue
    try:
        ln = int(r[li])
    except ValueError:
        continue
    a = agg[ln]
    a[0] += num(r[si])
    a[1] += num(r[ii])
    for c in stall_cols:
        v = num(r[c])
        if v:
            a[2][H[c][6:]] += v
tot_s = sum(a[0] for a in agg.values())
tot_i = sum(a[1] for a in agg.values())
src = {}
try:
    fpath = lines[0].split(",")[1].strip('"')
    for n, l in enumerate(open(fpath), 1):
        src[n] = l.rstrip()
except Exception:
    pass
print(f"# {path}: {tot_s} samples, {tot_i} warp instructions")
print(f"{'line':>5s} {'samp%':>6s} {'inst%':>6s}  top stalls | source")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = " ".join(f"{k}:{v}" for k, v in a[2].most_common(3))
    print(f"{ln:5d} {100 * a[0] / max(tot_s, 1):6.2f} {100 * a[1] / max(tot_i, 1):6.2f}  {st:45s} | {src.get(ln, '')[:110].strip()}")
