#!/usr/bin/env python
"""Runs the processor edge phase (cgnn_mp_edge_fwd) alone on random inputs of a BASELINE config size: the
command profiled by ncu for the dominant kernel.  Prints the CUDA-event time per call."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from cosmology_gnn_simulation_b200 import ops  # noqa: E402
from cosmology_gnn_simulation_b200.ops import MlpParams  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=32768)
ap.add_argument("--k", type=int, default=16)
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--no-agg", action="store_true")
ap.add_argument("--phase", default="edge", choices=["edge", "node", "rows"])
a = ap.parse_args()
L = 128
d = torch.device("cuda", 0)
g = torch.Generator(device=d).manual_seed(0)
in_dim = 3 * L if a.phase == "edge" else (2 * L if a.phase == "node" else L)
ws = [torch.randn(L, i, device=d, generator=g) / i ** 0.5 for i in (in_dim, L, L)]
bs = [torch.randn(L, device=d, generator=g) * 0.1 for _ in range(3)]
p = MlpParams(ws, bs, torch.ones(L, device=d), torch.zeros(L, device=d))
h = torch.randn(a.n, L, device=d, generator=g)
e = torch.randn(a.n * a.k, L, device=d, generator=g)
senders = torch.randint(0, a.n, (a.n * a.k,), device=d, generator=g, dtype=torch.int32)
e_out = torch.empty_like(e)
agg = None if a.no_agg else torch.empty_like(h)
h_out = torch.empty_like(h)
times = []
for _ in range(a.reps):
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    if a.phase == "edge":
        ops.mp_edge_fwd(p, h, e, senders, a.k, e_out, agg, a.precision)
    elif a.phase == "node":
        ops.mp_node_fwd(p, h, e[:a.n], h_out, a.precision)
    else:
        ops.mlp_rows_fwd(p, e, a.precision)          # 3-layer chain over the E rows, no gather / residual / sum
    t.record()
    torch.cuda.synchronize()
    times.append(s.elapsed_time(t))
flops = 2.0 * a.n * a.k * 5 * L * L if a.phase == "edge" else (2.0 * a.n * 4 * L * L if a.phase == "node" else 2.0 * a.n * a.k * 3 * L * L)
best = min(times)
print(f"{a.phase} fwd n={a.n} k={a.k} {a.precision}: ms per call {['%.3f' % t for t in times]}  best {best:.3f} ms = "
      f"{flops / best / 1e9:.1f} algorithmic TFLOP/s")
