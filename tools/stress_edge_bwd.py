#!/usr/bin/env python
"""Determinism stress of cgnn_mp_edge_bwd with halo rows (n_nodes > n): every kernel is deterministic, so repeated calls on the
same inputs must give bit-identical results -- any difference is a race.  python tools/stress_edge_bwd.py [--reps 30]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from cosmology_gnn_simulation_b200 import ops  # noqa: E402
from cosmology_gnn_simulation_b200.ops import MlpParams  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=3000)
ap.add_argument("--halo", type=int, default=600)
ap.add_argument("--k", type=int, default=16)
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--no-de-next", action="store_true", help="the last processor step: no gradient on the edge output")
a = ap.parse_args()
L, d = 128, torch.device("cuda", 0)
g = torch.Generator(device=d).manual_seed(0)
ws = [torch.randn(L, i, device=d, generator=g) / i ** 0.5 for i in (3 * L, L, L)]
bs = [torch.randn(L, device=d, generator=g) * 0.1 for _ in range(3)]
p = MlpParams(ws, bs, torch.ones(L, device=d), torch.zeros(L, device=d))
n, nn, k = a.n, a.n + a.halo, a.k
h = torch.randn(nn, L, device=d, generator=g)
e = torch.randn(n * k, L, device=d, generator=g)
senders = torch.randint(0, nn, (n * k,), device=d, generator=g, dtype=torch.int32)
rowptr, perm = ops.csr_transpose(senders, nn)
de_next0 = torch.randn(n * k, L, device=d, generator=g)
dagg = torch.randn(n, L, device=d, generator=g)
dh0 = torch.randn(nn, L, device=d, generator=g)
ref = None
bad = 0
for rep in range(a.reps):
    de = de_next0.clone()
    dh = dh0.clone()
    grads = ops.mp_edge_bwd(p, h, e, senders, rowptr, perm, k, None if a.no_de_next else de, dagg, de, dh, None, a.precision)
    torch.cuda.synchronize()
    cur = [de, dh] + list(grads)
    if ref is None:
        ref = [t.clone() for t in cur]
    else:
        diff = [i for i, (x, y) in enumerate(zip(cur, ref)) if not torch.equal(x, y)]
        if diff:
            bad += 1
            print(f"rep {rep}: tensors {diff} differ, max abs {[float((cur[i] - ref[i]).abs().max()) for i in diff]}")
print(f"n={n} halo={a.halo} k={k} {a.precision}{' no de_next' if a.no_de_next else ''}: {bad} of {a.reps - 1} repetitions differ from the first")
