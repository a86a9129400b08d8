#!/usr/bin/env python
"""Soak test of the edge forward whose residual comes from the input ring (C_RIN): repeats the same launch and reports every
output element that differs from the first result (rows, tile, columns) -- a refilled ring slot shows up as another tile's e.
  python tools/rin_diag.py [bf16x3|bf16] [launches]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from cosmology_gnn_simulation_b200 import ops
from cosmology_gnn_simulation_b200.ops import MlpParams
L = 128
d = torch.device('cuda', 0)
g = torch.Generator(device=d).manual_seed(0)
prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
n, k = 5000, 16
ws = [torch.randn(L, i, device=d, generator=g) / i ** 0.5 for i in (3 * L, L, L)]
bs = [torch.randn(L, device=d, generator=g) * 0.1 for _ in range(3)]
p = MlpParams(ws, bs, torch.ones(L, device=d), torch.zeros(L, device=d))
h = torch.randn(n, L, device=d, generator=g)
e = torch.randn(n * k, L, device=d, generator=g)
senders = torch.randint(0, n, (n * k,), device=d, generator=g, dtype=torch.int32)
ref = torch.empty_like(e); agg = torch.empty_like(h)
ops.mp_edge_fwd(p, h, e, senders, k, ref, agg, prec)
torch.cuda.synchronize()
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    out = torch.empty_like(e)
    ops.mp_edge_fwd(p, h, e, senders, k, out, agg, prec)
    torch.cuda.synchronize()
    bad = (out != ref).nonzero()
    if bad.numel():
        rows = bad[:, 0].unique()
        cols = bad[:, 1].unique()
        print(it, 'bad elements', bad.shape[0], 'rows', rows[:8].tolist(), '... n rows', rows.numel(), 'row%128 range', int((rows % 128).min()), int((rows % 128).max()),
              'tiles', (rows // 128).unique()[:6].tolist(), 'cols', int(cols.min()), int(cols.max()), 'maxdiff', float((out - ref).abs().max()),
              'diff==e?', float(((out - ref).abs()[bad[:, 0], bad[:, 1]]).mean()))
print('done')
