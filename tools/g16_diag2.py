#!/usr/bin/env python
"""Stale-read detector: the shared MLP workspace is filled with NaN bit patterns before every call of the edge backward; a result that
depends on anything the call did not write itself shows up as NaN (or differs between calls)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from cosmology_gnn_simulation_b200 import ops  # noqa: E402
from cosmology_gnn_simulation_b200.ops import MlpParams  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3g"
with_de_next = len(sys.argv) > 2 and sys.argv[2] == "de_next"
poison = len(sys.argv) > 3 and sys.argv[3] == "poison"
n, k = 20000, 16
L, d = 128, torch.device("cuda", 0)
g = torch.Generator(device=d).manual_seed(0)
ws_ = [torch.randn(L, i, device=d, generator=g) / i ** 0.5 for i in (3 * L, L, L)]
bs = [torch.randn(L, device=d, generator=g) * 0.1 for _ in range(3)]
p = MlpParams(ws_, bs, torch.ones(L, device=d), torch.zeros(L, device=d))
h = torch.randn(n, L, device=d, generator=g)
e = torch.randn(n * k, L, device=d, generator=g)
senders = torch.randint(0, n, (n * k,), device=d, generator=g, dtype=torch.int32)
rowptr, perm = ops.csr_transpose(senders, n)
dagg = torch.randn(n, L, device=d, generator=g)
dh0 = torch.randn(n, L, device=d, generator=g)
de0 = torch.randn(n * k, L, device=d, generator=g)
names = ["de", "dh", "W1", "b1", "W2", "b2", "W3", "b3", "gamma", "beta"]
ref = None
for rep in range(4):
    if poison:
        w = ops._mp_bwd_ws(p.c_struct(), n, k, prec, d, n_nodes=n)
        w.view(torch.int32).fill_(0x7FC00000 if rep % 2 == 0 else 0x7F817F81 - (1 << 32) if False else 0x7FC07FC0)
    de, dh = de0.clone(), dh0.clone()
    grads = ops.mp_edge_bwd(p, h, e, senders, rowptr, perm, k, de if with_de_next else None, dagg, de, dh, None, prec)
    torch.cuda.synchronize()
    cur = [de, dh] + list(grads)
    nans = [nm for nm, t in zip(names, cur) if not torch.isfinite(t).all()]
    msg = f"rep {rep}: non-finite in {nans}" if nans else f"rep {rep}: all finite"
    if ref is None:
        ref = [t.clone() for t in cur]
    else:
        diff = [nm for nm, a, b in zip(names, cur, ref) if not torch.equal(a, b)]
        msg += f"; differ from rep 0: {diff}"
        if "de" in diff:
            dd = (cur[0] - ref[0]).abs()
            rows = (dd.max(dim=1).values > 0).nonzero().flatten().cpu()
            msg += f"; de rows differing {rows.numel()}: first {rows[:12].tolist()} tiles256 {sorted(set((rows // 256).tolist()))[:16]} max {float(dd.max()):.3g}"
    print(msg)
print(f"{prec} de_next={with_de_next} poison={poison}")
