"""Diagnostic (not a test): prints the tensor-core backward errors of the node / edge phases vs float64."""
import sys
import torch
sys.path.insert(0, "tests")
from test_gpu_tc import _mlp_params, _mlp_ln64, _leaf64, _dev, L
from conftest import rel_l2
from cosmology_gnn_simulation_b200 import ops

def node(n, precision):
    gen = torch.Generator().manual_seed(n + 1)
    p, ws, bs, gamma, beta = _mlp_params(2 * L, gen)
    h = torch.randn(n, L, generator=gen); agg = torch.randn(n, L, generator=gen) * 2.0; dnext = torch.randn(n, L, generator=gen)
    w64, b64, g64, be64 = _leaf64(ws, bs, gamma, beta)
    h64, a64 = h.double().requires_grad_(True), agg.double().requires_grad_(True)
    out = h64 + _mlp_ln64(torch.cat([h64, a64], -1), w64, b64, g64, be64)
    (out * dnext.double()).sum().backward()
    d = _dev()
    dh, dagg = torch.empty((n, L), device=d), torch.empty((n, L), device=d)
    grads = ops.mp_node_bwd(p, h.to(d), agg.to(d), dnext.to(d), dh, dagg, precision)
    ref = []
    for w, b in zip(w64, b64): ref += [w.grad, b.grad]
    ref += [g64.grad, be64.grad]
    errs = [rel_l2(g.cpu(), r) for g, r in zip(grads, ref)]
    print(f"node n={n} {precision}: dh-dnext {rel_l2(dh.cpu()-dnext, h64.grad-dnext.double()):.2e} dagg {rel_l2(dagg.cpu(), a64.grad):.2e} params " + " ".join(f"{e:.1e}" for e in errs))

def edge(n, k, precision):
    gen = torch.Generator().manual_seed(n * 10 + k)
    p, ws, bs, gamma, beta = _mlp_params(3 * L, gen)
    h = torch.randn(n, L, generator=gen); e = torch.randn(n * k, L, generator=gen)
    senders = torch.randint(0, n, (n * k,), generator=gen, dtype=torch.int32)
    recv = torch.arange(n).repeat_interleave(k)
    de_next = torch.randn(n * k, L, generator=gen); dagg = torch.randn(n, L, generator=gen)
    w64, b64, g64, be64 = _leaf64(ws, bs, gamma, beta)
    h64, e64 = h.double().requires_grad_(True), e.double().requires_grad_(True)
    u = _mlp_ln64(torch.cat([h64[senders.long()], h64[recv], e64], -1), w64, b64, g64, be64)
    loss = (u.view(n, k, L).sum(1) * dagg.double()).sum() + ((e64 + u) * de_next.double()).sum()
    loss.backward()
    d = _dev(); sd = senders.to(d)
    rowptr, perm = ops.csr_transpose(sd, n)
    de = torch.empty((n * k, L), device=d); gs = torch.empty((n * k, L), device=d); dh = torch.zeros((n, L), device=d)
    grads = ops.mp_edge_bwd(p, h.to(d), e.to(d), sd, rowptr, perm, k, de_next.to(d), dagg.to(d), de, dh, gs, precision)
    ref = []
    for w, b in zip(w64, b64): ref += [w.grad, b.grad]
    ref += [g64.grad, be64.grad]
    errs = [rel_l2(g.cpu(), r) for g, r in zip(grads, ref)]
    print(f"edge n={n} k={k} {precision}: de-de_next {rel_l2(de.cpu()-de_next, e64.grad-de_next.double()):.2e} dh {rel_l2(dh.cpu(), h64.grad):.2e} params " + " ".join(f"{e:.1e}" for e in errs))

for prec in ("fp32", "bf16x3", "bf16"):
    for n in (200, 3000, 50000):
        node(n, prec)
    for n, k in ((300, 16), (4000, 8), (30000, 16)):
        edge(n, k, prec)
