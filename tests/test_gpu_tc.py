"""GPU parity of the tcgen05 tensor-core processor kernels (precision 'bf16x3' / 'bf16') against a float64
torch evaluation of the same phase (graph_network.py:83-101,177-183) and against the FP32 kernels."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

L = 128
TOL = {"bf16x3": 1e-4, "bf16": 3e-2}


def _dev():
    return torch.device("cuda", 0)


def _mlp_params(in_dim, gen, scale=1.0):
    from cosmology_gnn_simulation_b200.ops import MlpParams
    dims = [(L, in_dim), (L, L), (L, L)]
    ws = [((torch.rand(o, i, generator=gen) * 2 - 1) * scale / i ** 0.5) for o, i in dims]
    bs = [((torch.rand(o, generator=gen) * 2 - 1) * 0.1) for o, _ in dims]
    gamma = 1.0 + 0.1 * torch.randn(L, generator=gen)
    beta = 0.1 * torch.randn(L, generator=gen)
    dev = _dev()
    p = MlpParams([w.to(dev) for w in ws], [b.to(dev) for b in bs], gamma.to(dev), beta.to(dev))
    return p, ws, bs, gamma, beta


def _mlp_ln64(z, ws, bs, gamma, beta):
    z = z.double()
    for i, (w, b) in enumerate(zip(ws, bs)):
        z = z @ w.double().T + b.double()
        if i < len(ws) - 1:
            z = torch.relu(z)
    mu = z.mean(-1, keepdim=True)
    var = ((z - mu) ** 2).mean(-1, keepdim=True)
    return (z - mu) / torch.sqrt(var + 1e-5) * gamma.double() + beta.double()


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("n", [100, 256, 1000, 40000])
def test_tc_node_phase_forward(precision, n):
    from cosmology_gnn_simulation_b200 import ops
    gen = torch.Generator().manual_seed(n)
    p, ws, bs, gamma, beta = _mlp_params(2 * L, gen)
    h = torch.randn(n, L, generator=gen)
    agg = torch.randn(n, L, generator=gen) * 3.0
    ref = h.double() + _mlp_ln64(torch.cat([h, agg], -1), ws, bs, gamma, beta)
    out = torch.full((n, L), float("nan"), device=_dev())
    ops.mp_node_fwd(p, h.to(_dev()), agg.to(_dev()), out, precision)
    torch.cuda.synchronize()
    assert rel_l2(out.cpu(), ref) < TOL[precision]
    out32 = torch.empty((n, L), device=_dev())
    ops.mp_node_fwd(p, h.to(_dev()), agg.to(_dev()), out32, "fp32")
    assert rel_l2(out32.cpu(), ref) < 1e-5
    # deterministic
    out2 = torch.empty((n, L), device=_dev())
    ops.mp_node_fwd(p, h.to(_dev()), agg.to(_dev()), out2, precision)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("n,k", [(300, 16), (64, 8), (1000, 32), (5000, 16), (130, 4), (20000, 16)])
@pytest.mark.parametrize("with_agg", [True, False])
def test_tc_edge_phase_forward(precision, n, k, with_agg):
    from cosmology_gnn_simulation_b200 import ops
    gen = torch.Generator().manual_seed(n * 100 + k)
    p, ws, bs, gamma, beta = _mlp_params(3 * L, gen)
    h = torch.randn(n, L, generator=gen)
    e = torch.randn(n * k, L, generator=gen)
    senders = torch.randint(0, n, (n * k,), generator=gen, dtype=torch.int32)
    recv = torch.arange(n).repeat_interleave(k)
    u = _mlp_ln64(torch.cat([h[senders.long()], h[recv], e], -1), ws, bs, gamma, beta)     # graph_network.py:89-90
    ref_e = e.double() + u                                                                   # :182
    ref_agg = u.view(n, k, L).sum(1)                                                         # :92 (message = edge)
    d = _dev()
    e_out = torch.full((n * k, L), float("nan"), device=d)
    agg = torch.full((n, L), float("nan"), device=d) if with_agg else None
    ops.mp_edge_fwd(p, h.to(d), e.to(d), senders.to(d), k, e_out, agg, precision)
    torch.cuda.synchronize()
    assert rel_l2(e_out.cpu(), ref_e) < TOL[precision]
    if with_agg:
        assert rel_l2(agg.cpu(), ref_agg) < TOL[precision]
    e2 = torch.empty_like(e_out)
    agg2 = torch.empty_like(agg) if with_agg else None
    ops.mp_edge_fwd(p, h.to(d), e.to(d), senders.to(d), k, e2, agg2, precision)
    assert torch.equal(e_out, e2)
    if with_agg:
        assert torch.equal(agg, agg2)


def test_tc_edge_phase_in_place():
    """e_out may alias e_in (inference updates the edge stream in place)."""
    from cosmology_gnn_simulation_b200 import ops
    n, k = 2000, 16
    gen = torch.Generator().manual_seed(7)
    p, *_ = _mlp_params(3 * L, gen)
    d = _dev()
    h = torch.randn(n, L, generator=gen).to(d)
    e = torch.randn(n * k, L, generator=gen).to(d)
    senders = torch.randint(0, n, (n * k,), generator=gen, dtype=torch.int32).to(d)
    out = torch.empty_like(e)
    ops.mp_edge_fwd(p, h, e, senders, k, out, None, "bf16x3")
    e_inplace = e.clone()
    ops.mp_edge_fwd(p, h, e_inplace, senders, k, e_inplace, None, "bf16x3")
    assert torch.equal(out, e_inplace)


def test_tc_rejects_unsupported_shapes():
    from cosmology_gnn_simulation_b200 import ops
    from cosmology_gnn_simulation_b200.ops import MlpParams
    d = _dev()
    p = MlpParams([torch.randn(64, 192, device=d), torch.randn(64, 64, device=d), torch.randn(64, 64, device=d)],
                  [torch.zeros(64, device=d)] * 3, torch.ones(64, device=d), torch.zeros(64, device=d))
    h = torch.randn(100, 64, device=d)
    e = torch.randn(800, 64, device=d)
    s = torch.zeros(800, dtype=torch.int32, device=d)
    with pytest.raises(RuntimeError, match="tensor-core edge phase supports"):
        ops.mp_edge_fwd(p, h, e, s, 8, torch.empty_like(e), None, "bf16x3")


@pytest.mark.parametrize("message", ["sender", "edge"])
def test_model_tensor_core_mode_matches_oracle(message):
    """Whole model in 'bf16x3' (tensor-core processor forward): outputs and gradients within 1e-3 of float64."""
    from test_gpu_parity import _compare_with_oracle, TOL_TC
    _compare_with_oracle(message, dict(n=1500, k=16, L=128, H=128, nh=2, M=4), "bf16x3", TOL_TC)
