"""Fused Adam (csrc/optim.cu) against torch.optim.Adam + ExponentialLR as the reference configures them
(train.py:183-187,263-265,316)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("weight_decay", [0.0, 1e-2])
def test_fused_adam_matches_torch_adam(weight_decay):
    from cosmology_gnn_simulation_b200.optim import ExponentialLR, FusedAdam
    d = torch.device("cuda", 0)
    torch.manual_seed(0)
    shapes = [(128, 384), (128,), (128, 128), (3, 128), (1,), (5, 7)]                 # 35 + 1 elements: exercises the scalar tail
    ref_params = [torch.nn.Parameter(torch.randn(s, device=d)) for s in shapes]
    my_params = [torch.nn.Parameter(p.detach().clone()) for p in ref_params]
    ref = torch.optim.Adam(ref_params, lr=1e-3, weight_decay=weight_decay)
    ref_sched = torch.optim.lr_scheduler.ExponentialLR(ref, gamma=0.9)
    mine = FusedAdam(my_params, lr=1e-3, weight_decay=weight_decay)
    sched = ExponentialLR(mine, gamma=0.9)
    for it in range(6):
        grads = [torch.randn(s, device=d) * (10.0 ** (it - 3)) for s in shapes]
        for p, q, g in zip(ref_params, my_params, grads):
            p.grad = g.clone()
            q.grad = g.clone() if not (it == 2 and q.numel() == 1) else None      # a parameter without gradient: treated as zero
            if it == 2 and p.numel() == 1:
                p.grad = torch.zeros_like(p)
        ref.step()
        mine.step()
        if it % 2 == 1:
            ref_sched.step()
            sched.step()
        assert abs(ref_sched.get_last_lr()[0] - sched.get_last_lr()[0]) < 1e-12
        for p, q in zip(ref_params, my_params):
            assert torch.allclose(p, q, rtol=2e-6, atol=1e-7), (it, p.shape, (p - q).abs().max())


def test_fused_adam_keeps_module_semantics():
    """Parameters become views of one flat buffer: the module still trains, state_dict round-trips."""
    from cosmology_gnn_simulation_b200.optim import FusedAdam
    d = torch.device("cuda", 0)
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 2)).to(d)
    before = {k: v.clone() for k, v in net.state_dict().items()}
    opt = FusedAdam(net.parameters(), lr=1e-2)
    assert all(torch.equal(before[k], v) for k, v in net.state_dict().items())
    x = torch.randn(32, 8, device=d)
    for _ in range(3):
        opt.zero_grad()
        net(x).square().mean().backward()
        opt.step()
    after = net.state_dict()
    assert any(not torch.equal(before[k], after[k]) for k in before)
    net.load_state_dict(before)
    assert all(torch.equal(before[k], v) for k, v in net.state_dict().items())
    assert net[0].weight.data_ptr() == opt.flat.data_ptr()                            # still views of the flat buffer
