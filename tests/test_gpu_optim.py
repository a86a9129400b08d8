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
            # a parameter without gradient is skipped entirely -- no decay, no moment update, its own step count --
            # exactly as torch.optim.Adam skips it
            if it == 2 and q.numel() == 1:
                p.grad = q.grad = None
            else:
                p.grad, q.grad = g.clone(), g.clone()
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


def test_training_state_resume_is_bit_exact(tmp_path):
    """save_training_state / load_training_state: optimizer moments and step counts, scheduler, epoch and the RNG streams
    -- what the reference's weight-only checkpoints (train.py:329-351) lack.  Three more steps after a resume must equal
    the uninterrupted run bit for bit, with noisy preprocessing (the noise comes from the restored generator)."""
    from cosmology_gnn_simulation_b200 import synthetic
    from cosmology_gnn_simulation_b200.data_utils import preprocess
    from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode
    from cosmology_gnn_simulation_b200.loss import combined_loss
    from cosmology_gnn_simulation_b200.optim import ExponentialLR, FusedAdam, load_training_state, save_training_state
    d = torch.device("cuda", 0)
    box = synthetic.make_box(700, "uniform", seed=8)
    md = box["metadata"]

    def make():
        torch.manual_seed(5)
        model = EncodeProcessDecode(64, 64, 2, 2, 3, message="edge", precision="fp32").to(d)
        model(sample())                                            # materialise the lazy layers
        opt = FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-4)
        return model, opt, ExponentialLR(opt, 0.7)

    def sample():
        return preprocess(box["Coordinates"][:5], box["InternalEnergy"][:5], md, box["Coordinates"][5:6].clone(),
                          box["InternalEnergy"][5:6].clone(), noise_std=3e-4, num_neighbors=8, dt=md["dt"], box_size=md["box_size"], device=d)

    def run(model, opt, sched, steps):
        out = []
        for _ in range(steps):
            opt.zero_grad()
            ls = combined_loss(model(sample()), sample(), md["dt"], 1.0, 1.0, 0.1)
            ls["loss"].backward()
            opt.step()
            sched.step()
            out.append(float(ls["loss"].detach()))
        return out

    model, opt, sched = make()
    run(model, opt, sched, 3)
    save_training_state(str(tmp_path / "state.pt"), model, opt, sched, epoch=3)
    straight = run(model, opt, sched, 3)
    final = {k: v.clone() for k, v in model.state_dict().items()}

    model2, opt2, sched2 = make()
    info = load_training_state(str(tmp_path / "state.pt"), model2, opt2, sched2)
    assert info["epoch"] == 3 and opt2.step_count == 3 and abs(opt2.lr - 1e-3 * 0.7 ** 3) < 1e-12
    resumed = run(model2, opt2, sched2, 3)
    assert resumed == straight
    assert all(torch.equal(final[k], v) for k, v in model2.state_dict().items())
    # the state also loads into torch.optim.Adam (same layout)
    ref = torch.optim.Adam(model2.parameters(), lr=1.0)
    ref.load_state_dict(opt2.state_dict())
    assert ref.state_dict()["param_groups"][0]["lr"] == opt2.lr


def test_overlapped_step_equals_plain_step_on_one_rank():
    from cosmology_gnn_simulation_b200.optim import FusedAdam
    d = torch.device("cuda", 0)
    torch.manual_seed(2)
    shapes = [(128, 17), (128,), (3, 128), (3,), (1,)]
    a = [torch.nn.Parameter(torch.randn(s, device=d)) for s in shapes]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    oa, ob = FusedAdam(a, lr=1e-2), FusedAdam(b, lr=1e-2)
    for it in range(3):
        for p, q in zip(a, b):
            g = torch.randn_like(p)
            p.grad, q.grad = g.clone(), g.clone()
        oa.step()
        ob.step_overlapped(n_buckets=3)
    assert all(torch.equal(p, q) for p, q in zip(a, b))


def _overlap_worker(rank, world, port, out):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from cosmology_gnn_simulation_b200 import distributed as cd
    from cosmology_gnn_simulation_b200.optim import FusedAdam
    cd.init_from_env("nccl")
    d = torch.device("cuda", rank)
    torch.manual_seed(3)
    shapes = [(128, 384), (128,), (128, 128), (3, 128), (3,)]
    params = [torch.nn.Parameter(torch.randn(s, device=d)) for s in shapes]
    opt = FusedAdam(params, lr=1e-2)
    gen = torch.Generator(device=d).manual_seed(100 + rank)              # every rank has its own partial gradient
    for _ in range(3):
        for p in params:
            p.grad = torch.randn(p.shape, device=d, generator=gen)
        opt.step_overlapped(n_buckets=3, average=False)
    torch.save([p.detach().cpu() for p in params], f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_overlapped_step_on_two_ranks_equals_summed_gradient_step(tmp_path):
    """FusedAdam.step_overlapped (bucketed all-reduce, the update of bucket i under the reduction of bucket i+1) against a
    single-process Adam on the SUM of the two ranks' gradients."""
    import socket
    import torch.multiprocessing as mp
    from cosmology_gnn_simulation_b200.optim import FusedAdam
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "p")
    mp.spawn(_overlap_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = torch.load(f"{out}.0"), torch.load(f"{out}.1")
    assert all(torch.equal(a, b) for a, b in zip(r0, r1))
    d = torch.device("cuda", 0)
    torch.manual_seed(3)
    shapes = [(128, 384), (128,), (128, 128), (3, 128), (3,)]
    params = [torch.nn.Parameter(torch.randn(s, device=d)) for s in shapes]
    opt = FusedAdam(params, lr=1e-2)
    gens = [torch.Generator(device=d).manual_seed(100 + r) for r in range(2)]
    for _ in range(3):
        for p in params:
            p.grad = torch.randn(p.shape, device=d, generator=gens[0]) + torch.randn(p.shape, device=d, generator=gens[1])
        opt.step()
    for a, p in zip(r0, params):
        assert torch.allclose(a, p.detach().cpu(), rtol=1e-6, atol=1e-7)
