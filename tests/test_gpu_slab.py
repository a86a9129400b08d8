"""GPU: slab-sharded box (slab.py, preprocess_slab) against the single-GPU run of the same box.

world = 1 runs everywhere (the x-sort permutation and the range k-NN / edge-feature entry points);
world = 2 needs two GPUs (NCCL) and is skipped otherwise -- run it with `gpurun --gpus 2`."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(n, k, L, M, seed=0):
    from cosmology_gnn_simulation_b200 import synthetic
    from oracle import model_ref
    box = synthetic.make_box(n, "uniform", seed=seed)
    params = model_ref.init_params(L, L, 2, M, 3, seed=seed)
    return box, params


def _run_full(box, params, k, L, M, message, precision, dev):
    from cosmology_gnn_simulation_b200.data_utils import preprocess
    from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode
    from cosmology_gnn_simulation_b200.loss import combined_loss
    md = box["metadata"]
    g = preprocess(box["Coordinates"][:5], box["InternalEnergy"][:5], md, box["Coordinates"][5:6].clone(),
                   box["InternalEnergy"][5:6].clone(), num_neighbors=k, dt=md["dt"], box_size=md["box_size"], device=dev)
    model = EncodeProcessDecode(L, L, 2, M, 3, message=message, precision=precision)
    model.load_state_dict(params)
    model = model.to(dev)
    pred = model(g)
    ls = combined_loss(pred, g, md["dt"], 1.0, 1.0, 0.1)
    ls["loss"].backward()
    grads = {k_: (None if p.grad is None else p.grad.detach().cpu()) for k_, p in model.named_parameters()}
    return pred["acceleration"].detach().cpu(), pred["temp_rate"].detach().cpu(), float(ls["loss"]), grads


def _run_slab(box, params, k, L, M, message, precision, dev, rank, world):
    from cosmology_gnn_simulation_b200 import distributed as cd
    from cosmology_gnn_simulation_b200.data_utils import preprocess_slab
    from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode
    from cosmology_gnn_simulation_b200.slab import slab_loss
    md = box["metadata"]
    g = preprocess_slab(box["Coordinates"][:5], box["InternalEnergy"][:5], md, box["Coordinates"][5:6].clone(),
                        box["InternalEnergy"][5:6].clone(), num_neighbors=k, dt=md["dt"], box_size=md["box_size"],
                        rank=rank, world=world, device=dev)
    model = EncodeProcessDecode(L, L, 2, M, 3, message=message, precision=precision)
    model.load_state_dict(params)
    model = model.to(dev)
    pred = model(g)
    ls = slab_loss(pred, g, md["dt"], 1.0, 1.0, 0.1)
    ls["loss"].backward()
    cd.GradientBucket(model.parameters()).all_reduce(average=False)
    grads = {k_: (None if p.grad is None else p.grad.detach().cpu()) for k_, p in model.named_parameters()}
    lo, hi = g.own_range
    own = g.order[lo:hi].cpu()
    return own, pred["acceleration"].detach().cpu(), pred["temp_rate"].detach().cpu(), float(ls["loss"]), grads, g.halo.n_halo


def _compare(full, slabs, n, tol, gtol):
    acc_f, temp_f, loss_f, grads_f = full
    acc = torch.empty_like(acc_f)
    temp = torch.empty_like(temp_f)
    seen = torch.zeros(n, dtype=torch.bool)
    for own, a, t, loss, grads, _ in slabs:
        acc[own], temp[own] = a, t
        seen[own] = True
        assert abs(loss - loss_f) < 10 * tol * abs(loss_f)
    assert bool(seen.all())
    assert rel_l2(acc, acc_f) < tol and rel_l2(temp, temp_f) < tol
    grads = slabs[0][4]
    for name, gf in grads_f.items():
        if gf is None:
            assert grads[name] is None, name
        else:
            assert rel_l2(grads[name], gf) < gtol, (name, rel_l2(grads[name], gf))


@pytest.mark.parametrize("message", ["edge", "sender"])
@pytest.mark.parametrize("reorder", ["0", "1"])
def test_slab_world1_equals_plain_graph(message, reorder, monkeypatch):
    monkeypatch.setenv("CGNN_REORDER", reorder)
    dev = torch.device("cuda", 0)
    n, k, L, M = 3000, 16, 64, 3
    box, params = _setup(n, k, L, M)
    full = _run_full(box, params, k, L, M, message, "fp32", dev)
    one = _run_slab(box, params, k, L, M, message, "fp32", dev, 0, 1)
    assert one[5] == 0
    _compare(full, [one], n, 2e-5, 1e-3)


def _worker(rank, world, port, out, message, precision):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from cosmology_gnn_simulation_b200 import distributed as cd
    cd.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    n, k, L, M = 6000, 16, 128, 3
    box, params = _setup(n, k, L, M)
    res = _run_slab(box, params, k, L, M, message, precision, dev, rank, world)
    torch.save(res, f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("message,precision,reorder", [("edge", "fp32", "0"), ("sender", "fp32", "0"), ("edge", "bf16x3", "0"),
                                                       ("edge", "fp32", "1"), ("edge", "bf16x3", "1")])
def test_slab_world2_equals_single_gpu(tmp_path, message, precision, reorder, monkeypatch):
    # reorder = "1": every rank renumbers its owned rows along a Z-order curve inside the model (the halo plan's send lists move with them)
    monkeypatch.setenv("CGNN_REORDER", reorder)
    world = 2
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(world, _free_port(), out, message, precision), nprocs=world, join=True)
    slabs = [torch.load(f"{out}.{r}") for r in range(world)]
    assert all(s[5] > 0 for s in slabs)                       # both ranks really have halo rows
    dev = torch.device("cuda", 0)
    n, k, L, M = 6000, 16, 128, 3
    box, params = _setup(n, k, L, M)
    full = _run_full(box, params, k, L, M, message, precision, dev)
    tol, gtol = (2e-5, 1e-3) if precision == "fp32" else (1e-3, 1e-2)
    _compare(full, slabs, n, tol, gtol)


def _rollout_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from cosmology_gnn_simulation_b200 import distributed as cd
    from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode
    from cosmology_gnn_simulation_b200.rollout import rollout_slab
    cd.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    n, k, L, M, w = 4000, 16, 128, 2, 5
    box, params = _setup(n, k, L, M, seed=7)
    md = box["metadata"]
    model = EncodeProcessDecode(L, L, 2, M, 3, precision="bf16x3")
    model.load_state_dict(params)
    model = model.to(dev)
    data = {"Coordinates": box["Coordinates"][:w], "InternalEnergy": box["InternalEnergy"][:w]}
    res = rollout_slab(model, data, md, 0.0, md["dt"], md["box_size"], window_size=w, num_neighbors=k, n_steps=4, rank=rank, world=world)
    torch.save({k_: v.cpu() for k_, v in res.items()}, f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_rollout_world2_equals_single_gpu(tmp_path):
    """Config-4 shape in miniature: a rollout whose box is re-partitioned into x-slabs every step (particles change owner
    as they move) against the single-GPU rollout of the same box (render_rollout.py:39-85)."""
    from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode
    from cosmology_gnn_simulation_b200.rollout import rollout
    world = 2
    out = str(tmp_path / "roll")
    mp.spawn(_rollout_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r0, r1 = torch.load(f"{out}.0"), torch.load(f"{out}.1")
    assert torch.equal(r0["Coordinates"], r1["Coordinates"])             # every rank holds the same complete trajectory
    n, k, L, M, w = 4000, 16, 128, 2, 5
    box, params = _setup(n, k, L, M, seed=7)
    md = box["metadata"]
    model = EncodeProcessDecode(L, L, 2, M, 3, precision="bf16x3")
    model.load_state_dict(params)
    model = model.to(torch.device("cuda", 0))
    data = {"Coordinates": box["Coordinates"][:w], "InternalEnergy": box["InternalEnergy"][:w]}
    ref = rollout(model, data, md, 0.0, md["dt"], md["box_size"], window_size=w, num_neighbors=k, n_steps=4)
    d = (ref["Coordinates"].cpu() - r0["Coordinates"]).abs()
    d = torch.minimum(d, md["box_size"] - d)
    assert float(d.max()) < 1e-4 * md["box_size"]
    assert rel_l2(r0["InternalEnergy"], ref["InternalEnergy"].cpu()) < 1e-4


def test_halo_row_copies_match_torch_indexing():
    """cgnn_halo_pack / cgnn_halo_unpack_add: the row copies either side of the halo transport (slab.HaloPlan)."""
    from cosmology_gnn_simulation_b200 import ops
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(0)
    for rows, L, m in ((1000, 128, 333), (50, 64, 50), (7, 4, 1), (10, 128, 0)):
        src = torch.randn(rows, L, device=dev, generator=gen)
        idx = torch.randperm(rows, device=dev, generator=gen)[:m].contiguous()
        assert torch.equal(ops.halo_pack(src, idx), src[idx])
        dst = torch.randn(rows, L, device=dev, generator=gen)
        want = dst.clone().index_add_(0, idx, src[:m])
        ops.halo_unpack_add(src[:m].contiguous(), idx, dst)
        assert torch.equal(dst, want)


def test_own_share_of_a_pinned_host_sequence_reaches_the_device():
    """slab.copy_own_share: the PCIe half of preprocess_slab's sharded host-to-device transfer (one contiguous copy per frame)."""
    from cosmology_gnn_simulation_b200 import slab
    dev = torch.device("cuda", 0)
    gen = torch.Generator().manual_seed(5)
    for shape, pdim in (((6, 1001, 3), 1), ((1001, 1), 0), ((1, 1001, 3), 1)):
        t = torch.randn(*shape, generator=gen).pin_memory()
        for world in (2, 3):
            b = slab.slab_bounds(shape[pdim], world)
            for rank in range(world):
                out = torch.full(shape, float("nan"), device=dev)
                slab.copy_own_share(out, t, pdim, rank, world)
                torch.cuda.synchronize()
                mine = tuple([slice(None)] * pdim + [slice(b[rank], b[rank + 1])])
                got = out.cpu()
                assert torch.equal(got[mine], t[mine])
                assert int(torch.isnan(got).sum()) == t.numel() - t[mine].numel()
