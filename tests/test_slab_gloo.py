"""CPU, world_size 2 and 3, gloo: the slab-sharding index logic and its halo exchange (slab.py).

What is checked without a GPU: ownership ranges, the halo plan built from global sender ids (local ids, halo
grouping by owner), that `exchange` delivers exactly the owners' rows, that `reduce_grad` equals a dense
scatter-add over the global graph, and the sharded loss against the single-process loss of the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _global_problem(n, k, L, seed=0):
    g = torch.Generator().manual_seed(seed)
    senders = torch.randint(0, n, (n * k,), generator=g)            # global ids, edge e = receiver * k + rank
    # make the graph spatially local-ish like a k-NN graph: most senders close to the receiver id
    recv = torch.arange(n).repeat_interleave(k)
    near = (recv + torch.randint(-40, 41, (n * k,), generator=g)) % n
    senders = torch.where(torch.rand(n * k, generator=g) < 0.9, near, senders)
    h = torch.randn(n, L, generator=g)
    gh = torch.randn(n * k, L, generator=g)                         # a per-edge gradient to scatter by sender
    return senders, h, gh


def _worker(rank, world, port, out, n, k, L):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from cosmology_gnn_simulation_b200 import slab
    dist.init_process_group("gloo")
    senders, h, gh = _global_problem(n, k, L)
    bounds = slab.slab_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    plan, local, halo_gid = slab.plan_from_global_senders(senders[lo * k: hi * k], bounds, rank, world)
    assert plan.n_own == hi - lo and plan.n_halo == halo_gid.numel()
    # local ids point at the right global particles
    gid_of_local = torch.cat([torch.arange(lo, hi), halo_gid])
    assert torch.equal(gid_of_local[local.long()], senders[lo * k: hi * k])
    # halo grouped by owner, ascending
    owner = torch.searchsorted(torch.tensor(bounds), halo_gid, right=True) - 1
    assert bool((owner[1:] >= owner[:-1]).all()) and bool((halo_gid[1:] > halo_gid[:-1]).all())
    assert int((owner == rank).sum()) == 0
    # exchange: halo rows become the owners' rows
    h_loc = torch.full((plan.n_loc, L), float("nan"))
    h_loc[:plan.n_own] = h[lo:hi]
    plan.exchange(h_loc)
    assert torch.equal(h_loc, h[gid_of_local])
    # reduce_grad: local scatter-add by sender, then the halo rows go home
    dh_loc = torch.zeros(plan.n_loc, L).index_add_(0, local.long(), gh[lo * k: hi * k])
    plan.reduce_grad(dh_loc)
    assert float(dh_loc[plan.n_own:].abs().max()) == 0.0 if plan.n_halo else True
    # sharded loss
    g = torch.Generator().manual_seed(5)
    acc, temp = torch.randn(n, 3, generator=g), torch.randn(n, 1, generator=g)
    ya, yt = torch.randn(n, 3, generator=g), torch.randn(n, 1, generator=g)
    a = acc[lo:hi].clone().requires_grad_(True)
    t = temp[lo:hi].clone().requires_grad_(True)

    class G:
        pass
    graph = G()
    graph.halo, graph.y_acc, graph.y_temp_rate = plan, ya[lo:hi], yt[lo:hi]
    ls = slab.slab_loss({"acceleration": a, "temp_rate": t}, graph, 0.01, 1.0, 0.5, 0.1)
    ls["loss"].backward()
    torch.save({"dh_own": dh_loc[:plan.n_own].clone(), "loss": ls["loss"].detach(), "da": a.grad, "dt": t.grad,
                "parts": [ls["acc_loss"], ls["temp_rate_loss"], ls["momentum_loss"]]}, f"{out}.{rank}")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_plan_exchange_and_loss(tmp_path, world):
    from oracle import model_ref
    n, k, L = 600, 8, 16
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(world, _free_port(), out, n, k, L), nprocs=world, join=True)
    senders, h, gh = _global_problem(n, k, L)
    dense = torch.zeros(n, L).index_add_(0, senders, gh)
    res = [torch.load(f"{out}.{r}") for r in range(world)]
    got = torch.cat([r["dh_own"] for r in res])
    assert torch.allclose(got, dense, rtol=1e-5, atol=1e-5)
    # loss: identical on every rank, equal to the oracle's single-process loss and gradient
    g = torch.Generator().manual_seed(5)
    acc, temp = torch.randn(n, 3, generator=g).requires_grad_(True), torch.randn(n, 1, generator=g).requires_grad_(True)
    ya, yt = torch.randn(n, 3, generator=g), torch.randn(n, 1, generator=g)
    ref = model_ref.loss(acc, temp, ya, yt, 0.01, w_acc=1.0, w_temp=0.5, w_mom=0.1)
    ref["loss"].backward()
    for r in res:
        assert abs(float(r["loss"]) - float(ref["loss"])) < 1e-5 * abs(float(ref["loss"]))
    assert torch.allclose(torch.cat([r["da"] for r in res]), acc.grad, rtol=1e-4, atol=1e-7)
    assert torch.allclose(torch.cat([r["dt"] for r in res]), temp.grad, rtol=1e-4, atol=1e-7)


def test_slab_bounds_cover_everything():
    from cosmology_gnn_simulation_b200 import slab
    for n, w in ((10, 3), (7, 8), (1000, 8), (5, 1)):
        b = slab.slab_bounds(n, w)
        assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(w))
        assert max(b[i + 1] - b[i] for i in range(w)) - min(b[i + 1] - b[i] for i in range(w)) <= 1


def _share_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from cosmology_gnn_simulation_b200 import slab
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gen = torch.Generator().manual_seed(3)
    full = [torch.randn(6, 11, 3, generator=gen), torch.randn(11, 1, generator=gen), torch.randn(1, 11, 3, generator=gen)]
    res = []
    for t, pdim in zip(full, (1, 0, 1)):
        b = slab.slab_bounds(t.shape[pdim], world)
        mine = tuple([slice(None)] * pdim + [slice(b[rank], b[rank + 1])])
        part = torch.full_like(t, float("nan"))               # a rank starts with its own share only
        slab.copy_own_share(part, t, pdim, rank, world)
        res.append(bool(torch.equal(part[mine], t[mine])) and int(torch.isnan(part).sum()) == t.numel() - t[mine].numel())
        slab.share_over_ranks(part, pdim, world)
        res.append(bool(torch.equal(part, t)))
    # host tensor already where it should be / single rank: passes through untouched
    res.append(slab.sharded_to_device(full[0], 1, rank, world, "cpu") is full[0])
    torch.save(res, f"{out}.{rank}")
    dist.destroy_process_group()


def test_shares_of_a_host_tensor_reach_every_rank_world2(tmp_path):
    """slab.share_over_ranks: the NVLink half of the sharded host-to-device transfer of preprocess_slab (uneven shares, one
    block per frame)."""
    out = str(tmp_path / "share")
    mp.spawn(_share_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for r in range(2):
        assert all(torch.load(f"{out}.{r}"))
