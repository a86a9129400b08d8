"""CPU, world_size 2, gloo: the gradient all-reduce used by the multi-GPU bench path."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from cosmology_gnn_simulation_b200 import distributed as cd
    r, w, _ = cd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    dead = torch.nn.Parameter(torch.zeros(5))                 # a parameter that never gets a gradient
    params = list(model.parameters()) + [dead]
    x = torch.full((2, 4), float(rank + 1))
    model(x).sum().backward()
    local = [p.grad.clone() for p in model.parameters()]
    bucket = cd.GradientBucket(params)
    bucket.all_reduce(average=True)
    assert dead.grad is None
    torch.save({"local": local, "reduced": [p.grad.clone() for p in model.parameters()]}, f"{out}.{rank}")
    assert abs(cd.max_over_ranks(float(rank), "cpu") - (world - 1)) < 1e-12
    torch.distributed.destroy_process_group()


def test_gradient_bucket_allreduce_world2(tmp_path):
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    a, b = torch.load(out + ".0"), torch.load(out + ".1")
    for la, lb, ra, rb in zip(a["local"], b["local"], a["reduced"], b["reduced"]):
        assert torch.allclose(ra, (la + lb) / 2) and torch.equal(ra, rb)
