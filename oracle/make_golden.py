"""Generate tests/golden/*.npz by executing the REFERENCE's own files.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py

`/root/reference/graph_network.py` and `/root/reference/data_utils.py` are imported unmodified;
only the three third-party symbols that are not installable here are stubbed:

  torch_geometric.data.Data           attribute bag
  torch_geometric.nn.MessagePassing   nn.Module whose propagate(edge_index, x=..., edge_attr=...)
                                      = zeros_like(x).index_add_(0, edge_index[1], x[edge_index[0]])
                                      (PyG 2.6.1 default message(x_j) = x_j, aggr='add'; this ENCODES
                                      the un-vendored semantics, it does not verify them)
  torch_cluster.knn                   exhaustive search, canonical fp32 distance, stable sort
                                      (ENCODES the contract of torch-cluster 1.6.3, unpinned)

Everything else that runs is the reference's own code, so the fixtures pin the MLP/LayerNorm
stack, residual wiring, concat orders, state_dict key names, feature/target arithmetic, noise RNG
consumption and the loss.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")


def install_stubs():
    class Data:
        def __init__(self, **kw):
            for k, v in kw.items():
                setattr(self, k, v)

    class MessagePassing(nn.Module):
        def __init__(self, aggr="add"):
            super().__init__()
            assert aggr == "add"

        def propagate(self, edge_index, x=None, **unused):
            return torch.zeros_like(x).index_add_(0, edge_index[1], x[edge_index[0]])

    def knn(x, y, k):
        rows = []
        for i in range(y.shape[0]):
            d = x - y[i]
            d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            rows.append(torch.argsort(d2, stable=True)[:k])
        col = torch.stack(rows).reshape(-1)
        row = torch.arange(y.shape[0]).repeat_interleave(k)
        return torch.stack([row, col], dim=0)

    pyg = types.ModuleType("torch_geometric")
    pyg_data = types.ModuleType("torch_geometric.data")
    pyg_nn = types.ModuleType("torch_geometric.nn")
    pyg_data.Data = Data
    pyg_nn.MessagePassing = MessagePassing
    pyg_nn.knn_graph = None
    pyg.data, pyg.nn = pyg_data, pyg_nn
    tc = types.ModuleType("torch_cluster")
    tc.knn = knn
    sys.modules.update({"torch_geometric": pyg, "torch_geometric.data": pyg_data,
                        "torch_geometric.nn": pyg_nn, "torch_cluster": tc,
                        "torch_scatter": types.ModuleType("torch_scatter")})
    return Data


def main():
    Data = install_stubs()
    sys.path.insert(0, REF)
    import data_utils as ref_du          # noqa: E402  (the reference's file)
    import graph_network as ref_gn       # noqa: E402  (the reference's file)
    sys.path.insert(0, ROOT)
    from cosmology_gnn_simulation_b200 import synthetic

    os.makedirs(OUT, exist_ok=True)

    # ---------------- preprocess fixtures ------------------------------------------------
    for tag, n, k, kind, noise in [("pre_uniform", 300, 8, "uniform", 0.0),
                                   ("pre_clustered_noise", 257, 16, "clustered", 3e-4)]:
        box = synthetic.make_box(n, kind, window=5, seed=7)
        coords, energy, md = box["Coordinates"], box["InternalEnergy"], box["metadata"]
        torch.manual_seed(1234)
        g = ref_du.preprocess(position_seq=coords[:5].clone(), temperature_seq=energy[:5].clone(),
                              metadata=md, target_position=coords[5:6].clone(),
                              target_temperature=energy[5:6].clone(), noise_std=noise,
                              num_neighbors=k, dt=md["dt"], box_size=md["box_size"])
        rng_after = torch.rand(4)       # pins how much RNG preprocess consumed
        np.savez_compressed(
            os.path.join(OUT, tag + ".npz"),
            coords=coords.numpy(), energy=energy.numpy(),
            md_keys=np.array(list(md.keys())), md_vals=np.array([np.ravel(md[k_])[0] for k_ in md], dtype=np.float64),
            k=k, noise_std=noise, seed=1234,
            x=g.x.numpy(), edge_index=g.edge_index.numpy(), edge_attr=g.edge_attr.numpy(),
            y_acc=g.y_acc.numpy(), y_temp_rate=g.y_temp_rate.numpy(), pos=g.pos.numpy(),
            dt=g.dt.numpy(), box_size=g.box_size.numpy(), rng_after=rng_after.numpy())
        print("wrote", tag, "E =", g.edge_index.shape[1])

    # ---------------- model fixtures -----------------------------------------------------
    # (weight seeds are chosen so that no ReLU pre-activation on a gradient path sits within fp32
    #  rounding of zero: such a gate flips between two correct fp32 implementations and moves the
    #  gradient by O(1e-3) -- seed 5 did exactly that for model_deepmlp, see DESIGN.md "ReLU gates")
    for tag, n, k, L, H, nh, M, out, wseed in [("model_tiny", 48, 6, 32, 48, 2, 2, 3, 5),
                                               ("model_small", 384, 16, 64, 64, 2, 5, 3, 5),
                                               ("model_deepmlp", 96, 8, 32, 32, 3, 3, 3, 6)]:
        box = synthetic.make_box(n, "uniform", window=5, seed=11)
        md = box["metadata"]
        torch.manual_seed(99)
        g = ref_du.preprocess(position_seq=box["Coordinates"][:5].clone(),
                              temperature_seq=box["InternalEnergy"][:5].clone(), metadata=md,
                              target_position=box["Coordinates"][5:6].clone(),
                              target_temperature=box["InternalEnergy"][5:6].clone(),
                              noise_std=0.0, num_neighbors=k, dt=md["dt"], box_size=md["box_size"])
        torch.manual_seed(wseed)
        model = ref_gn.EncodeProcessDecode(L, H, nh, M, out)
        x = g.x.clone().requires_grad_(True)
        ea = g.edge_attr.clone().requires_grad_(True)
        graph = Data(x=x, edge_index=g.edge_index, edge_attr=ea)
        pred = model(graph)
        # loss exactly as train.py:255-260 with weights (1, 1, 0.1), one graph
        mse = nn.MSELoss()
        acc_loss = mse(pred["acceleration"], g.y_acc)
        temp_loss = mse(pred["temp_rate"], g.y_temp_rate)
        dv = pred["acceleration"] * md["dt"]
        mom = 0.1 * torch.sum(torch.sum(dv, dim=0) ** 2) / 1
        total = 1.0 * acc_loss + 1.0 * temp_loss + mom
        total.backward()
        sd = {k_: v.detach().numpy() for k_, v in model.state_dict().items()}
        grads = {}
        for name, p in model.named_parameters():
            grads[name] = None if p.grad is None else p.grad.numpy()
        payload = dict(x=g.x.numpy(), edge_index=g.edge_index.numpy(), edge_attr=g.edge_attr.numpy(),
                       y_acc=g.y_acc.numpy(), y_temp_rate=g.y_temp_rate.numpy(),
                       dt=md["dt"], cfg=np.array([L, H, nh, M, out]),
                       acceleration=pred["acceleration"].detach().numpy(),
                       temp_rate=pred["temp_rate"].detach().numpy(),
                       loss=total.item(), acc_loss=acc_loss.item(), temp_loss=temp_loss.item(),
                       mom_loss=mom.item(), grad_x=x.grad.numpy(), grad_edge_attr=(np.zeros_like(g.edge_attr.numpy()) if ea.grad is None else ea.grad.numpy()),
                       grad_edge_attr_is_none=(ea.grad is None),
                       sd_keys=np.array(list(sd.keys())),
                       grad_none=np.array([k_ for k_, v in grads.items() if v is None]))
        for k_, v in sd.items():
            payload["sd/" + k_] = v
        for k_, v in grads.items():
            if v is not None:
                payload["grad/" + k_] = v
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), **payload)
        print("wrote", tag, "params", sum(v.size for v in sd.values()),
              "grad None:", len(payload["grad_none"]), "of", len(grads))


if __name__ == "__main__":
    main()
