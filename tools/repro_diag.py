#!/usr/bin/env python
"""Which tensors of a whole training application differ between repetitions (every kernel is deterministic: any difference is a
race).  python tools/repro_diag.py [--n 20000] [--k 16] [--M 3] [--reps 6] [--grad-stream bf16|fp32]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from cosmology_gnn_simulation_b200 import synthetic  # noqa: E402
from cosmology_gnn_simulation_b200.data_utils import preprocess  # noqa: E402
from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode  # noqa: E402
from cosmology_gnn_simulation_b200.loss import combined_loss  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=20000)
ap.add_argument("--k", type=int, default=16)
ap.add_argument("--M", type=int, default=3)
ap.add_argument("--reps", type=int, default=6)
ap.add_argument("--grad-stream", default="bf16")
a = ap.parse_args()
dev = torch.device("cuda", 0)
box = synthetic.make_box(a.n, "uniform", seed=11)
md = box["metadata"]
g = preprocess(box["Coordinates"][:5], box["InternalEnergy"][:5], md, box["Coordinates"][5:6], box["InternalEnergy"][5:6],
               num_neighbors=a.k, dt=md["dt"], box_size=md["box_size"], device=dev)
torch.manual_seed(0)
model = EncodeProcessDecode(128, 128, 2, a.M, 3, message="edge", precision="bf16x3", grad_stream=a.grad_stream).to(dev)
ref, bad = None, 0
for rep in range(a.reps):
    for p in model.parameters():
        p.grad = None
    pred = model(g)
    combined_loss(pred, g, md["dt"], 1.0, 1.0, 0.1)["loss"].backward()
    cur = {"acc": pred["acceleration"].detach().clone(), "temp": pred["temp_rate"].detach().clone()}
    cur.update({nm: p.grad.clone() for nm, p in model.named_parameters() if p.grad is not None})
    if ref is None:
        ref = cur
    else:
        diff = [nm for nm in cur if not torch.equal(cur[nm], ref[nm])]
        if diff:
            bad += 1
            print(f"rep {rep}: {len(diff)} tensors differ: {diff[:12]}")
print(f"n={a.n} k={a.k} M={a.M} grad_stream={a.grad_stream} env={ {k_: v for k_, v in os.environ.items() if k_.startswith('CGNN_')} }: "
      f"{bad} of {a.reps - 1} repetitions differ")
