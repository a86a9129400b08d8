#!/usr/bin/env python
"""Where the end-to-end step (host buffers -> loss on the host) spends its time: wall clock per phase with a
device synchronisation after each (so the phases do not overlap; the sum is an upper bound of the real step)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from cosmology_gnn_simulation_b200 import synthetic  # noqa: E402
from cosmology_gnn_simulation_b200.data_utils import preprocess  # noqa: E402
from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode  # noqa: E402
from cosmology_gnn_simulation_b200.loss import combined_loss  # noqa: E402

n, k, L, M = 32768, 16, 128, 10
dev = torch.device("cuda", 0)
box = synthetic.make_box(n, "uniform", seed=0)
md = box["metadata"]
ch, eh = box["Coordinates"].pin_memory(), box["InternalEnergy"].pin_memory()
torch.manual_seed(0)
model = EncodeProcessDecode(L, L, 2, M, 3, message="edge", precision="bf16x3").to(dev)


def step(timed):
    marks = []

    def mark(name):
        if timed:
            torch.cuda.synchronize()
        marks.append((name, time.perf_counter()))

    mark("start")
    c, u = ch.to(dev, non_blocking=True), eh.to(dev, non_blocking=True)
    mark("h2d")
    g = preprocess(c[:5], u[:5], md, c[5:6], u[5:6], num_neighbors=k, dt=md["dt"], box_size=md["box_size"], device=dev)
    mark("preprocess")
    for p in model.parameters():
        p.grad = None
    pred = model(g)
    mark("forward")
    ls = combined_loss(pred, g, md["dt"], 1.0, 1.0, 0.1)
    mark("loss")
    ls["loss"].backward()
    mark("backward")
    out = torch.stack([ls["loss"].detach(), ls["acc_loss"], ls["temp_rate_loss"], ls["momentum_loss"]]).cpu()
    mark("d2h")
    return marks


for _ in range(3):
    step(False)
torch.cuda.synchronize()
acc = {}
for _ in range(5):
    m = step(True)
    for (a, ta), (b, tb) in zip(m[:-1], m[1:]):
        acc[b] = acc.get(b, 0.0) + (tb - ta) / 5
print("phase ms (synchronised):", {k_: round(v * 1e3, 3) for k_, v in acc.items()}, "sum", round(sum(acc.values()) * 1e3, 3))
t0 = time.perf_counter()
for _ in range(5):
    step(False)
torch.cuda.synchronize()
print("unsynchronised e2e step ms:", round((time.perf_counter() - t0) / 5 * 1e3, 3))
# host-side issue time of forward+backward (how far ahead of the GPU the CPU can run)
torch.cuda.synchronize()
g = preprocess(ch[:5].to(dev), eh[:5].to(dev), md, ch[5:6].to(dev), eh[5:6].to(dev), num_neighbors=k, dt=md["dt"], box_size=md["box_size"], device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
pred = model(g)
ls = combined_loss(pred, g, md["dt"], 1.0, 1.0, 0.1)
ls["loss"].backward()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"fwd+loss+bwd: host issue {1e3 * (t1 - t0):.2f} ms, device done after {1e3 * (t2 - t0):.2f} ms")
