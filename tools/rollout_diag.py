#!/usr/bin/env python
"""Where a rollout step spends its host and device time: preprocess (k-NN + features) and the model forward timed with and
without synchronisation, plus a cProfile of the unsynchronised loop (found the 15 ms cudaMemGetInfo call per forward)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from cosmology_gnn_simulation_b200 import synthetic, ops
from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode
from cosmology_gnn_simulation_b200.data_utils import preprocess
dev = torch.device('cuda', 0)
n, k, L, M = 32768, 16, 128, 10
box = synthetic.make_box(n, 'uniform', seed=0); md = box['metadata']
torch.manual_seed(0)
model = EncodeProcessDecode(L, L, 2, M, 3, message='edge', precision='bf16x3').to(dev).eval()
pos = box['Coordinates'][:5].to(dev); tmp = box['InternalEnergy'][:5].to(dev)
if tmp.dim() == 2: tmp = tmp.unsqueeze(-1)
def sync(): torch.cuda.synchronize()
with torch.no_grad():
    for it in range(6):
        sync(); t0 = time.perf_counter()
        g = preprocess(pos, tmp, md, noise_std=0.0, num_neighbors=k, box_size=md['box_size'], dt=md['dt'], device=dev)
        sync(); t1 = time.perf_counter()
        pred = model(g)
        sync(); t2 = time.perf_counter()
        print(f"preprocess {1e3*(t1-t0):.2f} ms  model {1e3*(t2-t1):.2f} ms")
    # without syncs
    sync(); t0 = time.perf_counter()
    for it in range(10):
        g = preprocess(pos, tmp, md, noise_std=0.0, num_neighbors=k, box_size=md['box_size'], dt=md['dt'], device=dev)
        pred = model(g)
    sync(); print(f"loop {1e2*(time.perf_counter()-t0):.2f} ms/step")
    sync(); t0 = time.perf_counter()
    for it in range(10):
        pred = model(g)
    sync(); print(f"model only {1e2*(time.perf_counter()-t0):.2f} ms/step")
    import cProfile, pstats
    pr = cProfile.Profile()
    sync(); pr.enable()
    for it in range(10):
        g = preprocess(pos, tmp, md, noise_std=0.0, num_neighbors=k, box_size=md['box_size'], dt=md['dt'], device=dev)
        pred = model(g)
    sync(); pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
