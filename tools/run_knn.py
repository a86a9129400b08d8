#!/usr/bin/env python
"""Times the periodic k-NN build alone (cgnn_knn_periodic + cgnn_edge_features) on uniform and clustered boxes."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from cosmology_gnn_simulation_b200 import ops, synthetic  # noqa: E402

d = torch.device("cuda", 0)
for n, k in ((32768, 16), (262144, 16), (2097152, 32)):
    for kind in ("uniform", "clustered"):
        pos = torch.from_numpy(synthetic.positions(n, kind, 1.0, seed=0)).to(d)
        for _ in range(2):
            ops.knn_periodic(pos, 1.0, k)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3):
            nbr = ops.knn_periodic(pos, 1.0, k)
            ops.edge_features(pos, nbr, 1.0)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 3
        print(f"k-NN + edge features  n={n:8d} k={k:2d} {kind:9s}: {ms:9.3f} ms  {n / ms / 1e3:8.1f} M particles/s")
