#!/usr/bin/env python
"""One eager training application (forward + loss + backward) of a bench workload inside a cudaProfiler range, for
`ncu --profile-from-start off` (launch list of exactly one step), or one call of a single phase for a full capture:

  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv \
      python tools/profile_step.py --workload config2
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_wgrad -c 1 -o out \
      python tools/profile_step.py --workload config2 --phase edge_bwd
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
from cosmology_gnn_simulation_b200 import ops, synthetic  # noqa: E402
from cosmology_gnn_simulation_b200.data_utils import preprocess  # noqa: E402
from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode  # noqa: E402
from cosmology_gnn_simulation_b200.loss import combined_loss  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config2", choices=sorted(bench.WORKLOADS))
ap.add_argument("--phase", default="step", choices=["step", "edge_fwd", "edge_bwd", "knn"])
ap.add_argument("--message", default="edge")
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--grad-stream", default="bf16", choices=["bf16", "fp32"])
a = ap.parse_args()
n, k, L, M, kind = bench.WORKLOADS[a.workload]
dev = torch.device("cuda", 0)
box = synthetic.make_box(n, kind, seed=0)
md = box["metadata"]
g = preprocess(box["Coordinates"][:5], box["InternalEnergy"][:5], md, box["Coordinates"][5:6], box["InternalEnergy"][5:6],
               num_neighbors=k, dt=md["dt"], box_size=md["box_size"], device=dev)
torch.manual_seed(0)
model = EncodeProcessDecode(L, L, 2, M if a.phase == "step" else 1, 3, message=a.message, precision=a.precision,
                            grad_stream=a.grad_stream).to(dev)


def step():
    for p in model.parameters():
        p.grad = None
    combined_loss(model(g), g, md["dt"], 1.0, 1.0, 0.1)["loss"].backward()


if a.phase == "step":
    run = step
elif a.phase == "knn":
    pos = g.pos.contiguous()
    run = lambda: ops.edge_features(pos, ops.knn_periodic(pos, md["box_size"], k), md["box_size"])  # noqa: E731
else:
    step()                                                     # materialise the lazy layers
    from cosmology_gnn_simulation_b200.graph_network import _mlp_params
    blk = model.processor[0]
    p = _mlp_params(blk.edge_model[0], blk.edge_model[1])
    gen = torch.Generator(device=dev).manual_seed(1)
    h = torch.randn(n, L, device=dev, generator=gen)
    e = torch.randn(n * k, L, device=dev, generator=gen)
    senders = g._cgnn_senders
    if n >= (1 << 17):                                         # the model's own Z-order renumbering of large graphs (graph_network.py)
        from cosmology_gnn_simulation_b200.graph_network import _morton_order
        o = _morton_order(g.pos, g.box_size.reshape(-1)[0])
        iv = torch.empty_like(o)
        iv[o] = torch.arange(n, device=dev)
        senders = iv.to(torch.int32)[senders.view(n, k)[o].long()].reshape(-1).contiguous()
    e_out, agg = torch.empty_like(e), torch.empty_like(h)
    if a.phase == "edge_fwd":
        run = lambda: ops.mp_edge_fwd(p, h, e, senders, k, e_out, agg, a.precision)  # noqa: E731
    else:
        rowptr, perm = ops.csr_transpose(senders, n)
        # the precision the model itself picks for this stream length (bfloat16 gradient streams for long edge streams)
        from cosmology_gnn_simulation_b200.graph_network import GRAD16_MIN_ROWS
        prec_e = "bf16x3g" if a.precision == "bf16x3" and a.grad_stream == "bf16" and n * k >= GRAD16_MIN_ROWS else a.precision
        de_next = torch.randn_like(e).to(ops.grad_stream_dtype(prec_e))
        dagg, dh = torch.randn_like(h), torch.zeros_like(h)
        run = lambda: ops.mp_edge_bwd(p, h, e, senders, rowptr, perm, k, de_next, dagg, de_next, dh, None, prec_e)  # noqa: E731
for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
run()
t.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"{a.workload} {a.phase}: {s.elapsed_time(t):.3f} ms")
