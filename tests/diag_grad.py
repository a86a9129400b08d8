"""Diagnostic (not a test): per-tensor gradient error of the CUDA path and of the fp32 CPU oracle vs the fp64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from conftest import rel_l2
from cosmology_gnn_simulation_b200 import synthetic
from cosmology_gnn_simulation_b200.graph import Data
from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode
from cosmology_gnn_simulation_b200.loss import combined_loss
from oracle import knn_ref, model_ref

def run(message, n, k, L, H, nh, M, precision="fp32", seed=0):
    pos = synthetic.positions(n, "uniform", 1.0, seed=seed)
    ext = knn_ref.knn_kdtree(pos, 1.0, k)
    ei = torch.from_numpy(knn_ref.edge_index_from_ext(ext, n))
    p = torch.from_numpy(pos); d = p[ei[0]] - p[ei[1]]
    ea = torch.cat([d, d.norm(dim=-1, keepdim=True)], dim=-1)
    gen = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(n, 17, generator=gen); ya, yt = torch.randn(n, 3, generator=gen), torch.randn(n, 1, generator=gen)
    params = model_ref.init_params(L, H, nh, M, 3, seed=seed)
    res = {}
    for dt in (torch.float64, torch.float32):
        pp = {k_: v.to(dt).requires_grad_(True) for k_, v in params.items()}
        o = model_ref.forward(pp, x.to(dt), ei, ea.to(dt), nh, M, message=message)
        model_ref.loss(o["acceleration"], o["temp_rate"], ya.to(dt), yt.to(dt), 0.01, w_mom=0.1)["loss"].backward()
        res[dt] = (o, pp)
    dev = torch.device("cuda", 0)
    model = EncodeProcessDecode(L, H, nh, M, 3, message=message, precision=precision); model.load_state_dict(params); model = model.to(dev)
    graph = Data(x=x.to(dev), edge_index=ei.to(dev), edge_attr=ea.to(dev), y_acc=ya.to(dev), y_temp_rate=yt.to(dev))
    pred = model(graph); combined_loss(pred, graph, 0.01, 1.0, 1.0, 0.1)["loss"].backward()
    o64, p64 = res[torch.float64]; o32, p32 = res[torch.float32]
    print(f"== {message} n={n} k={k} L={L} H={H} nh={nh} M={M} {precision}")
    print("  acc  gpu %.2e cpu32 %.2e" % (rel_l2(pred["acceleration"].detach().cpu(), o64["acceleration"].detach()), rel_l2(o32["acceleration"].detach(), o64["acceleration"].detach())))
    worst = []
    for name, prm in model.named_parameters():
        if p64[name].grad is None or prm.grad is None: continue
        worst.append((rel_l2(prm.grad.cpu(), p64[name].grad), rel_l2(p32[name].grad, p64[name].grad), name))
    worst.sort(reverse=True)
    for g, c, name in worst[:6]:
        print("  %-40s gpu %.2e cpu32 %.2e" % (name, g, c))
    print("  median gpu %.2e cpu32 %.2e" % (sorted(w[0] for w in worst)[len(worst)//2], sorted(w[1] for w in worst)[len(worst)//2]))

if __name__ == "__main__":
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    for message in ("sender", "edge"):
        run(message, 333, 8, 128, 128, 2, 2, prec)
        run(message, 1000, 16, 64, 64, 2, 3, prec)
        run(message, 4096, 16, 64, 64, 2, 5, prec)
