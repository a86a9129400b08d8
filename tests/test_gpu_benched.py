"""Parity of exactly what bench.py measures: precision 'bf16x3', message 'edge', the BASELINE config-2 shape
(32 768 particles, k = 16, latent 128, 10 message-passing steps) against the float64 oracle, and a 200-step
optimizer trajectory against the oracle's own training run.

Edge-message semantics are oracle-only: the reference as it runs sums SENDER latents and all five
reference-generated fixtures pin that mode; `message="edge"` is the Interaction Network the north star describes,
restated by `oracle/model_ref.py` (graph_network.py:83-101 with the updated edge latent as the message).

Bars (north_star): outputs / loss rel-L2 <= 1e-3.  Gradients: <= 1e-3 of the float64 truth, or no further from it
than 2x the reference's own fp32 CPU arithmetic (the oracle run in float32) -- a ReLU pre-activation within
rounding of zero takes the other subgradient in ANY finite-precision implementation; every
parameter gradient is printed with that floor beside it, and the table goes to gpurun_out/ for profiles/.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-3


def _graph_and_oracle_inputs(n, k, seed=0):
    from cosmology_gnn_simulation_b200 import synthetic
    from cosmology_gnn_simulation_b200.data_utils import preprocess
    dev = torch.device("cuda", 0)
    box = synthetic.make_box(n, "uniform", seed=seed)
    md = box["metadata"]
    torch.manual_seed(0)
    g = preprocess(box["Coordinates"][:5], box["InternalEnergy"][:5], md, box["Coordinates"][5:6],
                   box["InternalEnergy"][5:6], num_neighbors=k, dt=md["dt"], box_size=md["box_size"], device=dev)
    cpu = dict(x=g.x.cpu(), ei=g.edge_index.cpu(), ea=g.edge_attr.cpu(), ya=g.y_acc.cpu(), yt=g.y_temp_rate.cpu())
    return g, cpu, md


def _oracle_step(params, cpu, md, M, dtype, message="edge", checkpoint=True):
    from oracle import model_ref
    pp = {k_: v.detach().to(dtype).clone().requires_grad_(True) for k_, v in params.items()}
    oo = model_ref.forward(pp, cpu["x"].to(dtype), cpu["ei"], cpu["ea"].to(dtype), 2, M, message, checkpoint_steps=checkpoint)
    ll = model_ref.loss(oo["acceleration"], oo["temp_rate"], cpu["ya"].to(dtype), cpu["yt"].to(dtype), md["dt"], w_mom=0.1)
    ll["loss"].backward()
    return pp, oo, ll


def test_config2_benched_mode_matches_float64_oracle():
    """BASELINE configs[1] exactly as bench.py runs it (train.py:250-265 on the kernels)."""
    from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode
    from cosmology_gnn_simulation_b200.loss import combined_loss
    from oracle import model_ref
    n, k, L, M = 32 ** 3, 16, 128, 10
    g, cpu, md = _graph_and_oracle_inputs(n, k)
    params = model_ref.init_params(L, L, 2, M, 3, seed=0)
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    p64, o64, l64 = _oracle_step(params, cpu, md, M, torch.float64)
    p32, o32, l32 = _oracle_step(params, cpu, md, M, torch.float32)

    model = EncodeProcessDecode(L, L, 2, M, 3, message="edge", precision="bf16x3")
    model.load_state_dict(params)
    model = model.to(g.x.device)
    pred = model(g)
    ls = combined_loss(pred, g, md["dt"], 1.0, 1.0, 0.1)
    ls["loss"].backward()
    torch.cuda.synchronize()

    report = {"config": f"N={n} k={k} L={L} M={M} message=edge precision=bf16x3", "bar": TOL, "rows": []}
    out_acc = rel_l2(pred["acceleration"].detach().cpu(), o64["acceleration"].detach())
    out_temp = rel_l2(pred["temp_rate"].detach().cpu(), o64["temp_rate"].detach())
    loss_rel = abs(ls["loss"].item() - l64["loss"].item()) / abs(l64["loss"].item())
    report["outputs"] = {"acceleration": out_acc, "temp_rate": out_temp, "loss": loss_rel,
                         "fp32_oracle_acceleration": rel_l2(o32["acceleration"].detach(), o64["acceleration"].detach())}
    failures = []
    for name, prm in model.named_parameters():
        ref = p64[name].grad
        assert ref is not None and prm.grad is not None, name
        err = rel_l2(prm.grad.cpu(), ref)
        floor = rel_l2(p32[name].grad, ref)
        report["rows"].append({"param": name, "cuda_vs_fp64": err, "fp32_oracle_vs_fp64": floor, "within_1e-3": err <= TOL})
        if not err <= max(TOL, 2.0 * floor):
            failures.append((name, err, floor))
    worst = max(r["cuda_vs_fp64"] for r in report["rows"])
    report["worst_gradient"] = worst
    report["gradients_within_1e-3"] = sum(r["within_1e-3"] for r in report["rows"])
    report["gradients_total"] = len(report["rows"])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_config2_bf16x3_edge.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(f"\nconfig2 bf16x3/edge vs fp64 oracle: acc {out_acc:.2e} temp {out_temp:.2e} loss {loss_rel:.2e}; "
          f"gradients worst {worst:.2e}, {report['gradients_within_1e-3']}/{len(report['rows'])} within 1e-3")
    for r in report["rows"]:
        print(f"  {r['param']:45s} cuda {r['cuda_vs_fp64']:.2e}   fp32 oracle {r['fp32_oracle_vs_fp64']:.2e}")
    assert out_acc <= TOL and out_temp <= TOL and loss_rel <= TOL
    assert not failures, failures


def test_training_trajectory_matches_oracle():
    """200 Adam steps (train.py:183,263-265) on one sample: the loss curve of the CUDA path (bf16x3, edge messages)
    against the oracle trained in float64, with the oracle trained in float32 -- the reference's arithmetic --
    as the yardstick for how far two correct implementations drift apart.

    Training a ReLU network is chaotic in the rounding: the oracle's OWN float32 and float64 runs agree to 1e-3 over
    the first ten steps and are a factor 3 apart in loss by step 60 before both settle at the same level (measured:
    profiles/r02_trajectory_bf16x3_edge.json; plain SGD behaves alike).  So "the curves agree to 1e-3" can only be asked
    where two correct implementations still agree -- the first five steps; steps 5-20, where the float32 oracle itself
    leaves the 1e-3 band, and the rest of the trajectory are held to the yardstick: the CUDA path may not stray further
    from the float64 curve than three times (first 20 steps) / twice (largest gap of all 200, in units of the initial
    loss) what the float32 oracle does, and it must train to the same final loss within a factor 2."""
    from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode
    from cosmology_gnn_simulation_b200.loss import combined_loss
    from oracle import model_ref
    n, k, L, M, steps, lr = 256, 16, 128, 10, 200, 1e-4
    g, cpu, md = _graph_and_oracle_inputs(n, k, seed=3)
    params = model_ref.init_params(L, L, 2, M, 3, seed=1)
    torch.set_num_threads(len(os.sched_getaffinity(0)))

    def train_oracle(dtype):
        pp = {k_: v.detach().to(dtype).clone().requires_grad_(True) for k_, v in params.items()}
        opt = torch.optim.Adam(list(pp.values()), lr=lr)
        x, ea, ya, yt = cpu["x"].to(dtype), cpu["ea"].to(dtype), cpu["ya"].to(dtype), cpu["yt"].to(dtype)
        curve = []
        for _ in range(steps):
            opt.zero_grad()
            oo = model_ref.forward(pp, x, cpu["ei"], ea, 2, M, "edge")
            ll = model_ref.loss(oo["acceleration"], oo["temp_rate"], ya, yt, md["dt"], w_mom=0.1)
            ll["loss"].backward()
            opt.step()
            curve.append(float(ll["loss"].detach()))
        return np.array(curve)

    c64 = train_oracle(torch.float64)
    c32 = train_oracle(torch.float32)

    model = EncodeProcessDecode(L, L, 2, M, 3, message="edge", precision="bf16x3")
    model.load_state_dict(params)
    model = model.to(g.x.device)
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    losses = []
    for _ in range(steps):
        opt.zero_grad()
        ls = combined_loss(model(g), g, md["dt"], 1.0, 1.0, 0.1)
        ls["loss"].backward()
        opt.step()
        losses.append(ls["loss"].detach())
    cg = torch.stack(losses).cpu().double().numpy()

    dev_cuda = np.abs(cg - c64) / np.abs(c64)
    dev_fp32 = np.abs(c32 - c64) / np.abs(c64)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "trajectory_bf16x3_edge.json"), "w") as f:
        json.dump({"config": f"N={n} k={k} L={L} M={M} Adam lr={lr} {steps} steps, message=edge, precision=bf16x3",
                   "loss_first": float(c64[0]), "loss_last": float(c64[-1]),
                   "max_rel_dev_cuda_vs_fp64": float(dev_cuda.max()), "max_rel_dev_fp32_oracle_vs_fp64": float(dev_fp32.max()),
                   "mean_rel_dev_cuda_vs_fp64": float(dev_cuda.mean()), "mean_rel_dev_fp32_oracle_vs_fp64": float(dev_fp32.mean()),
                   "first10_max_rel_dev_cuda_vs_fp64": float(dev_cuda[:10].max()),
                   "first10_max_rel_dev_fp32_oracle_vs_fp64": float(dev_fp32[:10].max()),
                   "largest_gap_over_initial_loss_cuda": float(np.abs(cg - c64).max() / c64[0]),
                   "largest_gap_over_initial_loss_fp32_oracle": float(np.abs(c32 - c64).max() / c64[0]),
                   "curve_fp64": c64.tolist(), "curve_fp32": c32.tolist(), "curve_cuda": cg.tolist()}, f)
    print(f"\ntrajectory: loss {c64[0]:.4f} -> {c64[-1]:.4f}; rel deviation from the fp64 oracle, first 10 steps max / all steps mean / max: "
          f"CUDA bf16x3 {dev_cuda[:10].max():.2e} / {dev_cuda.mean():.2e} / {dev_cuda.max():.2e}, "
          f"fp32 oracle {dev_fp32[:10].max():.2e} / {dev_fp32.mean():.2e} / {dev_fp32.max():.2e}")
    assert c64[-1] < 0.9 * c64[0], "the trajectory must actually train"
    assert dev_cuda[:5].max() <= TOL, dev_cuda[:5]
    assert dev_cuda[:20].max() <= max(TOL, 3.0 * dev_fp32[:20].max()), (dev_cuda[:20].max(), dev_fp32[:20].max())
    gap_cuda, gap_fp32 = np.abs(cg - c64).max() / c64[0], np.abs(c32 - c64).max() / c64[0]
    assert gap_cuda <= max(TOL, 2.0 * gap_fp32), (gap_cuda, gap_fp32)
    assert 0.5 <= cg[-1] / c64[-1] <= 2.0, (cg[-1], c64[-1])
