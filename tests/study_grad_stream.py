"""Error study (not a test; CPU only): what a 2-byte gradient stream costs.

The tensor-core backward of a long row stream (csrc/tc_model.cu) writes three gradient intermediates per MLP to HBM --
dY (the LayerNorm backward's result), G2 and G1 (the gradients at the two hidden pre-activations) -- each read again by the
next dgrad chain, by a weight-gradient kernel and (G1) by the sender scatter: 10 of the 21 row-stream passes of a processor
step.  Stored as bfloat16 they are 5.  This script measures on the oracle what that rounding does to every parameter gradient:

  truth      float64 oracle
  fp32       float32 oracle (the yardstick: two correct float32 implementations differ by this much)
  g-bf16     float32 oracle whose gradients at dY / G2 / G1 of the chosen MLPs are rounded to bfloat16 (round to nearest even)
             before they are used by the weight gradient AND the input gradient -- a conservative model of the kernels, which
             take the per-receiver sums of G1 from the unrounded accumulators
  g-fp16     the same with float16 (no scaling)
  ... + de   additionally the gradient stream de^t that is carried from step to step (E x L) rounded after every step
  f64 + ...  the float64 oracle with the same rounding: the rounding's own contribution, free of float32 ReLU-gate noise

Forward values are untouched (the activations that decide the ReLU gates stay float32), so outputs and loss do not move.
Usage: python tests/study_grad_stream.py [n] [k] [M]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cosmology_gnn_simulation_b200 import synthetic  # noqa: E402
from oracle import knn_ref, model_ref  # noqa: E402


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


class RoundGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, dtype):
        ctx.dtype = dtype
        return z.view_as(z)

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.dtype).to(g.dtype), None


def mlp_ln(params, prefix, z, n_hidden, gdtype):
    names = model_ref.mlp_names(prefix + ".0", n_hidden)
    for i, name in enumerate(names):
        z = z @ params[name + ".weight"].T + params[name + ".bias"]
        if gdtype is not None:
            z = RoundGrad.apply(z, gdtype)          # G1, G2 (hidden pre-activations), dY (LayerNorm input)
        if i < n_hidden:
            z = torch.relu(z)
    return model_ref.layer_norm(z, params[prefix + ".1.weight"], params[prefix + ".1.bias"])


def forward(params, x, edge_index, edge_attr, n_hidden, n_steps, gdtype, which):
    src, dst = edge_index[0], edge_index[1]
    g_edge = gdtype if "edge" in which else None
    g_node = gdtype if "node" in which else None
    h = mlp_ln(params, "encoder.node_model", x, n_hidden, g_node)
    e = mlp_ln(params, "encoder.edge_model", edge_attr, n_hidden, g_edge)
    for t in range(n_steps):
        u_e = mlp_ln(params, f"processor.{t}.edge_model", torch.cat([h[src], h[dst], e], dim=-1), n_hidden, g_edge)
        agg = torch.zeros_like(h).index_add_(0, dst, u_e)
        u_n = mlp_ln(params, f"processor.{t}.node_model", torch.cat([h, agg], dim=-1), n_hidden, g_node)
        h, e = h + u_n, e + u_e
        if "de" in which:
            e = RoundGrad.apply(e, gdtype)          # the gradient stream de^t itself (E x L, carried from step to step)
    return {"acceleration": model_ref.mlp(params, "decoder_acc", h, n_hidden),
            "temp_rate": model_ref.mlp(params, "decoder_temp_rate", h, n_hidden)}


def run(n, k, M, L=128, seed=0, variants=None, quiet=False):
    pos = synthetic.positions(n, "uniform", 1.0, seed=seed)
    ext = knn_ref.knn_kdtree(pos, 1.0, k)
    ei = torch.from_numpy(knn_ref.edge_index_from_ext(ext, n))
    p = torch.from_numpy(pos)
    d = p[ei[0]] - p[ei[1]]
    ea = torch.cat([d, d.norm(dim=-1, keepdim=True)], dim=-1)
    gen = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(n, 17, generator=gen)
    ya, yt = torch.randn(n, 3, generator=gen), torch.randn(n, 1, generator=gen)
    params = model_ref.init_params(L, L, 2, M, 3, seed=seed)

    def grads(dt, gdtype, which):
        pp = {k_: v.detach().clone().to(dt).requires_grad_(True) for k_, v in params.items()}
        o = forward(pp, x.to(dt), ei, ea.to(dt), 2, M, gdtype, which)
        model_ref.loss(o["acceleration"], o["temp_rate"], ya.to(dt), yt.to(dt), 0.01, w_mom=0.1)["loss"].backward()
        return {k_: v.grad for k_, v in pp.items()}

    truth = grads(torch.float64, None, "")
    results = {}
    if not quiet:
        print(f"# n={n} k={k} L={L} M={M} edge messages; rel-L2 of every parameter gradient against the float64 oracle")
        print(f"{'variant':<28} {'worst':>9} {'median':>9}  worst tensor")
    for label, dt, gd, which in (("fp32", torch.float32, None, ""),
                                 ("g-bf16 edge MLPs", torch.float32, torch.bfloat16, "edge"),
                                 ("g-bf16 edge+node MLPs", torch.float32, torch.bfloat16, "edge node"),
                                 ("g-fp16 edge MLPs", torch.float32, torch.float16, "edge"),
                                 ("f64 + g-bf16 edge MLPs", torch.float64, torch.bfloat16, "edge"),
                                 ("f64 + g-bf16 edge+node", torch.float64, torch.bfloat16, "edge node"),
                                 ("f64 + g-bf16 edge + de", torch.float64, torch.bfloat16, "edge de"),
                                 ("g-bf16 edge + de", torch.float32, torch.bfloat16, "edge de")):
        if variants is not None and label not in variants:
            continue
        g = grads(dt, gd, which)
        errs = sorted(((rel_l2(g[k_], truth[k_]), k_) for k_ in truth), reverse=True)
        results[label] = (errs[0][0], errs[len(errs) // 2][0], errs[0][1])
        if not quiet:
            print(f"{label:<28} {errs[0][0]:9.2e} {errs[len(errs) // 2][0]:9.2e}  {errs[0][1]}")
            sys.stdout.flush()
    return results


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    run(a[0] if len(a) > 0 else 4096, a[1] if len(a) > 1 else 16, a[2] if len(a) > 2 else 10)
