"""GPU: a captured-and-replayed training step (graphed.py) equals the eager step bit for bit, also for a second
graph of the same size that was not the one captured."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("message,precision", [("edge", "bf16x3"), ("sender", "fp32")])
def test_graphed_step_equals_eager(message, precision):
    from cosmology_gnn_simulation_b200 import synthetic
    from cosmology_gnn_simulation_b200.data_utils import preprocess
    from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode
    from cosmology_gnn_simulation_b200.graphed import GraphedTrainStep
    from cosmology_gnn_simulation_b200.loss import combined_loss
    dev = torch.device("cuda", 0)
    n, k, L, M = 3000, 16, 128, 3
    graphs, md = [], None
    for seed in (0, 1):
        box = synthetic.make_box(n, "uniform", seed=seed)
        md = box["metadata"]
        graphs.append(preprocess(box["Coordinates"][:5], box["InternalEnergy"][:5], md, box["Coordinates"][5:6],
                                 box["InternalEnergy"][5:6], num_neighbors=k, dt=md["dt"], box_size=md["box_size"], device=dev))
    torch.manual_seed(0)
    model = EncodeProcessDecode(L, L, 2, M, 3, message=message, precision=precision).to(dev)
    model(graphs[0])                                   # materialise the lazy layers (no backward yet)
    eager = copy.deepcopy(model)

    def loss_fn(pred, g):
        return combined_loss(pred, g, md["dt"], 1.0, 1.0, 0.1)

    step = GraphedTrainStep(model, loss_fn)
    step.capture(graphs[0])
    assert step.launches_per_step > 0
    for g in (graphs[1], graphs[0], graphs[1]):
        out = step(g)
        for p in eager.parameters():
            p.grad = None
        ref = loss_fn(eager(g), g)
        ref["loss"].backward()
        torch.cuda.synchronize()
        assert torch.equal(out["loss"], ref["loss"].detach())
        for (name, a), b in zip(model.named_parameters(), eager.parameters()):
            if b.grad is None:
                assert a.grad is None or float(a.grad.abs().max()) == 0.0, name
            else:
                assert torch.equal(a.grad, b.grad), name
