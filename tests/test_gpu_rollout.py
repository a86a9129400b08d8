"""GPU: device-resident rollout (rollout.py) against the oracle restatement of render_rollout.py:26-90."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,L", [("fp32", 32), ("bf16x3", 128)])
def test_rollout_matches_oracle(precision, L):
    from cosmology_gnn_simulation_b200 import synthetic
    from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode
    from cosmology_gnn_simulation_b200.rollout import rollout
    from oracle import model_ref, rollout_ref
    n, k, M, w, total = 600, 16, 2, 5, 9
    box = synthetic.make_box(n, "uniform", seed=2)
    md = box["metadata"]
    params = model_ref.init_params(L, L, 2, M, 3, seed=1)
    ref = rollout_ref.rollout(params, box, md, md["dt"], md["box_size"], w, 2, M, total, num_neighbors=k)
    model = EncodeProcessDecode(L, L, 2, M, 3, precision=precision)
    model.load_state_dict(params)
    model = model.to(torch.device("cuda", 0))
    data = {"Coordinates": box["Coordinates"][:w], "InternalEnergy": box["InternalEnergy"][:w]}
    got = rollout(model, data, md, 0.0, md["dt"], md["box_size"], window_size=w, num_neighbors=k, n_steps=total - w)
    assert got["Coordinates"].is_cuda and tuple(got["Coordinates"].shape) == (total, n, 3)
    # the given frames are passed through untouched
    assert torch.equal(got["Coordinates"][:w].cpu(), box["Coordinates"][:w].float())
    tol = 1e-5 if precision == "fp32" else 1e-3
    # positions live on a torus: compare through the wrapped difference
    d = got["Coordinates"].cpu() - ref["Coordinates"]
    d = d - torch.round(d / md["box_size"]) * md["box_size"]
    assert float(d.abs().max()) < tol * md["box_size"]
    assert rel_l2(got["InternalEnergy"].cpu(), ref["InternalEnergy"]) < tol
    # the reference's calling convention (length taken from the data) and determinism
    padded = {"Coordinates": torch.cat([box["Coordinates"][:w], torch.zeros(total - w, n, 3)]),
              "InternalEnergy": torch.cat([box["InternalEnergy"][:w], torch.zeros(total - w, n, 1)])}
    again = rollout(model, padded, md, 0.0, md["dt"], md["box_size"], window_size=w, num_neighbors=k)
    assert torch.equal(again["Coordinates"], got["Coordinates"])


def test_sharded_rollout_world1_equals_rollout():
    """rollout_slab re-partitions the box every step (migration = re-partition); on one rank it must reproduce `rollout`
    exactly -- the slab path only permutes the particles (x-sorted order) and the kernels are order-independent per row."""
    from cosmology_gnn_simulation_b200 import synthetic
    from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode
    from cosmology_gnn_simulation_b200.rollout import rollout, rollout_slab
    from oracle import model_ref
    n, k, L, M, w = 900, 16, 128, 2, 5
    box = synthetic.make_box(n, "uniform", seed=5)
    md = box["metadata"]
    model = EncodeProcessDecode(L, L, 2, M, 3, precision="bf16x3")
    model.load_state_dict(model_ref.init_params(L, L, 2, M, 3, seed=2))
    model = model.to(torch.device("cuda", 0))
    data = {"Coordinates": box["Coordinates"][:w], "InternalEnergy": box["InternalEnergy"][:w]}
    a = rollout(model, data, md, 0.0, md["dt"], md["box_size"], window_size=w, num_neighbors=k, n_steps=3)
    b = rollout_slab(model, data, md, 0.0, md["dt"], md["box_size"], window_size=w, num_neighbors=k, n_steps=3)
    d = (a["Coordinates"] - b["Coordinates"]).abs()
    d = torch.minimum(d, md["box_size"] - d)
    assert float(d.max()) < 1e-5 * md["box_size"]
    assert rel_l2(b["InternalEnergy"].cpu(), a["InternalEnergy"].cpu()) < 1e-5
