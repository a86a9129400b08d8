"""The oracle against the fixtures produced by the reference's own files (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import golden_metadata, load_golden, rel_l2
from oracle import knn_ref, model_ref, preprocess_ref


@pytest.mark.parametrize("name", ["model_tiny", "model_small", "model_deepmlp"])
def test_model_forward_backward_matches_reference(name):
    g = load_golden(name)
    L, H, nh, M, out = [int(v) for v in g["cfg"]]
    params = {str(k): torch.from_numpy(g["sd/" + str(k)]).clone().requires_grad_(True) for k in g["sd_keys"]}
    # key set is exactly the reference state_dict
    shapes = model_ref.param_shapes(L, H, nh, M, out, g["x"].shape[1], g["edge_attr"].shape[1])
    assert list(shapes.keys()) == [str(k) for k in g["sd_keys"]]
    for k, s in shapes.items():
        assert tuple(params[k].shape) == tuple(s)
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    ea = torch.from_numpy(g["edge_attr"]).requires_grad_(True)
    ei = torch.from_numpy(g["edge_index"])
    pred = model_ref.forward(params, x, ei, ea, nh, M, message="sender")
    assert rel_l2(pred["acceleration"].detach(), g["acceleration"]) < 1e-5
    assert rel_l2(pred["temp_rate"].detach(), g["temp_rate"]) < 1e-5
    ls = model_ref.loss(pred["acceleration"], pred["temp_rate"], torch.from_numpy(g["y_acc"]),
                        torch.from_numpy(g["y_temp_rate"]), float(g["dt"]), w_mom=0.1)
    assert abs(ls["loss"].item() - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    assert abs(ls["momentum_loss"].item() - float(g["mom_loss"])) <= 1e-5 * abs(float(g["mom_loss"])) + 1e-12
    ls["loss"].backward()
    none = set(str(k) for k in g["grad_none"])
    for k, p in params.items():
        if k in none:
            # the reference leaves the whole edge stream without gradient
            assert "edge_model" in k
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
        else:
            assert rel_l2(p.grad, g["grad/" + k]) < 5e-5, k
    assert rel_l2(x.grad, g["grad_x"]) < 5e-5
    assert bool(g["grad_edge_attr_is_none"]) and ea.grad is None


@pytest.mark.parametrize("name", ["pre_uniform", "pre_clustered_noise"])
def test_preprocess_matches_reference(name):
    g = load_golden(name)
    md = golden_metadata(g)
    coords, energy = torch.from_numpy(g["coords"]), torch.from_numpy(g["energy"])
    torch.manual_seed(int(g["seed"]))
    out = preprocess_ref.preprocess(coords[:5], energy[:5], md, coords[5:6].clone(), energy[5:6].clone(),
                                    noise_std=float(g["noise_std"]), num_neighbors=int(g["k"]),
                                    dt=md["dt"], box_size=md["box_size"])
    assert np.array_equal(torch.rand(4).numpy(), g["rng_after"])   # same RNG consumption
    assert np.array_equal(out["edge_index"].numpy(), g["edge_index"])
    for key in ["x", "edge_attr", "y_acc", "y_temp_rate", "pos", "dt", "box_size"]:
        assert np.array_equal(out[key].numpy(), g[key]), key


def test_graph_layout_properties():
    g = load_golden("pre_uniform")
    ei, k = g["edge_index"], int(g["k"])
    n = g["pos"].shape[0]
    assert np.array_equal(ei[1], np.repeat(np.arange(n), k))       # receiver-sorted, fixed in-degree
    assert np.array_equal(ei[0].reshape(n, k)[:, 0], np.arange(n))  # rank 0 is the particle itself
    # raw (not minimum-image) displacements: some edges are longer than half the box
    assert (np.abs(g["edge_attr"][:, :3]) > 0.5).any()


@pytest.mark.parametrize("kind,n,k", [("uniform", 500, 16), ("clustered", 700, 32), ("lattice", 343, 8)])
def test_knn_c_oracles_agree_with_numpy(kind, n, k):
    from cosmology_gnn_simulation_b200 import synthetic
    pos = synthetic.positions(n, kind, 1.0, seed=3)
    a = knn_ref.knn_brute(pos, 1.0, k)
    assert np.array_equal(a, knn_ref.knn_brute_c(pos, 1.0, k))
    assert np.array_equal(a, knn_ref.knn_kdtree(pos, 1.0, k))


def test_knn_against_scipy_periodic_kdtree():
    """Independent cross-check of the restated torch_cluster contract (fp64 tree, so compare as
    distance multisets: ties may pick a different image/index)."""
    from scipy.spatial import cKDTree
    from cosmology_gnn_simulation_b200 import synthetic
    n, k = 2000, 16
    pos = synthetic.positions(n, "uniform", 1.0, seed=5)
    ext_idx = knn_ref.knn_kdtree(pos, 1.0, k)
    d_sc, i_sc = cKDTree(pos.astype(np.float64), boxsize=1.0).query(pos.astype(np.float64), k=k)
    assert np.array_equal(np.sort(ext_idx % n, axis=1), np.sort(i_sc, axis=1))


def test_knn_edge_cases():
    # fewer real particles than k: ghosts of the same particle fill the list
    pos = np.array([[0.1, 0.2, 0.3], [0.9, 0.9, 0.9]], dtype=np.float32)
    a = knn_ref.knn_brute(pos, 1.0, 16)
    assert np.array_equal(a, knn_ref.knn_kdtree(pos, 1.0, 16))
    assert a[0, 0] == 13 * 2 + 0 and a[1, 0] == 13 * 2 + 1
    # coincident particles and a particle exactly on the upper face
    pos = np.array([[0.5, 0.5, 0.5], [0.5, 0.5, 0.5], [1.0, 0.0, 1.0], [0.0, 0.0, 0.0]], dtype=np.float32)
    a = knn_ref.knn_brute(pos, 1.0, 4)
    assert np.array_equal(a, knn_ref.knn_kdtree(pos, 1.0, 4))
    assert a[1, 0] == 13 * 4 + 0          # tie at d2 = 0 resolved by the lower extended index
