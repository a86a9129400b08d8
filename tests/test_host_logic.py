"""CPU: host-side logic of the boundary modules (no kernels)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from cosmology_gnn_simulation_b200.graph import Batch, Data
from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode
from cosmology_gnn_simulation_b200 import data_utils


def test_state_dict_keys_match_reference():
    for name in ["model_tiny", "model_deepmlp"]:
        g = load_golden(name)
        L, H, nh, M, out = [int(v) for v in g["cfg"]]
        m = EncodeProcessDecode(L, H, nh, M, out)
        assert list(m.state_dict().keys()) == [str(k) for k in g["sd_keys"]]
        # optimizer can be built before the first forward, like train.py:173-183
        torch.optim.Adam(m.parameters(), lr=1e-3)
        sd = {str(k): torch.from_numpy(g["sd/" + str(k)]) for k in g["sd_keys"]}
        m.load_state_dict(sd)
        for k, v in m.state_dict().items():
            assert np.array_equal(v.numpy(), g["sd/" + k])


def test_lazy_first_layers_materialise_with_reference_shapes():
    m = EncodeProcessDecode(128, 128, 2, 2, 3)
    m._materialize_all(17, 4, torch.empty(1))
    sd = m.state_dict()
    assert tuple(sd["encoder.node_model.0.0.weight"].shape) == (128, 17)
    assert tuple(sd["encoder.edge_model.0.0.weight"].shape) == (128, 4)
    assert tuple(sd["processor.1.edge_model.0.0.weight"].shape) == (128, 384)
    assert tuple(sd["processor.1.node_model.0.0.weight"].shape) == (128, 256)
    assert tuple(sd["decoder_acc.4.weight"].shape) == (3, 128)
    assert tuple(sd["decoder_temp_rate.4.weight"].shape) == (1, 128)


def test_batch_offsets_and_ptr():
    def g(n, k):
        return Data(x=torch.randn(n, 17), edge_index=torch.stack([torch.randint(0, n, (n * k,)),
                    torch.arange(n).repeat_interleave(k)]), edge_attr=torch.randn(n * k, 4),
                    y_acc=torch.randn(n, 3), y_temp_rate=torch.randn(n, 1), pos=torch.rand(n, 3),
                    dt=torch.tensor([0.01]), box_size=torch.tensor([1.0]),
                    _cgnn_senders=None, _cgnn_k=k)
    a, b = g(5, 2), g(7, 2)
    a._cgnn_senders = a.edge_index[0].int()
    b._cgnn_senders = b.edge_index[0].int()
    batch = Batch.from_data_list([a, b])
    assert batch.num_graphs == 2 and batch.x.shape[0] == 12
    assert batch.ptr.tolist() == [0, 5, 12]
    assert torch.equal(batch.batch, torch.tensor([0] * 5 + [1] * 7))
    assert torch.equal(batch.edge_index[1], torch.arange(12).repeat_interleave(2))
    assert torch.equal(batch.edge_index[0][10:], b.edge_index[0] + 5)
    assert torch.equal(batch._cgnn_senders.long(), batch.edge_index[0])
    assert batch.dt.shape == (2,)


def test_extend_positions_matches_reference_order():
    pos = torch.rand(5, 3)
    ext, mapping = data_utils.extend_positions_torch(pos, 2.0)
    assert ext.shape == (135, 3) and mapping.shape == (135,)
    assert torch.equal(ext[13 * 5:14 * 5], pos)                      # zero shift is block 13
    assert torch.equal(ext[0:5], pos + torch.tensor([-2.0, -2.0, -2.0]))
    assert torch.equal(ext[5:10], pos + torch.tensor([-2.0, -2.0, 0.0]))   # z fastest
    assert torch.equal(mapping, torch.arange(5).repeat(27))


def test_noise_generators_match_oracle_and_rng_consumption():
    """The oracle's noise path is pinned by the reference fixture `pre_clustered_noise`
    (tests/test_oracle_golden.py); the product generators must draw the same numbers."""
    from oracle import preprocess_ref
    g = load_golden("pre_clustered_noise")
    coords = torch.from_numpy(g["coords"])[:5].permute(1, 0, 2)
    energy = torch.from_numpy(g["energy"])[:5].permute(1, 0, 2)
    for std in (0.0, 3e-4):
        torch.manual_seed(3)
        a = data_utils.generate_position_noise(coords, std, 1.0, 0.01)
        b = data_utils.generate_temperature_noise(energy, std, torch.tensor(0.7), 0.01)
        after = torch.rand(2)
        torch.manual_seed(3)
        vel = preprocess_ref._min_image_(coords[:, 1:] - coords[:, :-1], 1.0) / 0.01
        a_ref = preprocess_ref._random_walk_noise(vel, std, 0.01)
        b_ref = preprocess_ref._random_walk_noise((energy[:, 1:] - energy[:, :-1]) / 0.01, std * torch.tensor(0.7), 0.01)
        assert torch.equal(a, a_ref) and torch.equal(b, b_ref)
        assert torch.equal(after, torch.rand(2))          # draws happen even for std == 0
        assert a.shape == coords.shape and float(a[:, 0].abs().max()) == 0.0


def test_preprocess_without_gpu_raises():
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    g = load_golden("pre_uniform")
    with pytest.raises(RuntimeError, match="CUDA"):
        data_utils.preprocess(torch.from_numpy(g["coords"])[:5], torch.from_numpy(g["energy"])[:5],
                              {"temp_rate_std": 1.0}, dt=0.01, box_size=1.0)


def test_bench_work_model_matches_survey_table():
    """bench.py's algorithmic FLOP / HBM-byte model reproduces SURVEY §8d's table rows (config 2: 2.83 TFLOP and 14.9 GB per
    application with FP32 latents; config 3: 352.8 TFLOP)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    w = bench.phase_work(32768, 16, 128, 10, "edge")
    flop = w["forward"][0] + w["backward"][0]
    assert abs(flop - bench.flops_application(32768, 16, 128, 10, "edge")) < 1e-6 * flop
    assert abs(flop / 1e12 - 2.83) < 0.01
    assert abs((w["forward"][1] + w["backward"][1]) / 1e9 - 14.9) < 0.1
    assert abs(bench.flops_application(128 ** 3, 32, 128, 10, "edge") / 1e12 - 352.8) < 0.5
    s = bench.phase_work(32768, 16, 128, 10, "sender")
    assert abs((s["forward"][0] + s["backward"][0]) - bench.flops_application(32768, 16, 128, 10, "sender")) < 1e-6 * flop


def test_edge_stream_schedule_is_valid_and_minimal():
    """ckpt_plan.schedule: every step's backward sees e^t, buffers stay within the budget, nothing is recomputed when all
    copies fit, and the recompute count matches the dynamic programme's optimum for the sizes the configs use."""
    from cosmology_gnn_simulation_b200 import ckpt_plan
    for M in (1, 2, 5, 10, 15):
        for nbuf in (1, 2, 3, 4, M, M + 3):
            acts = ckpt_plan.schedule(M, nbuf)
            held, done, first = {}, [], ckpt_plan.split_forward(acts)
            for i, a in enumerate(acts):
                if a[0] == "enc":
                    held[a[1]] = 0
                elif a[0] == "adv":
                    _, t, src, dst = a
                    assert held[src] == t and dst < nbuf
                    held[dst] = t + 1
                else:
                    _, t, b = a
                    assert i >= first and held.pop(b) == t
                    done.append(t)
            assert done == list(range(M - 1, -1, -1))
            # the forward sweep visits every step below M-1 exactly once, in order
            assert [a[1] for a in acts[:first] if a[0] == "adv"] == list(range(M - 1))
            if nbuf >= M:
                assert ckpt_plan.recomputed_phases(acts) == 0.0
    assert ckpt_plan.recomputed_phases(ckpt_plan.schedule(10, 2)) == 12.5
    assert ckpt_plan.recomputed_phases(ckpt_plan.schedule(10, 3)) == 7.0


def test_model_plan_is_built_on_cpu_and_the_kernels_refuse_cpu_tensors():
    """Everything above the C ABI (plan, parameter groups, precision of the edge-stream backward) runs without a GPU; the first
    kernel entry point then refuses the CPU tensors -- there is no CPU path."""
    import pytest
    from cosmology_gnn_simulation_b200.graph import Data
    from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode, GRAD16_MIN_ROWS, _Plan
    n, k = 64, 4
    g = Data(x=torch.randn(n, 17), edge_index=torch.stack([torch.randint(0, n, (n * k,)), torch.arange(n).repeat_interleave(k)]),
             edge_attr=torch.randn(n * k, 4))
    for precision in ("fp32", "bf16x3"):
        model = EncodeProcessDecode(128, 128, 2, 2, 3, message="edge", precision=precision)
        with pytest.raises((RuntimeError, TypeError, ValueError)) as err:
            model(g)
        assert "CUDA" in str(err.value) or "cuda" in str(err.value), err.value
    plan = _Plan(2, "edge", "bf16x3", k, None, None, [], [], None, None, [], 0, grad_stream="bf16")
    assert plan.edge_bwd_precision(GRAD16_MIN_ROWS) == "bf16x3g" and plan.edge_bwd_precision(GRAD16_MIN_ROWS - 1) == "bf16x3"
    assert _Plan(2, "edge", "bf16x3", k, None, None, [], [], None, None, [], 0, grad_stream="fp32").edge_bwd_precision(1 << 30) == "bf16x3"
    assert _Plan(2, "edge", "fp32", k, None, None, [], [], None, None, [], 0, grad_stream="bf16").edge_bwd_precision(1 << 30) == "fp32"
