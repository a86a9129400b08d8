"""The literal drop-in: the reference's own scripts, UNMODIFIED, on this repo's modules (SURVEY section 4 "Drop-in", 8(f)3).

`PYTHONPATH=compat:<repo>` makes `graph_network` / `data_utils` resolve to the B200 implementation and supplies the
stand-ins for the third-party packages this image lacks (torch_geometric containers / DataLoader, an npz-backed h5py,
a no-op matplotlib); the scripts themselves are executed from where the reference lies -- /root/reference in the build
container, baseline/_ref/ (staged by tools/stage_reference.sh, git-ignored) on the GPU box -- and are never edited.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

COMPAT = os.path.join(ROOT, "compat")


def _reference_dir():
    for d in (os.environ.get("CGNN_REFERENCE_DIR"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if d and os.path.exists(os.path.join(d, "train.py")):
            return d
    return None


def _run(script, args, env_extra=None, check=True):
    ref = _reference_dir()
    env = dict(os.environ)
    # -P: the script's own directory is not put in front of sys.path, so compat/ wins for `graph_network` / `data_utils`
    # while `config`, `dataloader`, `validation` still come from the reference directory further down the path
    env["PYTHONPATH"] = os.pathsep.join([COMPAT, ROOT, ref, env.get("PYTHONPATH", "")])
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, "-P", os.path.join(ref, script)] + args, env=env, cwd=ref, capture_output=True, text=True, timeout=900)
    if check:
        assert r.returncode == 0, f"{script} failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    return r


def _dataset(tmp_path, n=600, frames=8):
    """Two tiny simulations in the layout dataloader.py:41-51 reads, plus the metadata JSON of generate_metadata.py:32-43."""
    sys.path.insert(0, COMPAT)
    try:
        import h5py
        from cosmology_gnn_simulation_b200 import synthetic
        md = None
        for split, seed in (("train", 0), ("val", 1)):
            os.makedirs(tmp_path / split, exist_ok=True)
            box = synthetic.make_box(n, "uniform", window=frames - 1, seed=seed)
            md = md or box["metadata"]
            with h5py.File(str(tmp_path / split / "sim0.hdf5"), "w") as f:
                f.create_dataset("Coordinates", data=box["Coordinates"].numpy())
                f.create_dataset("InternalEnergy", data=box["InternalEnergy"].numpy()[..., 0])
    finally:
        sys.path.remove(COMPAT)
    with open(tmp_path / "metadata.json", "w") as f:
        json.dump(md, f)
    return md


needs_reference = pytest.mark.skipif(_reference_dir() is None, reason="the reference scripts are not on this machine")


@needs_reference
@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check of everything around the hot path")
def test_train_script_reaches_the_hot_path_without_a_gpu(tmp_path):
    """No CUDA device: train.py must get through its imports, config, SequenceDataset, DataLoader and model construction on
    the shims and stop exactly where the hot path starts -- `preprocess` refusing to run without a GPU (no CPU fallback)."""
    _dataset(tmp_path)
    r = _run("train.py", ["--train_dir", str(tmp_path / "train"), "--val_dir", str(tmp_path / "val"), "--metadata_path",
                          str(tmp_path / "metadata.json"), "--output_dir", str(tmp_path / "out"), "--num_epochs", "1",
                          "--device", "cpu"], check=False)
    assert r.returncode != 0
    assert "Initializing dataset with 1 simulation file(s)" in r.stdout
    assert "Learning rate will decay" in r.stdout                      # model and optimizer were built (train.py:165-190)
    assert "cgnn preprocess needs a CUDA device" in r.stderr, r.stderr[-2000:]


@needs_reference
@pytest.mark.gpu
def test_reference_scripts_run_unmodified(tmp_path):
    """train.py (one epoch + validation + checkpoints + plots), one_step_test.py and render_rollout.py, as they are."""
    _dataset(tmp_path)
    out = tmp_path / "out"
    common = ["--metadata_path", str(tmp_path / "metadata.json")]
    r = _run("train.py", ["--train_dir", str(tmp_path / "train"), "--val_dir", str(tmp_path / "val"), "--output_dir", str(out),
                          "--num_epochs", "2", "--save_every", "1", "--momentum_loss_weight", "0.1", "--noise_std", "0.0003",
                          "--batch_size", "2"] + common)
    assert "Training complete" in r.stdout
    hist = json.load(open(out / "training_history.json"))
    assert len(hist["train_loss"]) == 2 and all(np.isfinite(hist["train_loss"])) and all(np.isfinite(hist["val_loss"]))
    for name in ("model_best.pth", "model_final.pth", "model_epoch_1.pth", os.path.join("plots", "losses_final.png")):
        assert os.path.exists(out / name), name
    sd = torch.load(out / "model_final.pth", map_location="cpu")
    assert "processor.9.edge_model.0.0.weight" in sd and tuple(sd["encoder.node_model.0.0.weight"].shape) == (128, 17)

    r = _run("one_step_test.py", ["--model_path", str(out / "model_final.pth"), "--test_data", str(tmp_path / "val" / "sim0.hdf5"),
                                  "--num_timesteps", "2"] + common)
    assert "ONE-STEP VALIDATION RESULTS" in r.stdout and "nan" not in r.stdout.lower()

    # the same checkpoint through the tensor-core kernels (CGNN_PRECISION / CGNN_MESSAGE: defaults of the extra keyword arguments)
    roll = tmp_path / "rollout"
    r = _run("render_rollout.py", ["--model_path", str(out / "model_final.pth"), "--test_data", str(tmp_path / "val" / "sim0.hdf5"),
                                   "--output_dir", str(roll)] + common, env_extra={"CGNN_PRECISION": "bf16x3"})
    assert "Evaluation complete" in r.stdout
    coords = np.load(roll / "rollout_coordinates.npy")
    assert coords.shape == (8, 600, 3) and np.isfinite(coords).all()
    assert os.path.exists(roll / "errors.png") and os.path.exists(roll / "rollout_summary.txt")
