"""Import shim: put this directory first on PYTHONPATH and the reference's scripts
(`from graph_network import EncodeProcessDecode`, train.py:16, one_step_test.py:9, render_rollout.py:10)
resolve to the B200 implementation."""
from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode, build_mlp  # noqa: F401
