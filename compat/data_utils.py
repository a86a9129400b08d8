"""Import shim: `from data_utils import preprocess` (train.py:15, validation.py:3, one_step_test.py:10,
render_rollout.py:11) resolves to the B200 implementation."""
from cosmology_gnn_simulation_b200.data_utils import (  # noqa: F401
    extend_positions_torch, generate_position_noise, generate_temperature_noise, preprocess)
