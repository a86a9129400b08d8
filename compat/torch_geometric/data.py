"""`torch_geometric.data.Data` / `Batch` (data_utils.py:218-227, train.py:247) -> graph.py containers."""
from cosmology_gnn_simulation_b200.graph import Batch, Data  # noqa: F401
