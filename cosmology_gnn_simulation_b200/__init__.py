"""B200-native (sm_100a) Interaction-Network hot path of mattpan-peregrinus/Cosmology_GNN_Simulation.

    from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode   # graph_network.py
    from cosmology_gnn_simulation_b200.data_utils import preprocess              # data_utils.py

Everything numerical runs in `libcgnn.so` (C ABI: include/cgnn.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
