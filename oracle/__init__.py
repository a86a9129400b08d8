"""CPU oracle for the Interaction-Network hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement of the reference algorithm
(`/root/reference/graph_network.py`, `/root/reference/data_utils.py`,
`/root/reference/train.py:107-118,255-260`).  It is the *checker*, never the
product: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it.  The product package
`cosmology_gnn_simulation_b200` never imports anything from here and fails
loudly when its CUDA library is missing.

Parity status
-------------
* model / loss / preprocess feature arithmetic: PINNED against the reference's
  own `graph_network.py` and `data_utils.py` executed in the build container
  with stubs for the three missing third-party symbols (`oracle/make_golden.py`
  -> `tests/golden/*.npz`).
* `torch_cluster.knn` (torch-cluster 1.6.3, un-vendored) and PyG's
  `MessagePassing.propagate` (torch-geometric 2.6.1, un-vendored) cannot be
  executed here: for those two call sites **parity is unpinned**; their
  published semantics are restated (`knn_ref.py`, `model_ref.py`) and
  cross-checked against `scipy.spatial.cKDTree(boxsize=...)` and `index_add_`.
"""
