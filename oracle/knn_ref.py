"""Oracle: periodic-box k-NN, restated.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference call site: data_utils.py:148-152
    extended_positions, mapping = extend_positions_torch(recent_position, box_size)
    edge_index = knn(extended_positions, recent_position, num_neighbors)   # torch_cluster 1.6.3

`torch_cluster` is an un-vendored dependency (pinned in setup_env.sh:13) that cannot be
installed here, so for this call **parity is unpinned**; what is restated is its published
contract: for each query row of `y`, the k rows of `x` with the smallest Euclidean distance,
in ascending distance order, returned as [2, N*k] (row 0 = query index, row 1 = x index).

Canonical specification used on both sides:
  ext[s*N + j] = fl32(pos[j] + shift_s),  shift_s = B * (s//9 - 1, (s//3)%3 - 1, s%3 - 1)
                                           (torch.cartesian_prod order, data_utils.py:24-32)
  d2(i, c)     = fl32(fl32(fl32(dx*dx) + fl32(dy*dy)) + fl32(dz*dz)),  d = fl32(ext[c] - pos[i])
                 (no FMA contraction: the x86-64 CPU wheel the reference executes)
  neighbours   = the k smallest candidates under the TOTAL order (d2, c) ascending.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def shift_table(box: float) -> np.ndarray:
    """27 periodic shifts in torch.cartesian_prod([-B,0,B]^3) order (x slowest). data_utils.py:23-24"""
    v = np.array([-box, 0.0, box], dtype=np.float32)
    s = np.arange(27)
    return np.stack([v[s // 9], v[(s // 3) % 3], v[s % 3]], axis=1).astype(np.float32)


def extend_positions(pos: np.ndarray, box: float):
    """data_utils.py:9-33 in numpy: [27N,3] float32 ghosts and the [27N] mapping to originals."""
    pos = np.asarray(pos, dtype=np.float32)
    n = pos.shape[0]
    sh = shift_table(box)
    ext = (pos[None, :, :] + sh[:, None, :]).astype(np.float32).reshape(27 * n, 3)
    mapping = np.tile(np.arange(n, dtype=np.int64), 27)
    return ext, mapping


def knn_brute(pos: np.ndarray, box: float, k: int, block: int = 256) -> np.ndarray:
    """Exhaustive search over all 27N candidates. Returns ext indices [N,k] int64, rows sorted by (d2,c)."""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    n = pos.shape[0]
    ext, _ = extend_positions(pos, box)
    assert ext.shape[0] >= k, "27*N must be >= k"
    out = np.empty((n, k), dtype=np.int64)
    cidx = np.arange(ext.shape[0], dtype=np.int64)
    for a in range(0, n, block):
        q = pos[a:a + block]
        d = (ext[None, :, :] - q[:, None, :]).astype(np.float32)            # fl(ext - pos)
        sq = (d * d).astype(np.float32)                                     # fl(d*d) per axis
        d2 = ((sq[..., 0] + sq[..., 1]).astype(np.float32) + sq[..., 2]).astype(np.float32)
        for r in range(q.shape[0]):
            # candidates that can be in the top-k: everything <= the k-th smallest distance
            kth = np.partition(d2[r], k - 1)[k - 1]
            cand = cidx[d2[r] <= kth]
            order = np.lexsort((cand, d2[r][cand]))                         # primary d2, secondary c
            out[a + r] = cand[order[:k]]
    return out


# ---- C restatement (KD-tree over the 27N ghosts, like nanoflann inside torch_cluster) -------
_lib = None


def build_c(force: bool = False) -> str:
    so = os.path.join(_HERE, "libknn_oracle.so")
    src = os.path.join(_HERE, "knn_kdtree.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src, "-lm"])
    return so


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c())
        _lib.knn_oracle_kdtree.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_float,
                                           ctypes.c_int, ctypes.c_void_p]
        _lib.knn_oracle_kdtree.restype = ctypes.c_int
        _lib.knn_oracle_brute.argtypes = _lib.knn_oracle_kdtree.argtypes
        _lib.knn_oracle_brute.restype = ctypes.c_int
    return _lib


def knn_kdtree(pos: np.ndarray, box: float, k: int) -> np.ndarray:
    """Same specification as `knn_brute`, computed by the C KD-tree (single thread)."""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    out = np.empty((pos.shape[0], k), dtype=np.int64)
    rc = _load().knn_oracle_kdtree(pos.ctypes.data, pos.shape[0], np.float32(box), k, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"knn_oracle_kdtree failed rc={rc}")
    return out


def knn_brute_c(pos: np.ndarray, box: float, k: int) -> np.ndarray:
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    out = np.empty((pos.shape[0], k), dtype=np.int64)
    rc = _load().knn_oracle_brute(pos.ctypes.data, pos.shape[0], np.float32(box), k, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"knn_oracle_brute failed rc={rc}")
    return out


def edge_index_from_ext(ext_idx: np.ndarray, n: int) -> np.ndarray:
    """data_utils.py:150-152: swap to [sender; receiver] and map ghosts back (c mod N)."""
    k = ext_idx.shape[1]
    senders = (ext_idx % n).reshape(-1)
    receivers = np.repeat(np.arange(n, dtype=np.int64), k)
    return np.stack([senders, receivers], axis=0)
