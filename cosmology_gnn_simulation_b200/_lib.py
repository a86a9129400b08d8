"""ctypes binding of libcgnn.so (include/cgnn.h).  No CPU fallback: if the library is missing the
first kernel call raises, loudly."""
from __future__ import annotations

import contextlib
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcgnn.so")
MAX_LAYERS = 4

# "bf16x3g" (CGNN_PREC_BF16X3_G16): bf16x3 whose long-stream backward keeps dY / G2 / G1 as bfloat16; chosen per call by the
# model for the edge MLPs (graph_network.py, grad_stream), not a model-level precision
PREC = {"fp32": 0, "bf16x3": 1, "bf16": 2, "bf16x3g": 3}
MSG = {"sender": 0, "edge": 1}
DISP = {"raw": 0, "min_image": 1}


class CgnnMlp(Structure):
    _fields_ = [("n_layers", c_int32), ("in_dim", c_int32), ("hidden", c_int32), ("out_dim", c_int32),
                ("W", c_void_p * MAX_LAYERS), ("b", c_void_p * MAX_LAYERS),
                ("ln_gamma", c_void_p), ("ln_beta", c_void_p), ("ln_dim", c_int32)]


class CgnnMlpGrad(Structure):
    _fields_ = [("W", c_void_p * MAX_LAYERS), ("b", c_void_p * MAX_LAYERS),
                ("ln_gamma", c_void_p), ("ln_beta", c_void_p)]


# name -> (restype, argtypes); exactly the declarations of include/cgnn.h
SIGNATURES = {
    "cgnn_last_error": (c_char_p, []),
    "cgnn_version": (c_char_p, []),
    "cgnn_launch_count": (c_int64, []),
    "cgnn_debug_stamps": (None, [c_void_p, c_int32, c_int32]),
    "cgnn_knn_workspace_bytes": (c_int64, [c_int64]),
    "cgnn_knn_periodic": (c_int, [c_void_p, c_int64, c_float, c_int32, c_void_p, c_void_p, c_int64, c_void_p]),
    "cgnn_knn_periodic_range": (c_int, [c_void_p, c_int64, c_float, c_int32, c_int64, c_int64, c_void_p, c_void_p, c_int64,
                                        c_void_p]),
    "cgnn_edge_features_range": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_float, c_int32, c_int64, c_int64, c_void_p,
                                         c_void_p, c_void_p, c_void_p]),
    "cgnn_edge_features": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_float, c_int32, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
    "cgnn_preprocess_features": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_float, c_float,
                                         POINTER(c_float), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cgnn_csr_transpose_workspace_bytes": (c_int64, [c_int64, c_int64]),
    "cgnn_csr_transpose": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "cgnn_edge_index_to_senders": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "cgnn_mlp_rows_workspace_bytes": (c_int64, [POINTER(CgnnMlp), c_int64, c_int32, c_int32]),
    "cgnn_mlp_rows_fwd": (c_int, [POINTER(CgnnMlp), c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    "cgnn_mlp_bwd_workspace_bytes": (c_int64, [POINTER(CgnnMlp)]),
    "cgnn_mlp_rows_bwd": (c_int, [POINTER(CgnnMlp), POINTER(CgnnMlpGrad), c_void_p, c_int64, c_void_p, c_void_p,
                                  c_void_p, c_int64, c_int32, c_void_p]),
    "cgnn_mp_edge_fwd_workspace_bytes": (c_int64, [POINTER(CgnnMlp), c_int64, c_int32]),
    "cgnn_mp_edge_fwd": (c_int, [POINTER(CgnnMlp), c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32, c_void_p,
                                 c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    "cgnn_aggregate_senders": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "cgnn_halo_pack": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "cgnn_halo_unpack_add": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "cgnn_mp_node_fwd_workspace_bytes": (c_int64, [POINTER(CgnnMlp), c_int64, c_int32]),
    "cgnn_mp_node_fwd": (c_int, [POINTER(CgnnMlp), c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int32,
                                 c_void_p]),
    "cgnn_mp_node_bwd": (c_int, [POINTER(CgnnMlp), POINTER(CgnnMlpGrad), c_void_p, c_void_p, c_void_p, c_int64,
                                 c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    "cgnn_mp_bwd_workspace_bytes": (c_int64, [POINTER(CgnnMlp), c_int64, c_int64, c_int32, c_int32]),
    "cgnn_mp_edge_bwd": (c_int, [POINTER(CgnnMlp), POINTER(CgnnMlpGrad), c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_int64, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_int64, c_int32, c_void_p]),
    "cgnn_scatter_to_senders": (c_int, [c_void_p, c_int32, c_void_p, c_void_p, c_int64, c_int32, c_int32,
                                        c_void_p, c_void_p]),
    "cgnn_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float, c_float,
                               c_int32, c_float, c_void_p]),
    "cgnn_adam_step_masked": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
                                      c_float, c_int32, c_float, c_void_p]),
    "cgnn_loss_workspace_bytes": (c_int64, [c_int64, c_int32]),
    "cgnn_loss_fwd_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32,
                                  c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int64, c_void_p]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """The loaded library.  Raises if it has not been built (python __graft_entry__.py / make -C csrc)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                "(there is no CPU fallback for the cgnn kernels)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().cgnn_last_error().decode(errors="replace")
        raise RuntimeError(f"libcgnn {what} failed (status {rc}): {msg}")


def ptr(t):
    """Device pointer of a tensor (or NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, name: str, dtype=None):
    if not t.is_cuda:
        raise RuntimeError(f"cgnn: `{name}` must be a CUDA tensor (got device {t.device}); "
                           "this package has no CPU path")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"cgnn: `{name}` must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"cgnn: `{name}` must be contiguous")
    return t


class Workspace:
    """Grow-only scratch buffers, one per (device, tag); the C ABI never allocates.

    A CUDA-graph capture records the raw addresses of the buffers it was handed, so a capture must own its buffers:
    `with workspace.scope() as held:` switches to a private set for the duration of the block (graphed.py keeps
    `held` alive next to the captured graph); buffers of the shared set are then free to grow or move without
    touching a captured step."""

    def __init__(self):
        self._bufs = {}

    def get(self, device, tag: str, nbytes: int) -> torch.Tensor:
        key = (str(device), tag)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            if buf is not None and torch.cuda.is_current_stream_capturing():
                raise RuntimeError(f"cgnn: workspace '{tag}' would have to grow inside a CUDA-graph capture "
                                   "(warm the step up at its final size before capturing)")
            self._bufs.pop(key, None)
            buf = None                       # the old buffer goes before the new one comes (they can be tens of GiB)
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf

    def tagged_bytes(self, device, prefix: str) -> int:
        """Bytes currently held under tags that start with `prefix` (the planner counts them as available)."""
        return sum(b.numel() for (dev, tag), b in self._bufs.items() if dev == str(device) and tag.startswith(prefix) and b is not None)

    @contextlib.contextmanager
    def scope(self):
        outer, self._bufs = self._bufs, {}
        try:
            yield self._bufs
        finally:
            self._bufs = outer


workspace = Workspace()


def launch_count() -> int:
    return int(lib().cgnn_launch_count())
