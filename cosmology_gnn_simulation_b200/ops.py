"""Thin torch-tensor wrappers over the C ABI (include/cgnn.h).  Tensors in, tensors out; every call
runs on the current CUDA stream of the tensors' device.  No CPU path."""
from __future__ import annotations

from ctypes import byref, c_void_p
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import CgnnMlp, CgnnMlpGrad, PREC, DISP, check, lib, ptr, require_cuda, stream_ptr, workspace


# ------------------------------------------------------------------------------------------------
# graph build (K1, K2)
# ------------------------------------------------------------------------------------------------
def knn_periodic(pos: torch.Tensor, box_size: float, k: int, query_range=None) -> torch.Tensor:
    """Replaces `extend_positions_torch` + `torch_cluster.knn` (data_utils.py:148-149).
    pos [N,3] fp32 CUDA -> ext index [N,k] int32 (c = shift*N + j), rows ascending in (d2, c).
    `query_range=(q0, nq)`: answer only the queries q0 <= i < q0+nq (a rank's slab) -> [nq,k]."""
    require_cuda(pos, "pos", torch.float32)
    n = pos.shape[0]
    if pos.dim() != 2 or pos.shape[1] != 3:
        raise ValueError("pos must be [N,3]")
    q0, nq = (0, n) if query_range is None else (int(query_range[0]), int(query_range[1]))
    out = torch.empty((nq, k), dtype=torch.int32, device=pos.device)
    with torch.cuda.device(pos.device):
        nbytes = lib().cgnn_knn_workspace_bytes(n)
        ws = workspace.get(pos.device, "knn", nbytes)
        check(lib().cgnn_knn_periodic_range(ptr(pos), n, float(box_size), int(k), q0, nq, ptr(out), ptr(ws), ws.numel(),
                                            stream_ptr(pos.device)), "cgnn_knn_periodic")
    return out


def edge_features(pos: torch.Tensor, nbr_ext: torch.Tensor, box_size: float, disp: str = "raw",
                  want_edge_index: bool = True, q0: int = 0):
    """Replaces data_utils.py:150-164.  Returns (senders int32 [E], edge_index int64 [2,E] | None,
    edge_attr fp32 [E,4]).  `nbr_ext` may cover only the receivers q0 <= i < q0 + nbr_ext.shape[0]
    (senders / edge_index then hold global ids)."""
    require_cuda(pos, "pos", torch.float32)
    require_cuda(nbr_ext, "nbr_ext", torch.int32)
    nq, k = nbr_ext.shape
    e = nq * k
    senders = torch.empty(e, dtype=torch.int32, device=pos.device)
    edge_index = torch.empty((2, e), dtype=torch.int64, device=pos.device) if want_edge_index else None
    edge_attr = torch.empty((e, 4), dtype=torch.float32, device=pos.device)
    with torch.cuda.device(pos.device):
        check(lib().cgnn_edge_features_range(ptr(pos), ptr(nbr_ext), pos.shape[0], k, float(box_size), DISP[disp],
                                             int(q0), nq, ptr(senders), ptr(edge_index), ptr(edge_attr),
                                             stream_ptr(pos.device)), "cgnn_edge_features")
    return senders, edge_index, edge_attr


def csr_transpose(senders: torch.Tensor, n: int):
    """Sender-sorted transpose: (rowptr int32 [N+1], perm int32 [E])."""
    require_cuda(senders, "senders", torch.int32)
    e = senders.numel()
    rowptr = torch.empty(n + 1, dtype=torch.int32, device=senders.device)
    perm = torch.empty(e, dtype=torch.int32, device=senders.device)
    with torch.cuda.device(senders.device):
        nbytes = lib().cgnn_csr_transpose_workspace_bytes(n, e)
        ws = workspace.get(senders.device, "csr", nbytes)
        check(lib().cgnn_csr_transpose(ptr(senders), n, e, ptr(rowptr), ptr(perm), ptr(ws), ws.numel(),
                                       stream_ptr(senders.device)), "cgnn_csr_transpose")
    return rowptr, perm


def senders_from_edge_index(edge_index: torch.Tensor, n: int) -> torch.Tensor:
    """Validates the receiver-sorted fixed-in-degree layout and returns int32 senders.
    Raises ValueError for any other graph layout (one 4-byte D2H read)."""
    require_cuda(edge_index, "edge_index", torch.int64)
    if edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise ValueError("edge_index must be [2,E]")
    e = edge_index.shape[1]
    if n <= 0 or e % n != 0 or e == 0:
        raise ValueError(f"cgnn needs a fixed in-degree graph: E={e} is not a multiple of N={n}")
    k = e // n
    senders = torch.empty(e, dtype=torch.int32, device=edge_index.device)
    bad = torch.empty(1, dtype=torch.int32, device=edge_index.device)
    with torch.cuda.device(edge_index.device):
        check(lib().cgnn_edge_index_to_senders(ptr(edge_index), n, k, ptr(senders), ptr(bad),
                                               stream_ptr(edge_index.device)), "cgnn_edge_index_to_senders")
    if int(bad.item()) != 0:
        raise ValueError("cgnn needs the receiver-sorted k-NN layout produced by preprocess "
                         "(edge_index[1] == arange(N).repeat_interleave(k), senders in [0,N))")
    return senders


# (every MLP entry point shares ONE scratch buffer, tag "mlp": calls are ordered on the stream and none keeps its
#  workspace beyond its own kernels, so the buffer only has to be as large as the largest request)
# ------------------------------------------------------------------------------------------------
# MLP descriptors
# ------------------------------------------------------------------------------------------------
class MlpParams:
    """Flat view of one reference `build_mlp` (+LayerNorm): weights [out,in], biases, gamma/beta."""

    def __init__(self, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
                 gamma: Optional[torch.Tensor] = None, beta: Optional[torch.Tensor] = None):
        self.weights = list(weights)
        self.biases = list(biases)
        self.gamma, self.beta = gamma, beta
        self.ln_dim = 0                      # LayerNorm width when the outputs are zero-padded to 128 (0: all of out_dim)
        nl = len(self.weights)
        if not (1 <= nl <= _lib.MAX_LAYERS):
            raise ValueError(f"cgnn supports 1..{_lib.MAX_LAYERS} Linear layers per MLP, got {nl}")
        self.in_dim = self.weights[0].shape[1]
        self.out_dim = self.weights[-1].shape[0]
        self.hidden = self.weights[0].shape[0] if nl > 1 else self.out_dim
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            require_cuda(w, f"weight[{i}]", torch.float32)
            require_cuda(b, f"bias[{i}]", torch.float32)
            exp_in = self.in_dim if i == 0 else self.hidden
            exp_out = self.out_dim if i == nl - 1 else self.hidden
            if tuple(w.shape) != (exp_out, exp_in) or tuple(b.shape) != (exp_out,):
                raise ValueError(f"layer {i}: expected weight {(exp_out, exp_in)}, got {tuple(w.shape)}")
        if gamma is not None:
            require_cuda(gamma, "ln.weight", torch.float32)
            require_cuda(beta, "ln.bias", torch.float32)

    def tensors(self) -> List[torch.Tensor]:
        out = []
        for w, b in zip(self.weights, self.biases):
            out += [w, b]
        if self.gamma is not None:
            out += [self.gamma, self.beta]
        return out

    def c_struct(self) -> CgnnMlp:
        m = CgnnMlp()
        m.n_layers, m.in_dim, m.hidden, m.out_dim = len(self.weights), self.in_dim, self.hidden, self.out_dim
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            m.W[i] = w.data_ptr()
            m.b[i] = b.data_ptr()
        m.ln_gamma = None if self.gamma is None else self.gamma.data_ptr()
        m.ln_beta = None if self.beta is None else self.beta.data_ptr()
        m.ln_dim = self.ln_dim
        return m

    def new_grads(self):
        """(CgnnMlpGrad, [grad tensors in `tensors()` order])"""
        g = CgnnMlpGrad()
        outs = []
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            gw, gb = torch.empty_like(w), torch.empty_like(b)
            g.W[i], g.b[i] = gw.data_ptr(), gb.data_ptr()
            outs += [gw, gb]
        if self.gamma is not None:
            gg, gb = torch.empty_like(self.gamma), torch.empty_like(self.beta)
            g.ln_gamma, g.ln_beta = gg.data_ptr(), gb.data_ptr()
            outs += [gg, gb]
        return g, outs


TC_WIDTH = 128            # latent = hidden width of the tcgen05 tiles


class PaddedMlp(MlpParams):
    """Zero-padded copy of an MLP for the 128-wide tensor-core tiles: hidden layers and LayerNorm'd outputs become 128 wide
    (padded rows / columns / biases / gamma / beta are zero, the LayerNorm keeps its own width through `ln_dim`), a first
    layer that reads a concatenation of `in_blocks` latents gets each block at a 128-column offset.  With zero padding every
    padded activation is exactly zero, so the real columns carry exactly the unpadded network (README.md:59-62 allows latent
    sizes 64 / 128 / 256; 64 runs this way).  `unpad(grads)` cuts the parameter gradients back to the real shapes."""

    def __init__(self, mp: MlpParams, in_blocks: int, latent: int):
        t, nl = TC_WIDTH, len(mp.weights)
        self.orig, self.cuts = mp, []
        ws, bs = [], []
        for i, (w, b) in enumerate(zip(mp.weights, mp.biases)):
            o_real, i_real = w.shape
            o_pad = t if (i < nl - 1 or mp.gamma is not None) else o_real
            if i > 0:
                cols = [(0, 0, i_real)]                                   # (offset in the padded input, offset in the real input, width)
                i_pad = t
            elif in_blocks == 0:
                cols, i_pad = [(0, 0, i_real)], i_real
            else:
                assert i_real == in_blocks * latent
                cols, i_pad = [(j * t, j * latent, latent) for j in range(in_blocks)], in_blocks * t
            wp = torch.zeros((o_pad, i_pad), dtype=torch.float32, device=w.device)
            for po, ro, wd in cols:
                wp[:o_real, po:po + wd] = w.detach()[:, ro:ro + wd]
            bp = torch.zeros(o_pad, dtype=torch.float32, device=w.device)
            bp[:o_real] = b.detach()
            ws.append(wp)
            bs.append(bp)
            self.cuts.append((o_real, i_real, cols))
        gamma = beta = None
        if mp.gamma is not None:
            gamma = torch.zeros(t, dtype=torch.float32, device=mp.gamma.device)
            beta = torch.zeros(t, dtype=torch.float32, device=mp.gamma.device)
            gamma[:mp.out_dim] = mp.gamma.detach()
            beta[:mp.out_dim] = mp.beta.detach()
        super().__init__(ws, bs, gamma, beta)
        if mp.gamma is not None:
            self.ln_dim = mp.out_dim

    def unpad(self, grads):
        out = []
        for i, (o_real, i_real, cols) in enumerate(self.cuts):
            gw, gb = grads[2 * i], grads[2 * i + 1]
            w = torch.empty((o_real, i_real), dtype=torch.float32, device=gw.device)
            for po, ro, wd in cols:
                w[:, ro:ro + wd] = gw[:o_real, po:po + wd]
            out += [w, gb[:o_real].contiguous()]
        if self.gamma is not None:
            n = self.orig.out_dim
            out += [grads[-2][:n].contiguous(), grads[-1][:n].contiguous()]
        return out


def _rows_ws(mlp_c: CgnnMlp, rows: int, precision: str, backward: int, device):
    nbytes = lib().cgnn_mlp_rows_workspace_bytes(byref(mlp_c), rows, PREC[precision], backward)
    if nbytes < 0:
        check(-1, "cgnn_mlp_rows_workspace_bytes")
    return workspace.get(device, "mlp", nbytes)


def _bwd_ws(mlp_c: CgnnMlp, device):
    nbytes = lib().cgnn_mlp_bwd_workspace_bytes(byref(mlp_c))
    if nbytes < 0:
        check(-1, "cgnn_mlp_bwd_workspace_bytes")
    return workspace.get(device, "mlp", nbytes)


# ------------------------------------------------------------------------------------------------
# K3/K6 rows, K4/K5 message passing
# ------------------------------------------------------------------------------------------------
def mlp_rows_fwd(p: MlpParams, x: torch.Tensor, precision: str = "fp32", out: Optional[torch.Tensor] = None) -> torch.Tensor:
    require_cuda(x, "x", torch.float32)
    if out is None:
        out = torch.empty((x.shape[0], p.out_dim), dtype=torch.float32, device=x.device)
    else:
        require_cuda(out, "out", torch.float32)
        if tuple(out.shape) != (x.shape[0], p.out_dim):
            raise ValueError(f"out must be {(x.shape[0], p.out_dim)}, got {tuple(out.shape)}")
    m = p.c_struct()
    with torch.cuda.device(x.device):
        ws = _rows_ws(m, x.shape[0], precision, 0, x.device)
        check(lib().cgnn_mlp_rows_fwd(byref(m), ptr(x), x.shape[0], ptr(out), ptr(ws), ws.numel(), PREC[precision],
                                      stream_ptr(x.device)), "cgnn_mlp_rows_fwd")
    return out


def grad_stream_dtype(precision: str):
    """Element type of the gradient stream an entry point exchanges with its caller (`de_next` / `de` of the edge phase, `dout` of
    an MLP with LayerNorm): bfloat16 with "bf16x3g" (CGNN_PREC_BF16X3_G16), float32 otherwise."""
    return torch.bfloat16 if precision == "bf16x3g" else torch.float32


def mlp_rows_bwd(p: MlpParams, x: torch.Tensor, dout: torch.Tensor, need_dx: bool, precision: str = "fp32"):
    require_cuda(dout, "dout", grad_stream_dtype(precision) if p.gamma is not None else torch.float32)
    m = p.c_struct()
    g, grads = p.new_grads()
    dx = torch.empty_like(x) if need_dx else None
    with torch.cuda.device(x.device):
        ws = _rows_ws(m, x.shape[0], precision, 1, x.device)
        check(lib().cgnn_mlp_rows_bwd(byref(m), byref(g), ptr(x), x.shape[0], ptr(dout), ptr(dx), ptr(ws),
                                      ws.numel(), PREC[precision], stream_ptr(x.device)), "cgnn_mlp_rows_bwd")
    return grads, dx


# When a list, every cgnn_mp_edge_fwd launch (the dominant kernel) is bracketed by CUDA events on the
# launching stream and (start, end) is appended: bench.py reads the per-launch durations from here.
EDGE_FWD_EVENTS = None


def mp_edge_fwd(p: MlpParams, h, e_in, senders, k: int, e_out, agg_edge, precision: str = "fp32", k_valid: int = 0):
    """h may carry halo rows after the receivers (slab sharding): receivers = e_in rows / k, nodes = h rows.
    `k_valid`: the real in-degree when the k rows per receiver are padded (tensor-core precisions)."""
    m = p.c_struct()
    n_recv = e_in.shape[0] // k
    with torch.cuda.device(h.device):
        ev = None
        if EDGE_FWD_EVENTS is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record(torch.cuda.current_stream(h.device))
        nbytes = lib().cgnn_mp_edge_fwd_workspace_bytes(byref(m), h.shape[0], PREC[precision])
        if nbytes < 0:
            check(-1, "cgnn_mp_edge_fwd_workspace_bytes")
        ws = workspace.get(h.device, "mlp", nbytes) if nbytes > 0 else None
        check(lib().cgnn_mp_edge_fwd(byref(m), ptr(h), ptr(e_in), ptr(senders), n_recv, h.shape[0], k, int(k_valid), ptr(e_out),
                                     ptr(agg_edge), ptr(ws), 0 if ws is None else ws.numel(), PREC[precision],
                                     stream_ptr(h.device)), "cgnn_mp_edge_fwd")
        if ev is not None:
            ev[1].record(torch.cuda.current_stream(h.device))
            EDGE_FWD_EVENTS.append(ev)


def aggregate_senders(h, senders, k: int, agg):
    with torch.cuda.device(h.device):
        check(lib().cgnn_aggregate_senders(ptr(h), ptr(senders), agg.shape[0], k, h.shape[1], ptr(agg),
                                           stream_ptr(h.device)), "cgnn_aggregate_senders")


def mp_node_fwd(p: MlpParams, h, agg, h_out, precision: str = "fp32"):
    m = p.c_struct()
    with torch.cuda.device(h.device):
        nbytes = lib().cgnn_mp_node_fwd_workspace_bytes(byref(m), h.shape[0], PREC[precision])
        if nbytes < 0:
            check(-1, "cgnn_mp_node_fwd_workspace_bytes")
        ws = workspace.get(h.device, "mlp", nbytes) if nbytes > 0 else None
        check(lib().cgnn_mp_node_fwd(byref(m), ptr(h), ptr(agg), h.shape[0], ptr(h_out), ptr(ws),
                                     0 if ws is None else ws.numel(), PREC[precision], stream_ptr(h.device)),
              "cgnn_mp_node_fwd")


def presize_workspaces(edge_mlp: MlpParams, enc_edge: MlpParams, n: int, n_nodes: int, k: int, n_edges: int, precision: str,
                        device) -> None:
    """Grows the shared scratch buffers to what one training application of this size will ask for (edge phase
    forward / backward, edge encoder forward / backward), so that a memory plan made afterwards sees them."""
    em, ee = edge_mlp.c_struct(), enc_edge.c_struct()
    with torch.cuda.device(device):
        nb = lib().cgnn_mp_edge_fwd_workspace_bytes(byref(em), n_nodes, PREC[precision])
        if nb > 0:
            workspace.get(device, "mlp", nb)
        _mp_bwd_ws(em, n, k, precision, device, n_nodes=n_nodes)
        _rows_ws(ee, n_edges, precision, 1, device)
        _rows_ws(ee, n_edges, precision, 0, device)


def _mp_bwd_ws(mlp_c: CgnnMlp, n: int, k: int, precision: str, device, n_nodes=None):
    nbytes = lib().cgnn_mp_bwd_workspace_bytes(byref(mlp_c), n, n if n_nodes is None else n_nodes, k, PREC[precision])
    if nbytes < 0:
        check(-1, "cgnn_mp_bwd_workspace_bytes")
    return workspace.get(device, "mlp", nbytes)


def mp_node_bwd(p: MlpParams, h, agg, dh_next, dh, dagg, precision: str = "fp32"):
    m = p.c_struct()
    g, grads = p.new_grads()
    with torch.cuda.device(h.device):
        ws = _mp_bwd_ws(m, h.shape[0], 0, precision, h.device)
        check(lib().cgnn_mp_node_bwd(byref(m), byref(g), ptr(h), ptr(agg), ptr(dh_next), h.shape[0], ptr(dh),
                                     ptr(dagg), ptr(ws), ws.numel(), PREC[precision], stream_ptr(h.device)),
              "cgnn_mp_node_bwd")
    return grads


def mp_edge_bwd(p: MlpParams, h, e_in, senders, rowptr, perm, k: int, de_next, dagg, de, dh, gs,
                precision: str = "fp32", k_valid: int = 0):
    """Edge-phase backward including the deterministic scatter of the sender gradients into `dh`
    (`rowptr`, `perm`: sender-sorted transpose from `csr_transpose`).  `de` may be `de_next` (in place); `gs` [E,L] is
    the scratch of the FP32 kernels (None for the tensor-core precisions)."""
    for name, t in (("de_next", de_next), ("de", de)):
        if t is not None:
            require_cuda(t, name, grad_stream_dtype(precision))
    m = p.c_struct()
    g, grads = p.new_grads()
    n_recv = e_in.shape[0] // k
    with torch.cuda.device(h.device):
        ws = _mp_bwd_ws(m, n_recv, k, precision, h.device, n_nodes=h.shape[0])
        check(lib().cgnn_mp_edge_bwd(byref(m), byref(g), ptr(h), ptr(e_in), ptr(senders), ptr(rowptr), ptr(perm),
                                     n_recv, h.shape[0], k, int(k_valid), ptr(de_next), ptr(dagg), ptr(de), ptr(dh), ptr(gs), ptr(ws),
                                     ws.numel(), PREC[precision], stream_ptr(h.device)), "cgnn_mp_edge_bwd")
    return grads


def halo_pack(src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """src[idx] as a fresh contiguous send buffer (slab.HaloPlan.exchange); idx int64."""
    require_cuda(src, "src", torch.float32)
    require_cuda(idx, "idx", torch.int64)
    out = torch.empty((idx.numel(), src.shape[1]), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        check(lib().cgnn_halo_pack(ptr(src), ptr(idx), idx.numel(), src.shape[1], ptr(out), stream_ptr(src.device)), "cgnn_halo_pack")
    return out


def halo_unpack_add(src: torch.Tensor, idx: torch.Tensor, dst: torch.Tensor) -> None:
    """dst[idx] += src for unique idx (slab.HaloPlan.reduce_grad)."""
    require_cuda(src, "src", torch.float32)
    require_cuda(idx, "idx", torch.int64)
    require_cuda(dst, "dst", torch.float32)
    with torch.cuda.device(dst.device):
        check(lib().cgnn_halo_unpack_add(ptr(src), ptr(idx), idx.numel(), dst.shape[1], ptr(dst), stream_ptr(dst.device)),
              "cgnn_halo_unpack_add")


def scatter_to_senders(src, per_receiver: bool, rowptr, perm, k: int, dh):
    with torch.cuda.device(dh.device):
        check(lib().cgnn_scatter_to_senders(ptr(src), int(per_receiver), ptr(rowptr), ptr(perm), dh.shape[0], k,
                                            dh.shape[1], ptr(dh), stream_ptr(dh.device)), "cgnn_scatter_to_senders")


# ------------------------------------------------------------------------------------------------
# K6 loss
# ------------------------------------------------------------------------------------------------
def loss_fwd_bwd(acc, temp, y_acc, y_temp, graph_ptr, num_graphs: int, dt: float, w_acc: float, w_temp: float,
                 w_mom: float, want_grads: bool = True):
    """Returns (losses[4] device tensor = total, acc_mse, temp_mse, momentum; d_acc | None; d_temp | None)."""
    for name, t in (("acc", acc), ("temp", temp), ("y_acc", y_acc), ("y_temp", y_temp)):
        require_cuda(t, name, torch.float32)
    n, out_dim = acc.shape
    losses = torch.empty(4, dtype=torch.float32, device=acc.device)
    d_acc = torch.empty_like(acc) if want_grads else None
    d_temp = torch.empty_like(temp) if want_grads else None
    with torch.cuda.device(acc.device):
        nbytes = lib().cgnn_loss_workspace_bytes(n, num_graphs)
        ws = workspace.get(acc.device, "loss", nbytes)
        check(lib().cgnn_loss_fwd_bwd(ptr(acc), ptr(temp), ptr(y_acc), ptr(y_temp), ptr(graph_ptr), n, out_dim,
                                      num_graphs, float(dt), float(w_acc), float(w_temp), float(w_mom),
                                      ptr(losses), ptr(d_acc), ptr(d_temp), ptr(ws), ws.numel(),
                                      stream_ptr(acc.device)), "cgnn_loss_fwd_bwd")
    return losses, d_acc, d_temp
