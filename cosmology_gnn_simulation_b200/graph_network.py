"""Drop-in for the reference's `graph_network` module (graph_network.py:108-187).

`EncodeProcessDecode` keeps the reference constructor, `forward(input_graph)` contract and
`state_dict` key names, but its forward and backward run as hand-written sm_100a
kernels behind the C ABI of `libcgnn.so` (include/cgnn.h):

  encoder (graph_network.py:52-64)          -> cgnn_mlp_rows_fwd / _bwd
  M x InteractionNetwork + residuals         -> cgnn_mp_edge_fwd, cgnn_aggregate_senders,
    (graph_network.py:83-101,177-183)          cgnn_mp_node_fwd  (+ _bwd with in-tile recompute)
  decoders (graph_network.py:151-152,158-164)-> cgnn_mlp_rows_fwd / _bwd

The torch modules below only hold parameters (so `.parameters()`, `.to()`, `state_dict()`, lazy first
layers and stock optimizers behave exactly like the reference); they are never called.  There is no
CPU or PyTorch fallback: CPU inputs raise.

`message` selects what is summed at the receivers (SURVEY finding F2):
  "sender" (default) - what the reference actually computes: PyG's default message, the sender
                        node latent; the edge stream then gets no gradient (its grads stay None).
  "edge"             - the intended Interaction Network: the updated edge latent is the message.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ckpt_plan, ops
from .ops import MlpParams

__all__ = ["EncodeProcessDecode", "build_mlp"]


def build_mlp(hidden_size: int, num_hidden_layers: int, output_size: int) -> nn.Module:
    """Parameter holder with the reference's layout: Linear layers at Sequential indices 0,2,4,...
    (first one lazy, graph_network.py:15-32).  The ReLUs are fused into the kernels."""
    mods: List[nn.Module] = []
    for depth in range(num_hidden_layers):
        mods += [nn.LazyLinear(hidden_size) if depth == 0 else nn.Linear(hidden_size, hidden_size),
                 nn.ReLU(inplace=True)]
    mods.append(nn.Linear(hidden_size, output_size))
    return nn.Sequential(*mods)


class _NodeEdgePair(nn.Module):
    """`node_model` / `edge_model` holder (registration order as graph_network.py:49-50,80-81)."""

    def __init__(self, node_model: nn.Module, edge_model: nn.Module):
        super().__init__()
        self.node_model = node_model
        self.edge_model = edge_model


def _linears(seq: nn.Sequential) -> List[nn.Module]:
    return [m for m in seq if isinstance(m, (nn.Linear, nn.LazyLinear))]


def _materialize(seq: nn.Sequential, in_dim: int, like: torch.Tensor) -> None:
    """Gives a still-lazy first layer its shape, as the reference's first forward would."""
    first = _linears(seq)[0]
    if isinstance(first, nn.LazyLinear) and first.has_uninitialized_params():
        probe = torch.empty((1, in_dim), dtype=like.dtype, device=like.device)
        first._infer_parameters(first, (probe,))     # initialises and turns the module into nn.Linear


def _mlp_params(seq: nn.Sequential, norm: Optional[nn.LayerNorm]) -> MlpParams:
    lin = _linears(seq)
    return MlpParams([m.weight for m in lin], [m.bias for m in lin],
                     None if norm is None else norm.weight, None if norm is None else norm.bias)


GRAD16_MIN_ROWS = 1 << 16         # edge rows from which the bfloat16 gradient stream is used (more than one wave of tiles: the
                                  # backward of shorter streams runs the fused chains, csrc/tc_model.cu bwd_mode)
REORDER_MIN_NODES = 1 << 17       # below this the per-node tables the edges gather from live in L2 whatever the numbering


def _morton_order(pos: torch.Tensor, box) -> torch.Tensor:
    """Particle ids sorted along a Z-order curve of a grid with ~8 particles per cell (device ops only, no synchronisation)."""
    n = pos.shape[0]
    nc = int(min(1024, max(1, round((n / 8.0) ** (1.0 / 3.0)))))
    if box is None:
        box = pos.max() + 1e-6
    c = torch.clamp((pos.float() / box * nc).long(), 0, nc - 1)

    def spread(v):                               # 10 bits -> every third bit
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        return (v | (v << 2)) & 0x09249249

    code = spread(c[:, 0]) | (spread(c[:, 1]) << 1) | (spread(c[:, 2]) << 2)
    return torch.sort(code, stable=True)[1]


class _Plan:
    """Static description of one forward call, shared by forward and backward."""

    def __init__(self, n_steps, message, precision, k, enc_node, enc_edge, proc_node, proc_edge, dec_acc,
                 dec_temp, groups, edge_buffers, halo=None, k_valid=0, grad_stream="fp32"):
        self.halo = halo                      # slab.HaloPlan of a sharded box, or None
        self.n_steps, self.message, self.precision, self.k = n_steps, message, precision, k
        # precision of the backward over the EDGE streams (processor edge MLPs, edge encoder): "bf16x3g" keeps the three gradient
        # intermediates of the long-stream composition as bfloat16 (csrc/tc_model.cu grad16; error study tests/study_grad_stream.py)
        self.grad_stream = grad_stream
        self.k_valid = k_valid                # real in-degree when k is the padded power of two (0: k itself)
        self.enc_node, self.enc_edge = enc_node, enc_edge
        self.proc_node, self.proc_edge = proc_node, proc_edge
        self.dec_acc, self.dec_temp = dec_acc, dec_temp
        self.groups = groups                  # [(MlpParams, first index into the flat parameter list)]
        self.edge_buffers = edge_buffers      # forced number of edge-stream buffers (0: from the free memory)
        self.grad_enabled = torch.is_grad_enabled()      # sampled where the module is called (autograd runs Function.forward in no-grad mode)

    def edge_bwd_precision(self, n_edges: int) -> str:
        """"bf16x3g" (CGNN_PREC_BF16X3_G16) for edge streams long enough for the layered backward composition, where it pays and where
        the rounding of the gradient stream averages out; the model's own precision otherwise."""
        if self.precision == "bf16x3" and self.grad_stream == "bf16" and n_edges >= GRAD16_MIN_ROWS:
            return "bf16x3g"
        return self.precision

    def padded_for_tensor_cores(self):
        """The same plan over zero-padded copies of the parameters (ops.PaddedMlp) when the tensor-core precisions meet a
        latent / hidden width below 128; None when no padding is needed (or possible: widths above 128 keep the FP32 path)."""
        L, H = self.enc_node.out_dim, self.enc_node.hidden
        t = ops.TC_WIDTH
        if self.precision == "fp32" or (L == t and H == t) or L > t or H > t or len(self.enc_node.weights) != 3:
            return None
        pad = {}

        def p(mp, blocks):
            pad[id(mp)] = ops.PaddedMlp(mp, blocks, L)
            return pad[id(mp)]

        groups_by_id = {id(mp): first for mp, first in self.groups}
        q = _Plan(self.n_steps, self.message, self.precision, self.k, p(self.enc_node, 0), p(self.enc_edge, 0),
                  [p(m, 2) for m in self.proc_node], [p(m, 3) for m in self.proc_edge], p(self.dec_acc, 1), p(self.dec_temp, 1),
                  None, self.edge_buffers, self.halo, self.k_valid, self.grad_stream)
        q.groups = [(pad[i], first) for i, first in groups_by_id.items()]
        q.grad_enabled = self.grad_enabled
        return q


_STREAM_PLANS = {}


class _StreamLease:
    """The edge-stream buffers of one training application: nbuf copies of e^t plus the gradient stream, E x L x 4 bytes each.

    They are slices of the grow-only shared scratch (`_lib.workspace`, tags "edge_stream_i"; a CUDA-graph capture has its own):
    32 GiB blocks that come and go through the caching allocator every step fragment it until a block no longer fits although
    the memory is there (seen at 2.1 M particles: 134 GiB allocated, 32 GiB reserved in pieces, 32 GiB request fails).  One
    application holds the lease from its forward to the end of its backward; a second one started in between (no reference
    loop does that) falls back to private tensors."""
    _busy = set()

    def __init__(self, device, n_buffers: int, e_count: int, latent: int, grad_dtype=None):
        """n_buffers FP32 copies of the edge stream in `bufs`; with `grad_dtype` also the gradient stream `grad` in that element
        type (bfloat16 for long streams, DESIGN.md section 3: half the size of a copy)."""
        from ._lib import workspace
        self.key = (str(device), id(workspace._bufs))
        self.shared = self.key not in _StreamLease._busy
        nbytes = e_count * latent * 4
        self.grad = None
        if self.shared:
            _StreamLease._busy.add(self.key)
            self.bufs = [workspace.get(device, f"edge_stream_{i}", nbytes)[:nbytes].view(torch.float32).view(e_count, latent)
                         for i in range(n_buffers)]
            if grad_dtype is not None:
                gbytes = e_count * latent * torch.empty((), dtype=grad_dtype).element_size()
                self.grad = workspace.get(device, "edge_stream_grad", gbytes)[:gbytes].view(grad_dtype).view(e_count, latent)
        else:
            self.bufs = [torch.empty((e_count, latent), dtype=torch.float32, device=device) for _ in range(n_buffers)]
            if grad_dtype is not None:
                self.grad = torch.empty((e_count, latent), dtype=grad_dtype, device=device)

    def release(self):
        if self.shared:
            _StreamLease._busy.discard(self.key)
            self.shared = False
        self.bufs = self.grad = None

    def __del__(self):
        self.release()


def _edge_stream_buffers(plan: _Plan, n: int, n_loc: int, e_count: int, L: int, device) -> int:
    """How many copies of the edge latent stream (E x L x 4 bytes each) message="edge" training may keep next to the
    gradient stream; ckpt_plan.schedule() turns that into the checkpoint / recompute schedule of the backward.
    Decided once per (steps, sizes, device): cudaMemGetInfo costs ~15 ms of host time per call, far too much for
    every step.  The workspaces of the backward are sized first so that the measurement sees them."""
    M = plan.n_steps
    forced = plan.edge_buffers or int(os.environ.get("CGNN_EDGE_BUFFERS", "0"))
    if forced:
        return min(int(forced), M)
    grad_copy = e_count * L * torch.empty((), dtype=ops.grad_stream_dtype(plan.edge_bwd_precision(e_count))).element_size()
    key = (M, n, n_loc, e_count, L, plan.precision, grad_copy, str(device))
    nbuf = _STREAM_PLANS.get(key)
    if nbuf is not None:
        return nbuf
    copy = e_count * L * 4
    ops.presize_workspaces(plan.proc_edge[0], plan.enc_edge, n, n_loc, plan.k, e_count, plan.precision, device)
    if copy >= (1 << 30):
        torch.cuda.empty_cache()          # large streams: hand cached fragments back before measuring
    free, total = torch.cuda.mem_get_info(device)
    free += torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
    from ._lib import workspace
    free += workspace.tagged_bytes(device, "edge_stream_")      # stream buffers of an earlier plan are reused, not added
    # still to come besides the stream copies: h^t (M+1) and agg^t (M) for every step, six node-sized temporaries of the
    # backward, the sender-sorted transpose, 1 GiB of small tensors -- and 3 % of the device is left untouched
    other = (2 * M + 7) * n_loc * L * 4 + 6 * e_count + (1 << 30) + (total * 3) // 100
    nbuf = min(M, (free - other - grad_copy) // copy)           # next to the gradient stream de (half a copy as bfloat16)
    if nbuf < 1:
        raise RuntimeError(
            f"cgnn: the edge latent stream does not fit: {copy / 2**30:.1f} GiB per copy, two copies needed, "
            f"{free / 2**30:.1f} GiB free; shard the box over more GPUs or use message='sender'")
    _STREAM_PLANS[key] = int(nbuf)
    return int(nbuf)


class _EncodeProcessDecodeFn(torch.autograd.Function):
    """Whole-model forward/backward over the C ABI; owns every saved activation."""

    @staticmethod
    def forward(ctx, plan: _Plan, senders, transpose_fn, x, edge_attr, *params):
        plan = plan.padded_for_tensor_cores() or plan       # widths below 128: zero-padded parameters, 128-wide latents inside
        p, M, k, prec = plan, plan.n_steps, plan.k, plan.precision
        n, L = x.shape[0], plan.enc_node.out_dim
        e_count = edge_attr.shape[0]
        dev = x.device
        # (needs_input_grad reflects requires_grad only: under torch.no_grad() nothing will be differentiated)
        train = plan.grad_enabled and any(ctx.needs_input_grad[3:])
        edge_mode = p.message == "edge"

        halo = p.halo
        n_loc = n if halo is None else halo.n_loc          # node array of a slab rank: [owned | halo senders]

        def with_halo(h_own_rows):
            if halo is None:
                return h_own_rows
            full = torch.empty((n_loc, L), dtype=torch.float32, device=dev)
            full[:n] = h_own_rows
            halo.exchange(full)
            return full

        h = with_halo(ops.mlp_rows_fwd(p.enc_node, x, prec))
        hs, aggs = [h], []

        def node_phase(t, h, agg):
            # inference updates h in place: a node tile only reads its own rows of h
            h_next = torch.empty_like(h) if train else h
            ops.mp_node_fwd(p.proc_node[t], h[:n], agg, h_next[:n], prec)
            if halo is not None:
                halo.exchange(h_next)                       # the one collective of a message-passing step
            if train:
                hs.append(h_next)
                aggs.append(agg)
            return h_next

        bufs, acts, pos, lease = None, None, 0, None
        if not edge_mode:
            # What the reference computes (PyG's default message): the receivers sum SENDER latents.  The edge
            # latents never reach h, the decoders or any gradient, so the edge stream is not computed at all.
            for t in range(M):
                agg = torch.empty((n, L), dtype=torch.float32, device=dev)
                ops.aggregate_senders(h, senders, k, agg)
                h = node_phase(t, h, agg)
        elif not train:
            lease = _StreamLease(dev, 1, e_count, L)
            e = ops.mlp_rows_fwd(p.enc_edge, edge_attr, prec, out=lease.bufs[0])
            for t in range(M):
                agg = torch.empty((n, L), dtype=torch.float32, device=dev)
                ops.mp_edge_fwd(p.proc_edge[t], h, e, senders, k, e if t + 1 < M else None, agg, prec, p.k_valid)   # e^M is never read
                h = node_phase(t, h, agg)
            del e
            lease.release()
        else:
            nbuf = _edge_stream_buffers(p, n, n_loc, e_count, L, dev)
            acts = ckpt_plan.schedule(M, nbuf)
            lease = _StreamLease(dev, nbuf, e_count, L,               # nbuf copies of e^t + the gradient stream
                                 grad_dtype=ops.grad_stream_dtype(p.edge_bwd_precision(e_count)))
            bufs = lease.bufs
            last = None
            for pos, act in enumerate(acts):
                if act[0] == "bwd":
                    last = act[2]
                    break
                _EncodeProcessDecodeFn._advance(p, act, bufs, edge_attr, senders, hs, e_count, L, first=True,
                                                node_phase=node_phase)
            agg = torch.empty((n, L), dtype=torch.float32, device=dev)
            ops.mp_edge_fwd(p.proc_edge[M - 1], hs[M - 1], bufs[last], senders, k, None, agg, prec, p.k_valid)     # e^M is never read
            node_phase(M - 1, hs[M - 1], agg)
            h = hs[M]
        acc = ops.mlp_rows_fwd(p.dec_acc, h[:n], prec)
        temp = ops.mlp_rows_fwd(p.dec_temp, h[:n], prec)

        if train:
            ctx.plan, ctx.senders, ctx.transpose_fn = plan, senders, transpose_fn
            ctx.x, ctx.edge_attr = x, edge_attr
            ctx.hs, ctx.aggs, ctx.bufs, ctx.acts, ctx.pos = hs, aggs, bufs, acts, pos
            ctx.lease = lease
            ctx.n_params = len(params)
        return acc, temp

    @staticmethod
    def _advance(p, act, bufs, edge_attr, senders, hs, e_count, L, first, node_phase=None):
        """One "enc" / "adv" action of the checkpoint schedule on the stream buffers."""
        prec, k = p.precision, p.k
        if act[0] == "enc":
            b = act[1]
            ops.mlp_rows_fwd(p.enc_edge, edge_attr, prec, out=bufs[b])
            return
        _, t, src, dst = act
        if first:                                           # the real forward of step t
            h = hs[t]
            agg = torch.empty((h.shape[0] if p.halo is None else p.halo.n_own, L), dtype=torch.float32, device=h.device)
            ops.mp_edge_fwd(p.proc_edge[t], h, bufs[src], senders, k, bufs[dst], agg, prec, p.k_valid)
            node_phase(t, h, agg)
        else:                                               # recompute of the edge phase alone
            ops.mp_edge_fwd(p.proc_edge[t], hs[t], bufs[src], senders, k, bufs[dst], None, prec, p.k_valid)

    @staticmethod
    def backward(ctx, d_acc, d_temp):
        p: _Plan = ctx.plan
        M, k, prec = p.n_steps, p.k, p.precision
        senders = ctx.senders
        hs, aggs = ctx.hs, ctx.aggs
        edge_mode = p.message == "edge"
        grads: List[Optional[torch.Tensor]] = [None] * ctx.n_params
        index_of = {id(mp): first for mp, first in p.groups}

        def put(mp: MlpParams, tensors):
            if isinstance(mp, ops.PaddedMlp):
                tensors = mp.unpad(tensors)
            first = index_of[id(mp)]
            for i, g in enumerate(tensors):
                grads[first + i] = g

        n_loc, L = hs[0].shape
        halo = p.halo
        n = n_loc if halo is None else halo.n_own
        d_acc = d_acc.contiguous() if d_acc is not None else torch.zeros((n, p.dec_acc.out_dim), device=hs[0].device)
        d_temp = d_temp.contiguous() if d_temp is not None else torch.zeros((n, 1), device=hs[0].device)

        g_acc, dh_a = ops.mlp_rows_bwd(p.dec_acc, hs[M][:n], d_acc, True, prec)
        g_temp, dh_t = ops.mlp_rows_bwd(p.dec_temp, hs[M][:n], d_temp, True, prec)
        put(p.dec_acc, g_acc)
        put(p.dec_temp, g_temp)
        if halo is None:
            dh = dh_a.add_(dh_t)
        else:
            dh = torch.zeros((n_loc, L), dtype=torch.float32, device=dh_a.device)
            dh[:n] = dh_a.add_(dh_t)
        del dh_a, dh_t
        de = None
        rowptr, perm = ctx.transpose_fn()
        fp32 = prec == "fp32"
        prec_e = p.edge_bwd_precision(ctx.edge_attr.shape[0]) if edge_mode else prec
        de_dtype = ops.grad_stream_dtype(prec_e)

        def step_backward(t, e_t, dh, de):
            if halo is not None:
                halo.reduce_grad(dh)              # gradients other ranks hold for my rows come home first
                dh_new = torch.zeros_like(dh)
            else:
                dh_new = torch.empty_like(dh)
            dagg = torch.empty((n, L), dtype=torch.float32, device=dh.device)
            put(p.proc_node[t], ops.mp_node_bwd(p.proc_node[t], hs[t][:n], aggs[t], dh[:n], dh_new[:n], dagg, prec))
            if edge_mode:
                # the gradient stream is updated in place (de^t over de^{t+1}); only the FP32 kernels need the per-edge
                # scratch gs, the tensor-core path accumulates the sender sums chunk by chunk in its workspace
                de_new = de if de is not None else ctx.lease.grad
                assert de_new.dtype == de_dtype
                gs = torch.empty_like(e_t) if fp32 else None
                put(p.proc_edge[t], ops.mp_edge_bwd(p.proc_edge[t], hs[t], e_t, senders, rowptr, perm, k, de, dagg,
                                                    de_new, dh_new, gs, prec_e, p.k_valid))
                return dh_new, de_new
            ops.scatter_to_senders(dagg, True, rowptr, perm, k, dh_new)
            return dh_new, None

        if edge_mode:
            bufs, acts = ctx.bufs, ctx.acts
            e_count = ctx.edge_attr.shape[0]
            for act in acts[ctx.pos:]:
                if act[0] == "bwd":
                    _, t, b = act
                    dh, de = step_backward(t, bufs[b], dh, de)
                    hs[t + 1] = None                         # h^{t+1} and agg^t have had their last reader
                    aggs[t] = None
                else:
                    _EncodeProcessDecodeFn._advance(p, act, bufs, ctx.edge_attr, senders, hs, e_count, L, first=False)
        else:
            for t in range(M - 1, -1, -1):
                dh, de = step_backward(t, None, dh, None)

        need_dx, need_dea = ctx.needs_input_grad[3], ctx.needs_input_grad[4]
        if halo is not None:
            halo.reduce_grad(dh)
        g_en, dx = ops.mlp_rows_bwd(p.enc_node, ctx.x, dh[:n].contiguous(), need_dx, prec)
        put(p.enc_node, g_en)
        dea = None
        if edge_mode:
            if need_dea and de.dtype != torch.float32:     # (an input gradient narrower than 128 columns comes from the FP32 kernels)
                de, prec_e = de.float(), prec
            g_ee, dea = ops.mlp_rows_bwd(p.enc_edge, ctx.edge_attr, de, need_dea, prec_e)
            put(p.enc_edge, g_ee)
        if ctx.lease is not None:
            ctx.lease.release()
        ctx.hs = ctx.aggs = ctx.bufs = ctx.lease = None
        return (None, None, None, dx, dea, *grads)


class EncodeProcessDecode(nn.Module):
    """Encode-process-decode Interaction Network (graph_network.py:108-164), B200-native.

    Constructor arguments as the reference; the keyword-only extras select kernel behaviour:
      num_neighbors  optional hint (k is otherwise read from the graph: E / N)
      message        "sender" (reference-actual, default) | "edge" (intended Interaction Network); env CGNN_MESSAGE
      precision      "fp32" (FP32 SIMT, <= 1e-5 parity, default) | "bf16x3" | "bf16" (tcgen05 tensor cores); env CGNN_PRECISION
      edge_buffers   message="edge" training: copies of the edge stream kept for the backward (0 = from the free memory)
      grad_stream    precision="bf16x3" only: "bf16" (default) | "fp32"; env CGNN_GRAD_STREAM.  "bf16": the backward of the edge MLPs
                     over long streams (GRAD16_MIN_ROWS edge rows or more) keeps the gradient stream de^t it carries from step to
                     step and its gradient intermediates dY / G2 / G1 as bfloat16 in HBM.  Forward values and ReLU gates are
                     unchanged; the rounding averages out in the weight gradients (tests/study_grad_stream.py: 1.3e-4 at 4 096
                     particles, shrinking with the square root of the size).
    """

    def __init__(self, latent_size: int, mlp_hidden_size: int, mlp_num_hidden_layers: int,
                 num_message_passing_steps: int, output_size: int, *, num_neighbors: Optional[int] = None,
                 message: Optional[str] = None, precision: Optional[str] = None, edge_buffers: int = 0,
                 grad_stream: Optional[str] = None):
        super().__init__()
        # the reference's scripts construct the model with its five arguments only: the environment chooses for them
        message = message if message is not None else os.environ.get("CGNN_MESSAGE", "sender")
        precision = precision if precision is not None else os.environ.get("CGNN_PRECISION", "fp32")
        if message not in ("sender", "edge"):
            raise ValueError("message must be 'sender' or 'edge'")
        if precision not in ("fp32", "bf16x3", "bf16"):
            raise ValueError("precision must be one of 'fp32', 'bf16x3', 'bf16'")
        grad_stream = grad_stream if grad_stream is not None else os.environ.get("CGNN_GRAD_STREAM", "bf16")
        if grad_stream not in ("bf16", "fp32"):
            raise ValueError("grad_stream must be 'bf16' or 'fp32'")
        self._latent_size = latent_size
        self._mlp_hidden_size = mlp_hidden_size
        self._mlp_num_hidden_layers = mlp_num_hidden_layers
        self._num_message_passing_steps = num_message_passing_steps
        self._output_size = output_size
        self.num_neighbors = num_neighbors
        self.message = message
        self.precision = precision
        self.grad_stream = grad_stream
        self.edge_buffers = edge_buffers        # message='edge' training: copies of the edge stream to keep (0: from the free memory)

        def mlp_ln():
            return nn.Sequential(build_mlp(mlp_hidden_size, mlp_num_hidden_layers, latent_size),
                                 nn.LayerNorm(latent_size))

        self.encoder = _NodeEdgePair(node_model=mlp_ln(), edge_model=mlp_ln())
        self.processor = nn.ModuleList()
        for _ in range(num_message_passing_steps):
            edge_model = mlp_ln()                     # constructed first, registered second (as the reference)
            node_model = mlp_ln()
            self.processor.append(_NodeEdgePair(node_model=node_model, edge_model=edge_model))
        self.decoder_acc = build_mlp(mlp_hidden_size, mlp_num_hidden_layers, output_size)
        self.decoder_temp_rate = build_mlp(mlp_hidden_size, mlp_num_hidden_layers, 1)
        self._graph_cache = {}

    # ------------------------------------------------------------------------------------------
    def _materialize_all(self, node_in: int, edge_in: int, like: torch.Tensor) -> None:
        L = self._latent_size
        _materialize(self.encoder.node_model[0], node_in, like)
        _materialize(self.encoder.edge_model[0], edge_in, like)
        for blk in self.processor:                          # the reference's first forward reaches edge_model first (graph_network.py:90,96)
            _materialize(blk.edge_model[0], 3 * L, like)
            _materialize(blk.node_model[0], 2 * L, like)
        _materialize(self.decoder_acc, L, like)
        _materialize(self.decoder_temp_rate, L, like)

    def _graph_tables(self, graph, n: int):
        """int32 senders (ELL: edge e = receiver*k + rank) and a lazy sender-sorted transpose."""
        senders = getattr(graph, "_cgnn_senders", None)
        edge_index = getattr(graph, "edge_index", None)
        if senders is None and edge_index is None:
            raise ValueError("cgnn: the graph has neither edge_index nor the int32 neighbour table of preprocess")
        if edge_index is not None and (senders is None or senders.device != edge_index.device):
            key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version)
            hit = self._graph_cache.get("key") == key
            if not hit:
                self._graph_cache = {"key": key, "edge_index": edge_index,
                                     "senders": ops.senders_from_edge_index(edge_index.contiguous(), n)}
            senders = self._graph_cache["senders"]
        k = senders.numel() // n
        if senders.numel() != n * k or k < 1:
            raise ValueError("cgnn: edge count is not a multiple of the node count")
        halo = getattr(graph, "halo", None)
        n_nodes = n if halo is None else halo.n_loc          # transpose rows: owned + halo senders
        # Large graphs are renumbered along a space-filling curve INSIDE the model: an edge's sender then sits close to its
        # receiver in memory, so the per-node tables the edge phases gather from (P_s in forward and recompute, the per-edge
        # sender gradients in backward) are served from L2 instead of HBM (2.1 M particles: 104.6 GB of DRAM traffic per edge
        # forward with the caller's numbering against 68.9 GB algorithmic).  Rows are computed independently, so the outputs are
        # bit-identical under the renumbering; sums over edges change their order (gradients agree to rounding).
        order = inv = None
        pos = getattr(graph, "pos", None)
        want = os.environ.get("CGNN_REORDER")
        if torch.is_tensor(pos) and pos.dim() == 2 and pos.shape[0] == n and pos.shape[1] == 3 and pos.device == senders.device \
                and (want == "1" or (want is None and n >= REORDER_MIN_NODES)):
            key = (senders.data_ptr(), senders._version, pos.data_ptr(), pos._version, n, k)
            if self._graph_cache.get("order_key") != key:
                box = getattr(graph, "box_size", None)
                box = box.reshape(-1)[0] if torch.is_tensor(box) and box.numel() >= 1 and box.device == pos.device else None
                o = _morton_order(pos, box)
                iv = torch.empty_like(o)
                iv[o] = torch.arange(n, device=o.device)
                # new id of every edge's sender, receivers in new order (a slab rank's halo senders keep their ids behind the owned rows)
                new_id = iv.to(torch.int32) if n_nodes == n else torch.cat([iv.to(torch.int32),
                                                                             torch.arange(n, n_nodes, dtype=torch.int32, device=o.device)])
                sp = new_id[senders.view(n, k)[o].long()].reshape(-1).contiguous()
                self._graph_cache.update(order_key=key, order=o, inv=iv, order_senders=sp,
                                         order_halo=None if halo is None else halo.renumbered(iv))
            order, inv, senders = self._graph_cache["order"], self._graph_cache["inv"], self._graph_cache["order_senders"]
            halo = self._graph_cache["order_halo"]
        if self.num_neighbors is not None and k != self.num_neighbors:
            raise ValueError(f"graph has in-degree {k}, model was built with num_neighbors={self.num_neighbors}")
        k_valid = 0
        if self.precision != "fp32" and self.message == "edge" and (k & (k - 1)) != 0:
            # The tensor-core chain takes a power-of-two in-degree (a receiver's rows are a warp-aligned group): pad every
            # receiver to the next power of two with dummy edges (sender = the receiver itself); the kernels keep rows of
            # rank >= k_valid out of the per-receiver sums and give them no gradient.
            if k > 32:
                raise ValueError(f"the tensor-core precisions support at most 32 neighbours (graph has in-degree {k})")
            k_pad = 1 << (k - 1).bit_length()
            key = (senders.data_ptr(), senders._version, n, k_pad)
            if self._graph_cache.get("pad_key") != key:
                sp = torch.arange(n, dtype=torch.int32, device=senders.device).view(n, 1).repeat(1, k_pad)
                sp[:, :k] = senders.view(n, k)
                self._graph_cache["pad_key"], self._graph_cache["pad_senders"] = key, sp.reshape(-1).contiguous()
            senders, k_valid, k = self._graph_cache["pad_senders"], k, k_pad
        holder = {}

        def transpose():
            if "t" not in holder:
                holder["t"] = ops.csr_transpose(senders, n_nodes)
            return holder["t"]

        return senders, k, k_valid, transpose, order, inv, halo

    def forward(self, input_graph) -> Dict[str, torch.Tensor]:
        x, edge_attr = input_graph.x, input_graph.edge_attr
        if edge_attr is None:
            raise ValueError("edge_attr must not be None in InteractionNetwork")     # graph_network.py:86-87
        if not x.is_cuda:
            raise RuntimeError("cgnn EncodeProcessDecode runs on CUDA tensors only (no CPU path): "
                               "move the graph and the model to the GPU")
        glob = getattr(input_graph, "globals", None)
        if glob is not None:                                                         # graph_network.py:168-173
            x = torch.cat([x, glob.unsqueeze(0).expand(x.size(0), -1)], dim=-1)
        x = x.float().contiguous()
        edge_attr = edge_attr.float().contiguous()
        self._materialize_all(x.shape[1], edge_attr.shape[1], x)
        if next(self.parameters()).device != x.device:
            raise RuntimeError("cgnn: model parameters and graph must live on the same CUDA device")

        enc_node = _mlp_params(self.encoder.node_model[0], self.encoder.node_model[1])
        enc_edge = _mlp_params(self.encoder.edge_model[0], self.encoder.edge_model[1])
        proc_node = [_mlp_params(b.node_model[0], b.node_model[1]) for b in self.processor]
        proc_edge = [_mlp_params(b.edge_model[0], b.edge_model[1]) for b in self.processor]
        dec_acc = _mlp_params(self.decoder_acc, None)
        dec_temp = _mlp_params(self.decoder_temp_rate, None)
        flat: List[torch.Tensor] = []
        groups = []
        for mp in [enc_node, enc_edge, *proc_node, *proc_edge, dec_acc, dec_temp]:
            groups.append((mp, len(flat)))
            flat += mp.tensors()

        n = x.shape[0]
        senders, k, k_valid, transpose, order, inv, halo = self._graph_tables(input_graph, n)
        if order is not None:                             # the model works in the curve order, the caller never sees it
            k_real = k_valid or k
            x = x[order]
            edge_attr = edge_attr.view(n, k_real, edge_attr.shape[1])[order].reshape(n * k_real, -1)
        if k_valid:                                       # dummy edges carry zero features (differentiable w.r.t. the real ones)
            padded = edge_attr.new_zeros((n, k, edge_attr.shape[1]))
            padded[:, :k_valid] = edge_attr.view(n, k_valid, edge_attr.shape[1])
            edge_attr = padded.view(n * k, -1)
        plan = _Plan(self._num_message_passing_steps, self.message, self.precision, k, enc_node, enc_edge,
                     proc_node, proc_edge, dec_acc, dec_temp, groups, self.edge_buffers,
                     halo=halo, k_valid=k_valid, grad_stream=getattr(self, "grad_stream", "fp32"))
        acc, temp = _EncodeProcessDecodeFn.apply(plan, senders, transpose, x, edge_attr, *flat)
        if order is not None:
            acc, temp = acc[inv], temp[inv]
        return {"acceleration": acc, "temp_rate": temp}
