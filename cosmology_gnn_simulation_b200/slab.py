"""Spatial slab decomposition of one periodic box over the GPUs of a node.

The reference has no distributed code; this is the multi-GPU form of its hot path the north star asks for.

* Particles are ordered by the x coordinate of their most recent position; rank p owns the contiguous
  range [p*N/P, (p+1)*N/P) of that order (equal particle counts, so clustered boxes stay balanced).
* A rank owns its particles as RECEIVERS and therefore all their k in-edges and edge latents: edges never
  move.  Its node array is [owned | halo], the halo being the remote senders its edges reference
  (grouped by owner rank, ascending global id inside a group).
* Per message-passing step there is ONE exchange: the owners send the fresh latents of the rows other
  ranks hold as halo (`HaloPlan.exchange`); in backward the halo rows' gradients travel the other way and
  are added on the owner in fixed peer order (`HaloPlan.reduce_grad`) -- deterministic.
* The loss needs one all-reduce of five floats (two squared-error sums and the momentum sum), the
  parameter gradients one all-reduce(SUM) (every edge and every owned node lives on exactly one rank).

Transport: `torch.distributed` point-to-point batches -- NCCL over NVLink on the B200 node, gloo in the CPU
tests (the index logic is plain torch and runs on either device).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


def _native(t: torch.Tensor) -> bool:
    """The row copies of the exchange run as libcgnn kernels for what the model hands over (float32 latent rows on the GPU, width a
    multiple of 4); the gloo tests of the index logic (CPU tensors) keep the torch expressions."""
    return t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.shape[1] % 4 == 0 and t.is_contiguous()


def _rows(t: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    if _native(t):
        from . import ops
        return ops.halo_pack(t, idx)                                         # cgnn_halo_pack
    return t[idx].contiguous()


def slab_bounds(n: int, world: int) -> List[int]:
    """Equal-count ownership ranges of the x-sorted particles: rank p owns [b[p], b[p+1])."""
    return [(p * n) // world for p in range(world + 1)]


class HaloPlan:
    """Which owned rows go to which peer, and where the rows received from each peer land."""

    def __init__(self, rank: int, world: int, n_own: int, n_global: int, send_idx: Sequence[torch.Tensor],
                 recv_counts: Sequence[int], group=None):
        self.rank, self.world, self.n_own, self.n_global, self.group = rank, world, n_own, n_global, group
        self.send_idx = [s.long() for s in send_idx]              # per peer: local owned row ids it needs from me
        self.recv_counts = [int(c) for c in recv_counts]          # per peer: halo rows I hold of its particles
        self.recv_off = [0]
        for c in self.recv_counts:
            self.recv_off.append(self.recv_off[-1] + c)
        self.n_halo = self.recv_off[-1]
        self.n_loc = n_own + self.n_halo

    # -- transport ---------------------------------------------------------------------------------
    def _p2p(self, sends: Dict[int, torch.Tensor], recvs: Dict[int, torch.Tensor]) -> None:
        ops = []
        for q in range(self.world):                                # same global order on every rank
            if q == self.rank:
                continue
            if q in sends and sends[q].numel():
                ops.append(dist.P2POp(dist.isend, sends[q], q, self.group))
            if q in recvs and recvs[q].numel():
                ops.append(dist.P2POp(dist.irecv, recvs[q], q, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def exchange(self, h_loc: torch.Tensor) -> None:
        """Fills the halo rows h_loc[n_own:] with the owners' current rows (in place)."""
        if self.world == 1 or h_loc.shape[0] != self.n_loc:
            if h_loc.shape[0] != self.n_loc:
                raise ValueError(f"halo exchange: expected {self.n_loc} rows, got {h_loc.shape[0]}")
            return
        sends = {q: _rows(h_loc, idx) for q, idx in enumerate(self.send_idx) if q != self.rank and idx.numel()}
        recvs = {q: h_loc[self.n_own + self.recv_off[q]: self.n_own + self.recv_off[q + 1]]
                 for q in range(self.world) if q != self.rank and self.recv_counts[q]}
        self._p2p(sends, recvs)

    def reduce_grad(self, dh_loc: torch.Tensor) -> None:
        """Adds the gradients other ranks accumulated on their halo copies of my rows into my owned rows
        (ascending peer order: deterministic) and clears my own halo rows (in place)."""
        if self.world == 1:
            return
        sends = {q: dh_loc[self.n_own + self.recv_off[q]: self.n_own + self.recv_off[q + 1]]
                 for q in range(self.world) if q != self.rank and self.recv_counts[q]}
        recvs = {q: torch.empty((idx.numel(),) + tuple(dh_loc.shape[1:]), dtype=dh_loc.dtype, device=dh_loc.device)
                 for q, idx in enumerate(self.send_idx) if q != self.rank and idx.numel()}
        self._p2p(sends, recvs)
        for q in sorted(recvs):
            # the ids a peer asks for are unique, so index_add_ has no duplicate targets and is deterministic
            if _native(dh_loc):
                from . import ops
                ops.halo_unpack_add(recvs[q], self.send_idx[q], dh_loc)        # cgnn_halo_unpack_add
            else:
                dh_loc.index_add_(0, self.send_idx[q], recvs[q])
        if self.n_halo:
            dh_loc[self.n_own:].zero_()

    def to(self, device):
        self.send_idx = [s.to(device) for s in self.send_idx]
        return self

    def renumbered(self, new_id_of_owned: torch.Tensor) -> "HaloPlan":
        """The same plan for a rank whose OWNED rows have been renumbered (row i now lives at new_id_of_owned[i]; the halo
        rows keep their places behind them): only the rows to send move."""
        import copy
        other = copy.copy(self)
        other.send_idx = [new_id_of_owned[s] if s.numel() else s for s in self.send_idx]
        return other


def sharded_to_device(t: torch.Tensor, particle_dim: int, rank: int, world: int, device, group=None) -> torch.Tensor:
    """A host tensor that every rank holds in full, as a device tensor on every rank -- with each rank copying only ITS share
    of the particles over PCIe (the range slab_bounds gives it along `particle_dim`) and the ranks handing their shares to one
    another over NVLink (`share_over_ranks`).  Host-to-device bytes per rank are 1 / world of the tensor instead of all of it.
    Tensors already on the device, and single-rank runs, pass through."""
    if t is None:
        return None
    device = torch.device(device)
    if t.device == device:
        return t
    if world == 1 or not dist.is_initialized():
        return t.to(device, non_blocking=True)
    out = torch.empty(t.shape, dtype=t.dtype, device=device)
    copy_own_share(out, t, particle_dim, rank, world)
    share_over_ranks(out, particle_dim, world, group)
    return out


def copy_own_share(out: torch.Tensor, t: torch.Tensor, particle_dim: int, rank: int, world: int) -> None:
    """out[..., own share, ...] = t[..., own share, ...], one copy per leading index (frame): every source is then a CONTIGUOUS
    slice of the (pinned) host tensor and goes out as a plain asynchronous DMA -- a strided [W, share, 3] view would be staged
    through pageable memory first."""
    import itertools
    b = slab_bounds(t.shape[particle_dim], world)
    for idx in itertools.product(*[range(s_) for s_ in t.shape[:particle_dim]]):
        sl = idx + (slice(b[rank], b[rank + 1]),)
        out[sl].copy_(t[sl], non_blocking=True)


def share_over_ranks(out: torch.Tensor, particle_dim: int, world: int, group=None) -> None:
    """`out` holds, on every rank, that rank's own share [b[rank], b[rank+1]) along `particle_dim` (b = slab_bounds); on return
    every rank holds all of it.  One broadcast per owner and contiguous block (a [W, N, 3] sequence has one block per frame), so
    the shares need not be equal in size."""
    import itertools
    b = slab_bounds(out.shape[particle_dim], world)
    lead = list(itertools.product(*[range(s) for s in out.shape[:particle_dim]]))
    for q in range(world):
        if b[q + 1] == b[q]:
            continue
        src = dist.get_global_rank(group, q) if group is not None else q
        for idx in lead:
            dist.broadcast(out[idx + (slice(b[q], b[q + 1]),)], src=src, group=group)


def plan_from_global_senders(senders_global: torch.Tensor, bounds: Sequence[int], rank: int, world: int, group=None):
    """Builds the halo plan and the local sender table of one rank.

    senders_global [n_own * k]: global (x-sorted) ids of the senders of this rank's edges.
    Returns (HaloPlan, senders_local int32 [n_own * k], halo_gid int64 [n_halo]); local ids are
    `gid - lo` for owned senders and `n_own + position in halo_gid` for remote ones."""
    dev = senders_global.device
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    n_own, n_global = hi - lo, int(bounds[-1])
    sg = senders_global.long()
    remote = (sg < lo) | (sg >= hi)
    halo_gid = torch.unique(sg[remote], sorted=True)
    b = torch.tensor(list(bounds), dtype=torch.long, device=dev)
    owner = torch.searchsorted(b, halo_gid, right=True) - 1
    recv_counts = torch.bincount(owner, minlength=world)[:world] if halo_gid.numel() else torch.zeros(world, dtype=torch.long, device=dev)
    local = torch.where(remote, n_own + torch.searchsorted(halo_gid, sg), sg - lo).to(torch.int32)

    # tell every owner which of its rows I need: counts first, then the id lists
    if world > 1:
        counts = recv_counts.clone()
        all_counts = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(all_counts, counts, group=group)
        need_from_me = [int(all_counts[q][rank]) for q in range(world)]       # rows peer q wants from me
        off = torch.cat([recv_counts.new_zeros(1), recv_counts.cumsum(0)]).tolist()
        sends = {q: halo_gid[off[q]:off[q + 1]].contiguous() for q in range(world) if q != rank and off[q + 1] > off[q]}
        recvs = {q: torch.empty(need_from_me[q], dtype=torch.long, device=dev) for q in range(world)
                 if q != rank and need_from_me[q] > 0}
        plan = HaloPlan(rank, world, n_own, n_global, [torch.empty(0, dtype=torch.long, device=dev)] * world,
                        recv_counts.tolist(), group)
        plan._p2p(sends, recvs)
        plan.send_idx = [(recvs[q] - lo) if q in recvs else torch.empty(0, dtype=torch.long, device=dev) for q in range(world)]
    else:
        plan = HaloPlan(rank, world, n_own, n_global, [torch.empty(0, dtype=torch.long, device=dev)], [0], group)
    return plan, local, halo_gid


# ------------------------------------------------------------------------------------------------
# loss over a sharded box (train.py:107-118,255-260 with the sums taken over all ranks)
# ------------------------------------------------------------------------------------------------
class _SlabLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, acc, temp, y_acc, y_temp, n_global, dt, w_acc, w_temp, w_mom, group):
        da, dtm = acc - y_acc, temp - y_temp
        part = torch.cat([(da * da).sum().reshape(1), (dtm * dtm).sum().reshape(1), acc.sum(0)])
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
        d = acc.shape[1]
        acc_loss = part[0] / (n_global * d)
        temp_loss = part[1] / n_global
        tot = part[2:] * dt
        mom = w_mom * (tot * tot).sum()
        ctx.save_for_backward(da, dtm, tot)
        ctx.c = (n_global, d, dt, w_acc, w_temp, w_mom)
        return torch.stack([w_acc * acc_loss + w_temp * temp_loss + mom, acc_loss, temp_loss, mom])

    @staticmethod
    def backward(ctx, g):
        da, dtm, tot = ctx.saved_tensors
        n_global, d, dt, w_acc, w_temp, w_mom = ctx.c
        s = g[0]
        d_acc = s * (w_acc * 2.0 / (n_global * d) * da + (w_mom * 2.0 * dt) * tot)
        d_temp = s * (w_temp * 2.0 / n_global) * dtm
        return d_acc, d_temp, None, None, None, None, None, None, None, None


def slab_loss(predictions, graph, dt: float, acc_loss_weight: float = 1.0, temp_rate_loss_weight: float = 1.0,
              momentum_loss_weight: float = 0.0):
    """The reference's training loss for a slab-sharded graph: every rank returns the GLOBAL scalars; backward
    leaves each rank with the partial parameter gradients of its own nodes and edges (all-reduce with SUM)."""
    halo: HaloPlan = graph.halo
    out = _SlabLossFn.apply(predictions["acceleration"], predictions["temp_rate"], graph.y_acc, graph.y_temp_rate,
                            halo.n_global, float(dt), float(acc_loss_weight), float(temp_rate_loss_weight),
                            float(momentum_loss_weight), halo.group)
    det = out.detach()
    return {"loss": out[0], "acc_loss": det[1], "temp_rate_loss": det[2], "momentum_loss": det[3]}
