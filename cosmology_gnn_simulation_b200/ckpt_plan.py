"""Checkpoint schedule of the edge latent stream for message="edge" training.

The backward of processor step t needs e^t, the INPUT of that step (graph_network.py:177-183: e^{t+1} = e^t +
u_e(h^t, e^t)); e^M is never read.  One copy of the stream is E x L x 4 bytes -- 32 GiB at 2 M particles, k = 32,
L = 128 -- so only a few copies fit next to the gradient stream, and the steps in between are recomputed from the
nearest copy below (h^t is kept for every t: node-sized).  e^0 itself is cheap to rebuild from the edge features
(the edge encoder is 4 % of the forward FLOPs), so it need not occupy a buffer.

`schedule(n_steps, n_buf)` returns the action list that visits t = M-1 .. 0 with at most `n_buf` stream buffers
live and the fewest recomputed edge phases (dynamic programme over (segment length, free buffers), the classic
"revolve" recursion):
    ("enc", b)           e^0 = edge encoder(edge_attr) -> buffer b
    ("adv", t, src, dst) e^{t+1} = step t applied to e^t in buffer src -> buffer dst (dst == src: in place)
    ("bwd", t, b)        backward of step t with e^t in buffer b; the buffer is free afterwards
The first time a step is advanced it is the real forward (aggregates, node phase); every later time it is a
recompute of the edge phase alone.  Buffers are small integers < n_buf.
"""
from __future__ import annotations

from functools import lru_cache
from typing import List, Tuple

ENC_COST = 0.5          # cost of rebuilding e^0, in units of one edge phase
INF = float("inf")


@lru_cache(maxsize=None)
def _cost(d: int, f: int, virt: bool) -> Tuple[float, int]:
    """Cheapest way to run the backward of the d steps above a base copy (held in a buffer, or `virt`: rebuilt from
    the encoder on demand) with f free buffers: (edge phases executed, first jump j)."""
    if d == 0:
        return 0.0, 0
    if f == 0:
        return INF, 0
    best, arg = INF, 0
    for j in range(1, d + 1):
        c = j + (ENC_COST if virt else 0.0) + _cost(d - j, f - 1, False)[0] + _cost(j - 1, f, virt)[0]
        if c < best:
            best, arg = c, j
    return best, arg


def schedule(n_steps: int, n_buf: int) -> List[tuple]:
    """Action list for M = n_steps processor steps with n_buf >= 1 stream buffers (excluding the gradient stream)."""
    if n_steps < 1 or n_buf < 1:
        raise ValueError("schedule needs n_steps >= 1 and n_buf >= 1")
    acts: List[tuple] = []
    free = list(range(n_buf - 1, -1, -1))

    def emit(a: int, d: int, virt: bool, base: int):
        if d == 0:
            return
        j = _cost(d, len(free), virt)[1]
        b = free.pop()
        if virt:
            acts.append(("enc", b))
            acts.append(("adv", a, b, b))
        else:
            acts.append(("adv", a, base, b))
        for t in range(a + 1, a + j):
            acts.append(("adv", t, b, b))
        emit(a + j, d - j, False, b)
        acts.append(("bwd", a + j, b))
        free.append(b)
        emit(a, j - 1, virt, base)

    d = n_steps - 1
    held = ENC_COST + _cost(d, n_buf - 1, False)[0] if n_buf >= 2 or d == 0 else INF
    virt = _cost(d, n_buf, True)[0] + ENC_COST
    if held <= virt:                         # e^0 keeps a buffer of its own
        b0 = free.pop()
        acts.append(("enc", b0))
        emit(0, d, False, b0)
        acts.append(("bwd", 0, b0))
    else:                                    # e^0 is rebuilt whenever a sweep starts from it
        emit(0, d, True, -1)
        b0 = free.pop()
        acts.append(("enc", b0))
        acts.append(("bwd", 0, b0))
    return acts


def recomputed_phases(acts: List[tuple]) -> float:
    """Edge phases (encoder rebuilds count ENC_COST) the schedule executes beyond the plain forward."""
    seen, extra, enc = set(), 0.0, 0
    for a in acts:
        if a[0] == "adv":
            if a[1] in seen:
                extra += 1.0
            seen.add(a[1])
        elif a[0] == "enc":
            enc += 1
    return extra + ENC_COST * max(enc - 1, 0)


def split_forward(acts: List[tuple]) -> int:
    """Index of the first backward action: everything before it is the forward sweep."""
    for i, a in enumerate(acts):
        if a[0] == "bwd":
            return i
    return len(acts)
