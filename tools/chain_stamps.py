#!/usr/bin/env python
"""Stage timeline of the tensor-core chain kernel (cgnn_debug_stamps): runs the processor edge phase forward
and backward at a BASELINE config size and prints, per chain launch, the median cycles block 0's epilogue
group 0 spends between the stamps of a tile (clock64; see CGNN_STAMP in csrc/mp_tc.cu).

  python tools/chain_stamps.py [--n 32768 --k 16 --tiles 12]
"""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from cosmology_gnn_simulation_b200 import ops  # noqa: E402
from cosmology_gnn_simulation_b200._lib import lib  # noqa: E402
from cosmology_gnn_simulation_b200.ops import MlpParams  # noqa: E402

SLOTS = ["start", "in_conv", "publish0", "mma1_done", "ps_ready", "ep1_math", "publish1", "mma2_done", "(ps)", "ep2_math",
         "publish2", "mma_last", "ln_stats", "stored", "-", "-"]

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=32768)
ap.add_argument("--k", type=int, default=16)
ap.add_argument("--tiles", type=int, default=12)
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--bwd", action="store_true")
ap.add_argument("--no-agg", action="store_true", help="forward without the per-receiver sum")
ap.add_argument("--local-senders", action="store_true", help="sender = receiver: the gathered P_s row is shared by the k lanes of a receiver")
a = ap.parse_args()
L = 128
d = torch.device("cuda", 0)
g = torch.Generator(device=d).manual_seed(0)
ws = [torch.randn(L, i, device=d, generator=g) / i ** 0.5 for i in (3 * L, L, L)]
bs = [torch.randn(L, device=d, generator=g) * 0.1 for _ in range(3)]
p = MlpParams(ws, bs, torch.ones(L, device=d), torch.zeros(L, device=d))
n, k = a.n, a.k
h = torch.randn(n, L, device=d, generator=g)
e = torch.randn(n * k, L, device=d, generator=g)
senders = torch.randint(0, n, (n * k,), device=d, generator=g, dtype=torch.int32)
if a.local_senders:
    senders = torch.arange(n, device=d, dtype=torch.int32).repeat_interleave(k)
e_out = torch.empty_like(e)
agg = torch.empty_like(h)


def run():
    if a.bwd:
        rowptr, perm = ops.csr_transpose(senders, n)
        de_next = torch.randn_like(e)
        dagg = torch.randn_like(h)
        de = torch.empty_like(e)
        dh = torch.zeros_like(h)
        gs = torch.empty_like(e)
        ops.mp_edge_bwd(p, h, e, senders, rowptr, perm, k, de_next, dagg, de, dh, gs, a.precision)
    else:
        ops.mp_edge_fwd(p, h, e, senders, k, e_out, None if a.no_agg else agg, a.precision)


run()                                   # warm-up (module load, workspaces)
torch.cuda.synchronize()
LAUNCHES = 16
buf = torch.zeros(LAUNCHES * 2 * a.tiles * 16, dtype=torch.int64, device=d)
lib().cgnn_debug_stamps(ctypes.c_void_p(buf.data_ptr()), a.tiles, LAUNCHES)
s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
run()
t.record()
torch.cuda.synchronize()
lib().cgnn_debug_stamps(None, 0, 0)
print(f"{'bwd' if a.bwd else 'fwd'} n={n} k={k} {a.precision}: {s.elapsed_time(t):.3f} ms for the call")
st = buf.cpu().view(LAUNCHES, 2, a.tiles, 16)
for li in range(LAUNCHES):
    x = st[li, 0]                       # group 0
    if int(x.max()) == 0:
        continue
    rows = x[x[:, 0] > 0]
    if rows.shape[0] < 3:
        print(f"launch {li}: {rows.shape[0]} stamped tiles (small launch)")
        continue
    rows = rows[1:]                      # skip the first tile (pipeline fill)
    used = [j for j in range(14) if int(rows[:, j].min()) > 0]
    per_tile = (rows[1:, 0] - rows[:-1, 0]).float().median().item() if rows.shape[0] > 1 else float("nan")
    parts = []
    for a_, b_ in zip(used[:-1], used[1:]):
        parts.append(f"{SLOTS[b_]} {int((rows[:, b_] - rows[:, a_]).float().median().item())}")
    total = (rows[:, used[-1]] - rows[:, used[0]]).float().median().item()
    print(f"launch {li}: tile {int(total)} cyc in group, {int(per_tile)} cyc tile-to-tile | " + " | ".join(parts))
