#!/usr/bin/env python
"""Headline benchmark: particle-steps/s of one training application (forward + loss + backward) of
the encode / M x message-passing / decode Interaction Network on a synthetic periodic box.

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # N ranks, one box replica each + gradient all-reduce

Workload (default `config3`): BASELINE.json configs[2], the configuration the metric is quoted on -- 128^3 =
2 097 152 particles, k=32, latent 128, 10 MP steps, acceleration + temperature-rate + momentum loss, edge
messages, on ONE B200.  `--workload config2` is configs[1] (32^3 particles, k=16).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (particles, k, latent, mp_steps, box kind)
    "config1": (4096, 16, 64, 5, "uniform"),
    "config2": (32 ** 3, 16, 128, 10, "uniform"),
    "config3": (128 ** 3, 32, 128, 10, "uniform"),
    "config4": (128 ** 3, 16, 128, 10, "uniform"),       # --rollout: per-GPU share of the 256^3 rollout box at 8 GPUs (k = 16, render_rollout.py:49)
    "config2-clustered": (32 ** 3, 16, 128, 10, "clustered"),
    "config3-clustered": (128 ** 3, 32, 128, 10, "clustered"),
    "tiny": (2048, 8, 32, 2, "uniform"),
}
W_ACC, W_TEMP, W_MOM = 1.0, 1.0, 0.1
METRIC = "particle-steps/sec (fwd+bwd, 10 MP, L=128, k=32) at 1/2/4/8 B200; % roofline"      # BASELINE.json, = config3's shape
UNIT = "particle-steps/s"
GRAPH_MAX_EDGES = 8 << 20         # above this a step's kernels are long enough that replaying a captured step buys nothing


def metric_name(workload):
    """BASELINE.json's metric string; for the other workloads the same wording with their own shape."""
    n, k, L, M, kind = WORKLOADS[workload]
    return f"particle-steps/sec (fwd+bwd, {M} MP, L={L}, k={k}) at 1/2/4/8 B200; % roofline"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_sustained": d["bf16_tflops_sustained"], "bf16_burst": d["bf16_tflops"],
                "hbm": d["hbm_gbs"], "source": "MEASURED_PEAKS.json"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def flops_edge_fwd(n, k, L):
    """Algorithmic FLOPs of one processor edge phase (proc_edge = 2E(3LH + H^2 + HL), H = L)."""
    return 2.0 * n * k * (3 * L * L + L * L + L * L)


def flops_application(n, k, L, M, message):
    fn, fe = 17, 4
    enc_n = 2.0 * n * (fn * L + L * L + L * L)
    enc_e = 2.0 * n * k * (fe * L + L * L + L * L)
    proc_e = flops_edge_fwd(n, k, L)
    proc_n = 2.0 * n * (2 * L * L + L * L + L * L)
    dec = 2.0 * n * (L * L + L * L + 3 * L) + 2.0 * n * (L * L + L * L + L)
    fwd = enc_n + enc_e + M * (proc_e + proc_n) + dec
    if message == "edge":
        return 3.0 * fwd
    return fwd + 2.0 * (enc_n + M * proc_n + dec)


def phase_work(n, k, L, M, message):
    """Algorithmic FLOPs and HBM bytes of the forward and the backward of one training application (SURVEY §8d formulas,
    FP32 latents b = 4, the fused design's traffic: only latents and indices touch HBM; recompute is not useful work)."""
    e, b = n * k, 4
    fn, fe = 17, 4
    enc_n = 2.0 * n * (fn * L + L * L + L * L)
    enc_e = 2.0 * e * (fe * L + L * L + L * L)
    proc_e = flops_edge_fwd(n, k, L)
    proc_n = 2.0 * n * (2 * L * L + L * L + L * L)
    dec = 2.0 * n * (L * L + L * L + 3 * L) + 2.0 * n * (L * L + L * L + L)
    fwd_flop = enc_n + enc_e + M * (proc_e + proc_n) + dec
    bwd_flop = 2.0 * fwd_flop if message == "edge" else 2.0 * (enc_n + M * proc_n + dec)
    enc_bytes = fn * 4 * n + fe * 4 * e + n * L * b + e * L * b
    dec_bytes = n * L * b + 16 * n
    fwd_bytes = enc_bytes + M * (2 * e * L * b + 2 * n * L * b + 4 * e) + dec_bytes
    bwd_bytes = enc_bytes + dec_bytes + M * ((3 * e * L * b + 3 * n * L * b + 8 * e) if message == "edge" else (3 * n * L * b + 8 * e))
    return {"forward": (fwd_flop, float(fwd_bytes)), "backward": (bwd_flop, float(bwd_bytes))}


# ------------------------------------------------------------------------------------------------
# clocks (sampled DURING the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _loop(self):
        nv = self._nv
        names = {"hw_slowdown": "nvmlClocksThrottleReasonHwSlowdown",
                 "hw_thermal_slowdown": "nvmlClocksThrottleReasonHwThermalSlowdown",
                 "sw_thermal_slowdown": "nvmlClocksThrottleReasonSwThermalSlowdown",
                 "sw_power_cap": "nvmlClocksThrottleReasonSwPowerCap"}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for label, attr in names.items():
                    if mask & getattr(nv, attr, 0):
                        self.reasons.add(label)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# CPU reference leg (oracle port of the reference algorithm; the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------
def cpu_training_step_rate(k, L, M, steps, warmup, sample_n=8192, seed=0, message="edge", keep=False):
    """Times the reference algorithm (gathers, cat, Linear, ReLU, LayerNorm, index_add_, autograd backward;
    `message` as the GPU arm) on the host cores with all threads.
    The particle count is a bounded sample of the workload (cost is linear in N at fixed k, L, M)."""
    from cosmology_gnn_simulation_b200 import synthetic
    from oracle import knn_ref, model_ref
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    pos = synthetic.positions(sample_n, "uniform", 1.0, seed=seed)
    t0 = time.perf_counter()
    ext = knn_ref.knn_kdtree(pos, 1.0, k)                   # KD-tree over the 27N ghosts, one thread (F5)
    knn_s = time.perf_counter() - t0
    ei = torch.from_numpy(knn_ref.edge_index_from_ext(ext, sample_n))
    p = torch.from_numpy(pos)
    d = p[ei[0]] - p[ei[1]]
    ea = torch.cat([d, d.norm(dim=-1, keepdim=True)], dim=-1)
    gen = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(sample_n, 17, generator=gen)
    ya, yt = torch.randn(sample_n, 3, generator=gen), torch.randn(sample_n, 1, generator=gen)
    params = {k_: v.requires_grad_(True) for k_, v in model_ref.init_params(L, L, 2, M, 3, seed=seed).items()}
    times = []
    for it in range(warmup + steps):
        for v in params.values():
            v.grad = None
        t0 = time.perf_counter()
        o = model_ref.forward(params, x, ei, ea, 2, M, message=message)
        ls = model_ref.loss(o["acceleration"], o["temp_rate"], ya, yt, 0.01, w_acc=W_ACC, w_temp=W_TEMP, w_mom=W_MOM)
        ls["loss"].backward()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    out = {"rate": sample_n * len(times) / total, "ms_per_step": 1e3 * total / len(times), "cores": cores,
           "sample_n": sample_n, "knn_rate": sample_n / knn_s, "knn_s": knn_s}
    if keep:                                               # the sample and the oracle's result on it, for the parity block
        out["sample"] = {"x": x, "ei": ei, "ea": ea, "ya": ya, "yt": yt, "params": params, "loss": ls["loss"].detach(),
                         "acc": o["acceleration"].detach(), "temp": o["temp_rate"].detach()}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                              # rank 0 alone runs the CPU arm
    n, k, L, M, kind = WORKLOADS[args.workload]
    sample_n = min(n, 8192)
    r = cpu_training_step_rate(k, L, M, args.steps, max(args.warmup, 1), sample_n=sample_n, message=args.message)
    sample = (f"{sample_n} of {n} particles per step at the workload's k={k}, L={L}, M={M} "
              f"(cost linear in N); graph build (KD-tree on 27N ghosts, 1 thread) timed apart: "
              f"{r['knn_rate']:.3g} particles/s")
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": r["rate"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, args.message, args.precision, args.gpus, args.sharding),
        "cpu_baseline": {"value": r["rate"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample,
                         "arithmetic": "fp32 torch CPU ops (the reference's own), all host threads"},
        "e2e": {"value": r["rate"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(name, message, precision, gpus, sharding="slab", scaling="weak"):
    n, k, L, M, kind = WORKLOADS[name]
    if gpus == 1:
        par = "single GPU"
    elif sharding == "slab" and scaling == "strong":
        par = (f"slab{gpus}, strong: the workload's ONE box of {n} particles cut into {gpus} equal-count x-slabs; per MP step one "
               f"halo exchange of boundary latents (NCCL P2P), 5-float loss all-reduce, gradient all-reduce(SUM)")
        n = n // gpus
    elif sharding == "slab":
        par = (f"slab{gpus}: ONE box of {gpus}x{n} particles cut into {gpus} equal-count x-slabs; per MP step one halo "
               f"exchange of boundary latents (NCCL P2P), 5-float loss all-reduce, gradient all-reduce(SUM)")
    else:
        par = f"dp{gpus} (one box replica per GPU, gradient all-reduce)"
    return {"workload": f"{name}: training step, {n} particles per GPU ({kind} periodic box), k={k}, latent={L}, "
                        f"{M} MP steps, acc+temp+momentum loss",
            "particles_per_gpu": n, "k": k, "latent": L, "mp_steps": M, "message": message, "precision": precision,
            "parallelism": par,
            "l2": "inputs larger than L2 (edge latent stream alone exceeds 126 MB)" if n * k * L * 4 > 126e6
                  else "L2 flushed between timed steps (256 MB write)"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    from cosmology_gnn_simulation_b200 import _lib, ops, synthetic
    from cosmology_gnn_simulation_b200 import distributed as cd
    from cosmology_gnn_simulation_b200.data_utils import preprocess, preprocess_slab
    from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode
    from cosmology_gnn_simulation_b200.loss import combined_loss
    from cosmology_gnn_simulation_b200.slab import slab_loss
    import torch.distributed as dist

    rank, world, local = cd.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    n, k, L, M, kind = WORKLOADS[args.workload]
    message, precision = args.message, args.precision

    slab = world > 1 and args.sharding == "slab"
    # slab: every rank holds the same box of world * n particles and owns one x-slab of it (weak scaling);
    # replica: every rank has its own box of n particles
    strong = slab and args.scaling == "strong"        # strong: ONE box of n particles cut into `world` slabs (a box too large for one GPU)
    n_box = n if strong else n * world
    box = synthetic.make_box(n_box, kind, seed=0) if slab else synthetic.make_box(n, kind, seed=rank)
    if strong:
        n = n_box // world                            # particles per GPU, for the throughput count below
    md = box["metadata"]
    coords_host = box["Coordinates"].pin_memory()
    energy_host = box["InternalEnergy"].pin_memory()
    h2d_bytes = coords_host.numel() * 4 + energy_host.numel() * 4
    if slab and world > 1:
        h2d_bytes = (h2d_bytes + world - 1) // world      # per rank: its own share of the box (the rest arrives over NVLink)

    torch.manual_seed(0)
    model = EncodeProcessDecode(L, L, 2, M, 3, message=message, precision=precision).to(dev)
    bucket = cd.GradientBucket(model.parameters())
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev) if n * k * L * 4 <= 126e6 else None

    def build_graph(coords, energy):
        if slab:
            return preprocess_slab(coords[:5], energy[:5], md, coords[5:6], energy[5:6], noise_std=0.0, num_neighbors=k,
                                   dt=md["dt"], box_size=md["box_size"], rank=rank, world=world, device=dev)
        return preprocess(coords[:5], energy[:5], md, coords[5:6], energy[5:6], noise_std=0.0, num_neighbors=k,
                          dt=md["dt"], box_size=md["box_size"], device=dev)

    graph = build_graph(coords_host, energy_host)        # resident inputs for the device-timed region

    PHASE_EVENTS = {"on": False, "marks": []}         # (start, after forward + loss, after backward) CUDA events of eager steps

    def _mark():
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    def train_step(g):
        for p in model.parameters():
            p.grad = None
        e0 = _mark() if PHASE_EVENTS["on"] else None
        pred = model(g)
        ls = slab_loss(pred, g, md["dt"], W_ACC, W_TEMP, W_MOM) if slab else combined_loss(pred, g, md["dt"], W_ACC, W_TEMP, W_MOM)
        e1 = _mark() if PHASE_EVENTS["on"] else None
        ls["loss"].backward()
        e2 = _mark() if PHASE_EVENTS["on"] else None
        # slab: every rank holds the partial gradient of ONE global loss (SUM); replicas average
        bucket.all_reduce(average=not slab)
        if PHASE_EVENTS["on"]:
            PHASE_EVENTS["marks"].append((e0, e1, e2))
        return ls

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # One GPU, small boxes: the ~2000 launches of a step are captured once into a CUDA graph and replayed (graphed.py);
    # the capture has to come before any eager step on the default stream.  Large boxes (kernels of milliseconds) and
    # N > 1 (NCCL point-to-point inside the step) run eagerly.
    graphed = None
    launches_per_graphed_step = 0
    if world == 1 and not args.no_cuda_graph and n * k <= GRAPH_MAX_EDGES:
        from cosmology_gnn_simulation_b200.graphed import GraphedTrainStep
        graphed = GraphedTrainStep(model, lambda pred, g: combined_loss(pred, g, md["dt"], W_ACC, W_TEMP, W_MOM))
        graphed.capture(graph)
        launches_per_graphed_step = graphed.launches_per_step
    run_step = graphed if graphed is not None else train_step

    def timed_loop(step_fn, collect):
        """`collect`: eager steps -- bracket every cgnn_mp_edge_fwd call and the forward / backward phases with CUDA events
        (event records only; nothing is synchronised inside the loop)."""
        for _ in range(args.warmup):
            step_fn(graph)
        barrier()
        ops.EDGE_FWD_EVENTS = [] if collect else None
        PHASE_EVENTS["on"], PHASE_EVENTS["marks"] = collect, []
        l0 = _lib.launch_count()
        marks = []
        with ClockSampler(local) as clk:
            for _ in range(args.steps):
                if flush is not None:
                    flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                step_fn(graph)
                e.record()
                marks.append((s, e))
            barrier()
        ev, ops.EDGE_FWD_EVENTS = ops.EDGE_FWD_EVENTS, None
        PHASE_EVENTS["on"] = False
        ms = cd.max_over_ranks(sum(s.elapsed_time(e) for s, e in marks), dev)
        return ms, _lib.launch_count() - l0, ev, clk

    # ---- device-resident throughput (`value`) -------------------------------------------------
    total_ms, launches, events, clocks = timed_loop(run_step, graphed is None)
    if graphed is not None:
        launches = launches_per_graphed_step * args.steps
    value = world * n * args.steps / (total_ms * 1e-3)
    phase_marks = PHASE_EVENTS["marks"]

    # ---- the dominant kernel, bracketed by CUDA events.  Launches inside a replayed CUDA graph cannot be bracketed one
    # by one, so a graphed run adds an eager replica of the timed region; an eager run measured them in the region itself
    eager_ms = None
    if graphed is not None:
        eager_ms, _, events, _ = timed_loop(train_step, True)
        phase_marks = PHASE_EVENTS["marks"]
    edge_ms = sum(a.elapsed_time(b) for a, b in events) / max(len(events), 1)

    # ---- graph build alone (reported apart; not part of the model application) -----
    pos_dev = graph.pos.contiguous()
    n_pos = pos_dev.shape[0]
    del graph
    if not slab:
        for _ in range(2):
            ops.knn_periodic(pos_dev, md["box_size"], k)
        torch.cuda.synchronize(dev)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        reps = 5
        for _ in range(reps):
            nbr = ops.knn_periodic(pos_dev, md["box_size"], k)
            feats = ops.edge_features(pos_dev, nbr, md["box_size"], want_edge_index=False)
        e.record()
        torch.cuda.synchronize(dev)
        knn_ms = s.elapsed_time(e) / reps
        del nbr, feats
    else:
        knn_ms = None

    # ---- end to end through the public API with host buffers (`e2e`) ---------------------------
    def e2e_step():
        if slab and world > 1:
            # host tensors go to preprocess_slab as they are: every rank copies its own share of the particles over PCIe, the
            # shares travel between the GPUs over NVLink (slab.sharded_to_device)
            c, u = coords_host, energy_host
        else:
            c = coords_host.to(dev, non_blocking=True)
            u = energy_host.to(dev, non_blocking=True)
        g = build_graph(c, u)
        ls = run_step(g)
        return torch.stack([ls["loss"].detach(), ls["acc_loss"], ls["temp_rate_loss"], ls["momentum_loss"]]).cpu()

    # (steps of seconds: a few are enough, the default run has to end within minutes)
    e2e_steps = args.steps if total_ms / args.steps < 500.0 else min(args.steps, 4)
    for _ in range(min(args.warmup, 3) if e2e_steps == args.steps else 1):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        last = e2e_step()
    barrier()
    e2e_s = cd.max_over_ranks(time.perf_counter() - t0, dev)
    e2e_value = world * n * e2e_steps / e2e_s

    # ---- N > 1: the sharded step against the single-GPU step on a small box (the world-2 pytest cannot run on the
    # driver's one-GPU test lease, so the check travels with the bench line) -----------------------------------
    parity = slab_parity(rank, world, dev, k, L, M, message, precision) if slab else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    fl = flops_edge_fwd(n, k, L)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "edge_fwd_traffic.json")     # dram bytes of the same launch from `ncu --set full`
    if os.path.exists(tpath):
        with open(tpath) as f:
            tt = json.load(f)
        for t in (tt if isinstance(tt, list) else [tt]):
            if t.get("workload") == args.workload and t.get("precision") == precision:
                traffic = t.get("dram_bytes_per_launch")
                traffic_src = "profiles/edge_fwd_traffic.json: " + t.get("source", "ncu --set full capture of this kernel at this workload (not measured in this run)")
    achieved = fl / (edge_ms * 1e-3) / 1e12 if edge_ms > 0 else 0.0
    step_ms = (eager_ms if eager_ms is not None else total_ms) / args.steps
    line = {
        "metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16": "bf16",
                  "bf16x3": "bf16x3 (split bf16 operands, f32 accumulate, f32 latents in HBM" +
                            ("; gradient streams of the edge MLPs' backward bf16)" if os.environ.get("CGNN_GRAD_STREAM", "bf16") == "bf16"
                             else ")")}[precision],
        "data": "synthetic",
        "config": workload_config(args.workload, message, precision, world, args.sharding, args.scaling),
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 16,
                "steps": e2e_steps,
                "includes": "H2D of the 6 frames, graph build (k-NN + features), forward, loss, backward, D2H of the 4 loss scalars"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "cgnn_mp_edge_fwd (processor edge phase, forward)",
                     "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                     "executed_flop_per_launch": (2.0 * n * k * 3 * L * L + 2.0 * n * 2 * L * L) * (3 if precision == "bf16x3" else 1)
                     if precision != "fp32" else fl,
                     "flop_per_launch": fl, "ms_per_launch": edge_ms, "launches_timed": len(events),
                     "peak_source": peaks["source"] + " (bf16 dense, sustained)",
                     "share_of_step": edge_ms * len(events) / max(args.steps, 1) / step_ms,
                     "timed": "eager replica of the timed region, same process" if eager_ms is not None else "the timed region"},
        "cuda_graph": graphed is not None,
        "eager_ms_per_step": None if eager_ms is None else eager_ms / args.steps,
        "model_tflops": flops_application(n, k, L, M, message) / (total_ms / args.steps * 1e-3) / 1e12,
        "graph_build": None if knn_ms is None else {"ms": knn_ms, "particles_per_s": n_pos / (knn_ms * 1e-3),
                                                    "what": "cgnn_knn_periodic + cgnn_edge_features, device resident"},
        "loss_check": [float(v) for v in last],
        "peak_memory_gib": {"allocated": torch.cuda.max_memory_allocated(dev) / 2**30, "reserved": torch.cuda.max_memory_reserved(dev) / 2**30,
                            "device": torch.cuda.mem_get_info(dev)[1] / 2**30},
    }
    if message == "edge" and precision != "fp32":
        from cosmology_gnn_simulation_b200 import graph_network as gn
        plans = [v for kk, v in gn._STREAM_PLANS.items() if kk[1] == n and kk[3] == n * k]       # (the parity block plans a small box too)
        forced = int(os.environ.get("CGNN_EDGE_BUFFERS", "0"))
        if plans or forced:
            from cosmology_gnn_simulation_b200 import ckpt_plan
            nb = min(forced, M) if forced else plans[-1]
            line["edge_stream"] = {"buffers": nb, "bytes_per_copy": n * k * L * 4,
                                   "recomputed_edge_phases_per_step": ckpt_plan.recomputed_phases(ckpt_plan.schedule(M, nb)),
                                   "what": "copies of the FP32 edge latent stream kept for the backward (ckpt_plan.py); the rest is recomputed"}
    if parity is not None:
        line["parity"] = parity
    if phase_marks:
        marks = phase_marks[-args.steps:]
        work = phase_work(n, k, L, M, message)
        phases = {}
        for name, (a, b_) in (("forward", (0, 1)), ("backward", (1, 2))):
            ms = sum(m[a].elapsed_time(m[b_]) for m in marks) / len(marks)
            fl_, by = work[name]
            phases[name] = {"ms": ms, "tflops": fl_ / (ms * 1e-3) / 1e12, "frac_tensor": fl_ / (ms * 1e-3) / 1e12 / peaks["bf16_sustained"],
                            "gbs": by / (ms * 1e-3) / 1e9, "frac_hbm": by / (ms * 1e-3) / 1e9 / peaks["hbm"]}
        phases["what"] = ("eager steps of the same process, CUDA events around forward + loss and around backward; algorithmic FLOPs and "
                          "HBM bytes of SURVEY 8d (FP32 latents, fused-design traffic) against the measured BF16 and copy peaks")
        line["phases"] = phases
    if not args.no_cpu_baseline and world == 1:
        torch.cuda.empty_cache()
        line["cpu_baseline"] = cpu_baseline_leg(k, L, M, n, message, precision, dev)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_leg(k, L, M, n, message, precision, dev):
    """The oracle port timed on the host cores on a bounded sample of the workload -- and, as the checker it is, the CUDA
    path run on that very sample (same graph, features, targets, weights): loss and outputs side by side."""
    from cosmology_gnn_simulation_b200.graph import Data
    from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode
    from cosmology_gnn_simulation_b200.loss import combined_loss
    r = cpu_training_step_rate(k, L, M, steps=2, warmup=1, sample_n=min(n, 8192), message=message, keep=True)
    s = r.pop("sample")
    model = EncodeProcessDecode(L, L, 2, M, 3, message=message, precision=precision)
    model.load_state_dict({k_: v.detach() for k_, v in s["params"].items()})
    model = model.to(dev)
    g = Data(x=s["x"].to(dev), edge_index=s["ei"].to(dev), edge_attr=s["ea"].to(dev), y_acc=s["ya"].to(dev), y_temp_rate=s["yt"].to(dev))
    pred = model(g)
    ls = combined_loss(pred, g, 0.01, W_ACC, W_TEMP, W_MOM)
    ls["loss"].backward()

    def rel(a, b):
        return float((a.detach().double().cpu() - b.detach().double()).norm() / b.detach().double().norm())
    grad_err = max(rel(prm.grad, s["params"][name].grad) for name, prm in model.named_parameters()
                   if prm.grad is not None and s["params"][name].grad is not None)
    return {"value": r["rate"], "unit": UNIT, "cores": r["cores"], "kind": "port",
            "sample": f"oracle port (message='{message}'), {r['sample_n']} of {n} particles per step, "
                      f"k={k}, L={L}, M={M}, fwd+loss+bwd, 1 warm-up + 2 timed steps; k-NN oracle "
                      f"(KD-tree on 27N ghosts, 1 thread) {r['knn_rate']:.3g} particles/s",
            "parity_on_sample": {"loss_cuda": float(ls["loss"].detach()), "loss_oracle_fp32": float(s["loss"]),
                                 "loss_rel": abs(float(ls["loss"].detach()) - float(s["loss"])) / abs(float(s["loss"])),
                                 "acceleration_rel_l2": rel(pred["acceleration"], s["acc"]),
                                 "temp_rate_rel_l2": rel(pred["temp_rate"], s["temp"]),
                                 "worst_gradient_rel_l2": grad_err,
                                 "what": "this repo's CUDA path on the oracle's sample (same graph, inputs, weights) against the "
                                         "oracle's float32 result; gradients carry the ReLU-gate noise of two fp32-class "
                                         "implementations (DESIGN.md section 2)"}}


def slab_parity(rank, world, dev, k, L, M, message, precision, n_small=8192):
    """Runs one training application of a small box slab-sharded over all ranks and again on rank 0 alone; returns the
    relative differences of loss, outputs and parameter gradients (rank 0; None elsewhere)."""
    import torch.distributed as dist
    from cosmology_gnn_simulation_b200 import distributed as cd, synthetic
    from cosmology_gnn_simulation_b200.data_utils import preprocess, preprocess_slab
    from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode
    from cosmology_gnn_simulation_b200.loss import combined_loss
    from cosmology_gnn_simulation_b200.slab import slab_loss
    n_small -= n_small % world
    box = synthetic.make_box(n_small, "uniform", seed=123)
    md = box["metadata"]
    args_ = (box["Coordinates"][:5], box["InternalEnergy"][:5], md, box["Coordinates"][5:6], box["InternalEnergy"][5:6])
    kw = dict(noise_std=0.0, num_neighbors=k, dt=md["dt"], box_size=md["box_size"], device=dev)
    torch.manual_seed(1234)
    model = EncodeProcessDecode(L, L, 2, M, 3, message=message, precision=precision).to(dev)
    g = preprocess_slab(*[a.clone() if torch.is_tensor(a) else a for a in args_], rank=rank, world=world, **kw)
    torch.manual_seed(1234)                             # the lazy first layers draw their weights here: same on every rank
    pred = model(g)
    ls = slab_loss(pred, g, md["dt"], W_ACC, W_TEMP, W_MOM)
    ls["loss"].backward()
    cd.GradientBucket(model.parameters()).all_reduce(average=False)
    own = torch.cat([pred["acceleration"].detach(), pred["temp_rate"].detach()], dim=1).contiguous()
    parts = [torch.empty_like(own) for _ in range(world)]
    dist.all_gather(parts, own)
    torch.cuda.synchronize(dev)
    if rank != 0:
        dist.barrier()
        return None
    out_sorted = torch.cat(parts, dim=0)
    out_slab = torch.empty_like(out_sorted)
    out_slab[g.order] = out_sorted                       # back to the input order
    single = EncodeProcessDecode(L, L, 2, M, 3, message=message, precision=precision)
    single.load_state_dict(model.state_dict())
    single = single.to(dev)
    g1 = preprocess(*[a.clone() if torch.is_tensor(a) else a for a in args_], **kw)
    pred1 = single(g1)
    ls1 = combined_loss(pred1, g1, md["dt"], W_ACC, W_TEMP, W_MOM)
    ls1["loss"].backward()
    torch.cuda.synchronize(dev)

    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm())
    out1 = torch.cat([pred1["acceleration"].detach(), pred1["temp_rate"].detach()], dim=1)
    grads = [(rel(p.grad, q.grad), nm) for (nm, p), q in zip(model.named_parameters(), single.parameters())
             if p.grad is not None and q.grad is not None]
    worst = max(grads)
    res = {"box": f"{n_small} particles, k={k}, L={L}, M={M}, message={message}, precision={precision}: {world}-rank slab step vs the "
                  f"same box on rank 0 alone",
           "loss_rel": abs(float(ls["loss"].detach()) - float(ls1["loss"].detach())) / abs(float(ls1["loss"].detach())),
           "outputs_rel_l2": rel(out_slab, out1), "worst_gradient_rel_l2": worst[0], "worst_gradient": worst[1],
           "gradients_compared": len(grads),
           "note": "sums over a receiver's k edges and over rows are taken in the same order on both sides; the sender scatter "
                   "adds halo contributions in peer order, so gradients agree to rounding, not bitwise"}
    dist.barrier()
    return res


def run_rollout(args):
    """Inference rollout (BASELINE configs[3]): per step the k-NN graph is rebuilt, the model runs forward only and the
    integrator advances the box, all device resident (rollout.py).  N > 1: ONE box of N x the workload's particles,
    re-partitioned into x-slabs every step (rollout_slab: particles change owner as they move), one all-gather of the new
    frame per step."""
    from cosmology_gnn_simulation_b200 import _lib, synthetic
    from cosmology_gnn_simulation_b200 import distributed as cd
    from cosmology_gnn_simulation_b200.graph_network import EncodeProcessDecode
    from cosmology_gnn_simulation_b200.rollout import rollout, rollout_slab
    import torch.distributed as dist
    rank, world, local = cd.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    n, k, L, M, kind = WORKLOADS[args.workload]
    box = synthetic.make_box(n * world, kind, seed=0)
    md = box["metadata"]
    torch.manual_seed(0)
    model = EncodeProcessDecode(L, L, 2, M, 3, message=args.message, precision=args.precision).to(dev)
    data = {"Coordinates": box["Coordinates"][:6].to(dev), "InternalEnergy": box["InternalEnergy"][:6].to(dev)}

    def run(steps):
        if world == 1:
            return rollout(model, data, md, 0.0, md["dt"], md["box_size"], window_size=5, num_neighbors=k, n_steps=steps)
        return rollout_slab(model, data, md, 0.0, md["dt"], md["box_size"], window_size=5, num_neighbors=k, n_steps=steps,
                            rank=rank, world=world)

    run(max(args.warmup, 3))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    l0 = _lib.launch_count()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        s.record()
        out = run(args.steps)
        e.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
    ms = cd.max_over_ranks(s.elapsed_time(e), dev)
    if rank == 0:
        line = {"metric": "particle-steps/sec of an inference rollout (k-NN rebuild + forward + integrator per step)",
                "value": n * world * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.precision, "data": "synthetic",
                "config": {"workload": f"rollout of {args.workload}: {n} particles per GPU, ONE box of {n * world}, k={k}, latent={L}, {M} MP steps, "
                                       f"graph rebuilt every step, device-resident trajectory"
                                       + (", re-partitioned into x-slabs every step (migration), one all-gather of the new frame per step" if world > 1 else ""),
                           "message": args.message},
                "clocks": clocks.summary(), "gpu_launches": int(_lib.launch_count() - l0),
                "finite": bool(torch.isfinite(out["Coordinates"]).all())}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cgnn", choices=["cgnn", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--message", default="edge", choices=["sender", "edge"],
                    help="edge: the Interaction Network of the north star (default); sender: what PyG's default message() computes")
    ap.add_argument("--precision", default="bf16x3", choices=["fp32", "bf16x3", "bf16"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1, slab sharding: weak = one box of N x the workload's particles (default), strong = the workload's box itself cut into N slabs")
    ap.add_argument("--sharding", default="slab", choices=["slab", "replica"],
                    help="N > 1: slab = one box of N x particles cut into x-slabs with halo exchange (default); "
                         "replica = one independent box per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true", help="one GPU: issue every launch eagerly instead of replaying a captured step")
    ap.add_argument("--rollout", action="store_true", help="time an inference rollout of the workload instead of a training step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cgnn" else args.warmup
    if args.rollout:
        run_rollout(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
