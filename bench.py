#!/usr/bin/env python
"""bench.py — Reddit-shape sampled training steps/s of the BLISS hot path on B200 (BASELINE.json metric).

One step = 3-layer poisson-bandit sampling (fan-out 4096/2048/1024, batch 256) + input-feature
gather + SAGE forward/backward + Adam + EXP3 bandit update, on a synthetic Reddit-shaped graph
(232,965 nodes, ~114.6 M edges, 602 features, 41 classes).  Data-parallel runs shard
the seed batches over the ranks (replicated graph) and all-reduce gradients + all-gather the sparse
bandit updates over NCCL.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = whole-job steps/s with the seed batches resident in HBM;
`e2e` = the same through the public sampler/trainer API with each step's seeds coming from pinned
host memory and the loss read back to the host; `roofline` = the dominant kernel's algorithmic bytes
÷ its CUDA-event time against the measured HBM peak; `cpu_baseline` = the CPU oracle (a port of the
reference's algorithm — DGL is not installable, so the reference itself cannot run) timed on this
box's host cores.  `--impl reference` times only that CPU path.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FANOUT = [4096, 2048, 1024]
BATCH = 256
HIDDEN = 256
ETA = 0.1
LR = 0.002
DROPOUT = 0.1
METRIC = "reddit_shape_sampled_steps_per_s"


def _peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.mask, self.max_mhz, self._stop_evt = index, [], 0, None, threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(0.05)     # NVML queries take driver locks: keep them sparse

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": [n for b, n in self.REASONS.items() if self.mask & b], "samples": len(self.samples)}


def _visible_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ---------------------------------------------------------------------------------------------
# CPU path (oracle port of the reference) — used for cpu_baseline and --impl reference
# ---------------------------------------------------------------------------------------------
def cpu_reference_steps(g_cpu, seed_batches, n_warm, n_steps, budget_s, threads):
    """Times the oracle's full training step (sample + gather + SAGE fwd/bwd + Adam + exp3) on the
    host cores.  Returns (steps timed, seconds, sampled edges per step)."""
    import torch.nn.functional as F
    from oracle import model as omodel
    from oracle import samplers as osamp
    from oracle import philox
    torch.set_num_threads(threads)
    state = {"step": 0}

    def ufn(layer, nids, prob=None):
        return torch.from_numpy(philox.uniform_for_nodes(2, state["step"], layer, nids.numpy()))

    if "w" not in g_cpu.edata:
        g_cpu.edata["w"] = osamp.normalized_edata(g_cpu)          # train_lightning.py:362
    smp = osamp.PoissonBanditLadiesSampler(FANOUT, eta=ETA, model="sage", uniform_fn=ufn)
    torch.manual_seed(3)
    in_feats, n_classes = g_cpu.ndata["features"].shape[1], g_cpu.n_classes
    mdl = omodel.SAGE(in_feats, HIDDEN, n_classes, 3, F.relu, DROPOUT)
    opt = torch.optim.Adam(mdl.parameters(), lr=LR)
    feats, labels = g_cpu.ndata["features"], g_cpu.ndata["labels"]

    def one(seeds):
        inp, _, blocks = smp.sample_blocks(g_cpu, seeds)
        x = feats[inp]
        y = labels[seeds.long()]
        loss = F.cross_entropy(mdl(blocks, x), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        smp.exp3(blocks, g_cpu)
        state["step"] += 1
        return sum(b.num_edges() for b in blocks)

    t_first = time.perf_counter()
    for i in range(n_warm):
        one(seed_batches[i % len(seed_batches)])
    t_warm = (time.perf_counter() - t_first) / max(n_warm, 1)
    k = n_steps if t_warm <= 0 else max(1, min(n_steps, int(budget_s / max(t_warm, 1e-3))))
    edges = 0
    t0 = time.perf_counter()
    for i in range(k):
        edges += one(seed_batches[(n_warm + i) % len(seed_batches)])
    dt = time.perf_counter() - t0
    return k, dt, edges / k


def build_graph(shape, scale, device):
    from bliss_gnn_b200.graph import synthetic_graph
    return synthetic_graph(shape, seed=0, device=device, scale=scale)


def seed_batches_for(g, rank, world, n, seed=1):
    """Batches of the shared epoch permutation of the training nodes; rank r takes r, r+R, …"""
    train = torch.nonzero(g.ndata["train_mask"], as_tuple=True)[0].cpu()
    gen = torch.Generator().manual_seed(seed)
    out = []
    perm = train[torch.randperm(train.numel(), generator=gen)]
    per_epoch = perm.numel() // BATCH
    b = rank
    while len(out) < n:
        if b >= per_epoch - (per_epoch % world):
            perm = train[torch.randperm(train.numel(), generator=gen)]
            b = rank
        out.append(perm[b * BATCH:(b + 1) * BATCH].to(torch.int32))
        b += world
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--shape", type=str, default="reddit")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink |V|,|E| (debug only; the metric is scale 1.0)")
    ap.add_argument("--normalize", type=str, default="lazy", choices=["lazy", "literal"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="disable the static-shape CUDA-graph replay of fwd/bwd/Adam")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    threads = os.cpu_count() or 1
    workload = f"{args.shape}-shaped synthetic graph" + ("" if args.scale == 1.0 else f" (scale {args.scale})")
    config = {"workload": f"{workload}, 3-layer SAGE, poisson-bandit, batch {BATCH}/rank, fan-out 4096/2048/1024",
              "sampler": "poisson-bandit", "batch_per_rank": BATCH, "fan_out": FANOUT, "hidden": HIDDEN,
              "eta": ETA, "normalize": args.normalize, "parallelism": f"dp{world}",
              "model_step": "eager" if args.eager else "cuda-graph replay over capacity-padded blocks",
              "l2": "inputs (CSC + 3 EXP3 weight layers + features ≈ 2.4 GB, random rows per step) exceed the 126 MB L2"}

    # ---------------- reference arm: the CPU path only ----------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        g = build_graph(args.shape, args.scale, dev).to("cpu")
        batches = seed_batches_for(g, 0, 1, max(args.steps + args.warmup, 4))
        k, dt, edges = cpu_reference_steps(g, batches, min(args.warmup, 1) or 1, args.steps, 150.0, threads)
        v = k / dt
        sample = f"{k} full training steps (batch {BATCH}) of the oracle port on {threads} threads, bounded to ~150 s"
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "steps/s", "n_gpus": args.gpus, "steps": k,
            "warmup": min(args.warmup, 1) or 1, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "sampled_edges_per_s": edges * v,
            "cpu_baseline": {"value": v, "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ---------------- B200 arm ----------------
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the B200 arm has no CPU fallback"}))
        return 1
    import torch.nn.functional as F  # noqa: F401
    from bliss_gnn_b200 import _native as N
    from bliss_gnn_b200.train import DataModule, Trainer, build_model
    N.build()
    device = torch.device(f"cuda:{local}")
    torch.cuda.set_device(device)
    pg = None
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("BENCH_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        torch.distributed.init_process_group("nccl", device_id=device)
        pg = torch.distributed.group.WORLD
    torch.set_float32_matmul_precision("medium")      # the reference's --precision default (train_lightning.py:550)

    g = build_graph(args.shape, args.scale, device)
    dm = DataModule(args.shape, fan_out=FANOUT, eta=ETA, device=device, batch_size=BATCH, sampler="poisson-bandit",
                    model="sage", seed=0, rank=rank, world_size=world, graph=g, normalize=args.normalize)
    torch.manual_seed(3)
    model = build_model("sage", dm.in_feats, HIDDEN, dm.n_classes, 3, DROPOUT).to(device)
    tr = Trainer(dm, model, LR, pg, static_graph=not args.eager, eager_warmup=8)   # 8 ordinary steps size the pools
    n_need = args.warmup + 3 * args.steps + 24
    host_batches = [b.pin_memory() for b in seed_batches_for(g, rank, world, n_need)]
    dev_batches = [b.to(device) for b in host_batches]
    it = iter(range(n_need))

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3) + (0 if args.eager else tr.eager_warmup + 2)):   # + pool sizing and graph capture
        tr.training_step(dev_batches[next(it)])

    # ---- (1) device-resident throughput: EXACTLY `steps` steps between two events ----
    clocks = ClockSampler(_visible_index(local))
    N.STATS.reset(timing=False)
    replays0 = tr.graph_replays
    tr.flush()
    resizes0 = tr.pool_resizes
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    edges0 = tr.total_sampled_edges
    e0.record()
    for _ in range(args.steps):
        tr.training_step(dev_batches[next(it)])
    tr.flush()               # the last step's counters (sizes, capacity flags) are consumed inside the timed region
    e1.record()
    barrier()
    edges = tr.total_sampled_edges - edges0
    clock_info = clocks.stop()
    # kernels launched eagerly + the hand-written kernels inside every CUDA-graph replay
    launches = N.STATS.launches + (tr.graph_replays - replays0) * tr.graph_kernels
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * args.steps / (ms_total / 1e3)

    # ---- (2) end to end: seeds from pinned host memory every step, loss read back every step ----
    # The loss of every step is copied to pinned host memory stream-ordered and consumed one step later
    # (like the step's counters), so the host never stalls the device inside the loop.
    barrier()
    loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    losses = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = tr.training_step(host_batches[next(it)])
        loss_host[i & 1].copy_(loss.reshape(1), non_blocking=True)
        loss_ev[i & 1].record()
        if i:
            loss_ev[(i - 1) & 1].synchronize()
            losses.append(float(loss_host[(i - 1) & 1]))
    loss_ev[(args.steps - 1) & 1].synchronize()
    losses.append(float(loss_host[(args.steps - 1) & 1]))
    tr.flush()
    e1.record()
    barrier()
    assert len(losses) == args.steps and all(l == l for l in losses), "e2e: a loss did not come back"
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        torch.distributed.all_reduce(ms2, op=torch.distributed.ReduceOp.MAX)
    e2e_value = world * args.steps / (float(ms2.item()) / 1e3)

    # ---- (3) per-entry-point CUDA-event timing over the same kind of steps (roofline) ----
    n_prof = min(args.steps, 50)
    tr.flush()
    dm.sampler.force_stage_path = True     # one FFI call per kernel, so each is bracketed by its own events
    N.STATS.reset(timing=True)
    sizes = []
    for _ in range(n_prof):
        tr.training_step(dev_batches[next(it)])
        sizes.append([(c.n_seeds, c.e_in, c.n_cand, c.n_src, c.n_edges) for c in dm.sampler.last_counters])
    torch.cuda.synchronize()
    per_fn = N.STATS.elapsed_ms()
    N.STATS.reset(timing=False)
    dm.sampler.force_stage_path = False
    peak, peak_src = _peaks()
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
    except Exception:
        traffic_tab = {}

    layer_dims = [HIDDEN, HIDDEN, dm.n_classes]          # width the aggregation runs at in each layer (SAGE, lin_before_mp)

    def alg_bytes(name):
        """Algorithmic (compulsory) bytes of one entry point over the profiled steps: every distinct
        input read once, every output written once."""
        total = 0.0
        for step_sizes in sizes:
            for l, (n_s, e_in, n_c, n_src, e_b) in enumerate(step_sizes):
                if name == "bliss_frontier_prob":    # (index, weight) per in-edge + chunk records/partials + |V| accumulator scan + candidates
                    total += 8.0 * e_in + 64.0 * (e_in / 256.0 + n_s) + 8.0 * g.num_nodes() + 8.0 * n_c
                elif name == "bliss_block_count":    # indices + chunk records + keep bits + per kept edge (info, weight, first key)
                    total += 4.0 * e_in + 76.0 * (e_in / 256.0 + n_s) + 20.0 * e_b
                elif name in ("bliss_block_fill", "bliss_sample_layer_back"):
                    total += 80.0 * (e_in / 256.0 + n_s) + 16.0 * e_b + 36.0 * e_b + 36.0 * n_c
                elif name == "bliss_spmm":           # forward + backward of the layer at its aggregation width
                    total += 2 * (8.0 * e_b + 8.0 * (n_s + 1) + 4.0 * layer_dims[l] * (n_src + n_s))
                else:
                    total += 8.0 * e_in
        return total

    def roofline_of(name, note):
        calls, t_ms = per_fn[name]
        achieved = alg_bytes(name) / 1e9 / (t_ms / 1e3) if t_ms > 0 else 0.0
        tr_ = traffic_tab.get(name, {}).get("dram_bytes_per_launch")
        return {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": tr_, "alg_bytes_per_launch": alg_bytes(name) / max(calls, 1),
                "peak_source": peak_src, "launches_timed": calls, "avg_launch_us": 1e3 * t_ms / max(calls, 1),
                "note": note}

    name = max(per_fn.items(), key=lambda kv: kv[1][1])[0]
    roofline = roofline_of(name, "entry point with the largest total CUDA-event time over the profiled steps. The SpMM "
                                 "gathers source rows that are L2-resident (<= 8 K rows x 1 KB), so its DRAM traffic is "
                                 "about its compulsory bytes and the HBM fraction is low by construction: the kernel is "
                                 "bound by L2->SM gather bandwidth, reported as l2_gather (DESIGN.md section 3)")
    roofline["per_entry_point_ms_per_step"] = {k: v[1] / n_prof for k, v in sorted(per_fn.items())}
    if "bliss_spmm" in per_fn:      # the bound that actually applies: bytes gathered through L2 per second
        gathered = sum(2 * 4.0 * layer_dims[l] * e_b for st in sizes for l, (_, _, _, _, e_b) in enumerate(st))
        t_ms = per_fn["bliss_spmm"][1]
        roofline["l2_gather"] = {"achieved_gbs": gathered / 1e9 / (t_ms / 1e3),
                                 "peak_gbs": 12000.0, "peak_source": "chip-level L2->SM throughput ~6300 B/clk "
                                 "(/opt/skills/guides/B300_MICROARCH.md, TMA/LDG chip-throughput) x 1.9 GHz",
                                 "bytes_per_step": gathered / n_prof}
    roofline_sampling = roofline_of("bliss_frontier_prob", "the HBM-streaming sampling passes the north-star names "
                                    "(three chunk passes + candidate scan); the scatter pass is bound by L2 atomic "
                                    "throughput (one 64-bit RED per in-edge), not by HBM") if "bliss_frontier_prob" in per_fn else None

    out = {"metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
           "sampled_edges_per_s": world * edges / (ms_total / 1e3), "clocks": clock_info,
           "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": BATCH * 4 * world,
                   "d2h_bytes_per_step": (4 + dm.sampler._wsp.ctr_all.numel()) * world,
                   "note": "seeds H2D from pinned memory every step; loss + per-layer counters D2H every step, "
                           "copied stream-ordered and consumed one step later"},
           "gpu_launches": launches, "graph_replays": tr.graph_replays,
           "pool_resizes_in_timed_regions": tr.pool_resizes - resizes0, "roofline": roofline,
           "roofline_sampling": roofline_sampling}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        g_cpu = g.to("cpu")
        k, dt, e_cpu = cpu_reference_steps(g_cpu, [b.clone() for b in host_batches[:8]], 1, 4, args.cpu_budget_s,
                                           threads)
        out["cpu_baseline"] = {"value": k / dt, "unit": "steps/s", "cores": threads, "kind": "port",
                               "sample": f"{k} full training steps of the same workload (oracle port, batch {BATCH})"}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
