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
    """Times the oracle's full training step (sample + gather + SAGE fwd/bwd + Adam + exp3) on the host cores,
    step by step (``time.perf_counter`` around each step; BASELINE.md §2: 3 warm-up steps, >= 20 timed steps,
    median and p10/p90).  ``budget_s`` bounds the timed part.  Returns a dict."""
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

    for i in range(n_warm):
        one(seed_batches[i % len(seed_batches)])
    times, edges = [], 0
    t_begin = time.perf_counter()
    for i in range(n_steps):
        t0 = time.perf_counter()
        edges += one(seed_batches[(n_warm + i) % len(seed_batches)])
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s:
            break
    times.sort()
    k = len(times)
    pick = lambda q: times[min(k - 1, int(q * k))]
    return {"steps": k, "warmup": n_warm, "seconds": sum(times), "median_s": statistics.median(times),
            "p10_s": pick(0.10), "p90_s": pick(0.90), "edges_per_step": edges / k}


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
              "model_step": "eager" if args.eager else "cuda-graph replay over capacity-padded blocks; the next batch's "
                            "blocks are sampled beside the backward pass (after this step's exp3)",
              "l2": "inputs (CSC + 3 EXP3 weight layers + features ≈ 2.4 GB, random rows per step) exceed the 126 MB L2"}

    # ---------------- reference arm: the CPU path only ----------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        g = build_graph(args.shape, args.scale, dev).to("cpu")
        n_warm = max(1, min(args.warmup, 3))
        batches = seed_batches_for(g, 0, 1, args.steps + n_warm)
        r = cpu_reference_steps(g, batches, n_warm, args.steps, 150.0, threads)
        v = 1.0 / r["median_s"]
        sample = (f"{r['steps']} full training steps (batch {BATCH}) of the oracle port on {threads} threads after "
                  f"{n_warm} warm-up steps, timed step by step, bounded to ~150 s; value = 1 / median step time")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "steps/s", "n_gpus": args.gpus, "steps": r["steps"],
            "warmup": n_warm, "ms_per_step": 1e3 * r["median_s"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "sampled_edges_per_s": r["edges_per_step"] * v,
            "cpu_baseline": {"value": v, "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample,
                             "ms_per_step_median": 1e3 * r["median_s"], "ms_per_step_p10": 1e3 * r["p10_s"],
                             "ms_per_step_p90": 1e3 * r["p90_s"]},
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
    json_fd = 1
    if world > 1:
        # NCCL's init lines (version, rank / nranks / transport of every communicator) are written to stdout by the
        # library: file descriptor 1 is pointed at stderr for the run and the ONE JSON line goes to the saved stdout,
        # so the driver can read the rank count off the run's own log and still parse stdout
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "INFO"
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,ENV")
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        torch.distributed.init_process_group("nccl", device_id=device)
        pg = torch.distributed.group.WORLD
    torch.set_float32_matmul_precision("medium")      # the reference's --precision default (train_lightning.py:550)

    g = build_graph(args.shape, args.scale, device)
    dm = DataModule(args.shape, fan_out=FANOUT, eta=ETA, device=device, batch_size=BATCH, sampler="poisson-bandit",
                    model="sage", seed=0, rank=rank, world_size=world, graph=g, normalize=args.normalize)
    torch.manual_seed(3)
    model = build_model("sage", dm.in_feats, HIDDEN, dm.n_classes, 3, DROPOUT).to(device)
    tr = Trainer(dm, model, LR, pg, static_graph=not args.eager, eager_warmup=8)   # 8 ordinary steps size the pools
    REPEATS = 5                    # timed regions of exactly `steps` steps each; the median region is reported
    n_prof = min(args.steps, 50)
    n_need = max(args.warmup, 3) + tr.eager_warmup + 2 + 2 * REPEATS * args.steps + n_prof + 8
    host_batches = [b.pin_memory() for b in seed_batches_for(g, rank, world, n_need)]
    dev_batches = [b.to(device) for b in host_batches]
    it = iter(range(n_need))

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=device)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def run_steps(batches, n, each=None):
        """n training steps over consecutive batches, announcing each next batch to the trainer (the data loader's
        look-ahead: its blocks are sampled in the shadow of the current step's backward pass)."""
        i = next(it)
        for k in range(n):
            j = next(it) if k + 1 < n else None
            loss = tr.training_step(batches[i], batches[j] if j is not None else None)
            if each is not None:
                each(k, loss)
            i = j

    run_steps(dev_batches, max(args.warmup, 3) + (0 if args.eager else tr.eager_warmup + 2))   # + pool sizing, capture

    # ---- (1) device-resident throughput: EXACTLY `steps` steps between two events, REPEATS times ----
    clocks = ClockSampler(_visible_index(local))
    N.STATS.reset(timing=False)
    launches0 = tr.graph_kernel_launches
    tr.flush()
    resizes0 = tr.pool_resizes
    barrier()
    clocks.start()
    region_ms, edges = [], 0
    for _ in range(REPEATS):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        edges0 = tr.total_sampled_edges
        e0.record()
        run_steps(dev_batches, args.steps)
        tr.flush()           # the last step's counters (sizes, capacity flags) are consumed inside the timed region
        e1.record()
        barrier()
        region_ms.append(max_over_ranks(e0.elapsed_time(e1)))
        edges += tr.total_sampled_edges - edges0
    clock_info = clocks.stop()
    # kernels launched eagerly + the hand-written kernels inside every CUDA-graph replay
    launches = (N.STATS.launches + tr.graph_kernel_launches - launches0) // REPEATS
    ms_total = statistics.median(region_ms)
    value = world * args.steps / (ms_total / 1e3)

    # ---- (2) end to end: seeds from pinned host memory every step, loss read back every step ----
    # The step graph copies its loss and its blocks' counters to pinned host memory itself (Trainer.host_loss,
    # Trainer._ctr_pin); the host consumes them one step later, so it never stalls the device inside the loop.
    # The seeds are host tensors: staged through pinned memory on a copy stream (Trainer._stage_seeds).
    e2e_ms = []
    for _ in range(REPEATS):
        barrier()
        losses = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()

        def read_back(i, loss):
            if i:                                  # the previous step's loss: already in pinned memory, no stall
                losses.append(tr.host_loss(back=1))

        run_steps(host_batches, args.steps, read_back)
        losses.append(tr.host_loss())
        tr.flush()
        e1.record()
        barrier()
        assert len(losses) == args.steps and all(l == l for l in losses), "e2e: a loss did not come back"
        e2e_ms.append(max_over_ranks(e0.elapsed_time(e1)))
    e2e_value = world * args.steps / (statistics.median(e2e_ms) / 1e3)

    # ---- (3) per-KERNEL CUDA-event timing (csrc/profile.cu) over the same kind of steps: roofline ----
    # The steps run eagerly (every kernel its own launch, bracketed by an event pair on its launch stream), so the
    # times are the kernels' own durations; inside the replayed graph they overlap each other.
    tr.flush()
    dm.sampler.force_stage_path = True
    N.STATS.reset(timing=False)
    N.profile_enable(True)
    sizes = []
    for _ in range(n_prof):
        tr.training_step(dev_batches[next(it)])
        sizes.append([(c.n_seeds, c.e_in, c.n_cand, c.n_src, c.n_edges) for c in dm.sampler.last_counters])
    torch.cuda.synchronize()
    per_kernel = N.profile_read()
    N.profile_enable(False)
    dm.sampler.force_stage_path = False
    peak, peak_src = _peaks()
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    except Exception:
        traffic_tab = {}
    rl = roofline_tables(per_kernel, sizes, n_prof, dm.in_feats, [HIDDEN, HIDDEN, dm.n_classes], peak, peak_src,
                         traffic_tab)
    rl["roofline"]["l2_gather"] = l2_gather_probe(N, device, per_kernel, sizes, [HIDDEN, HIDDEN, dm.n_classes])

    if world > 1:
        ex, gx = tr._exchange, tr._gradx
        how = lambda o: ("peer-memory push" + (" through the NVSwitch multicast address" if getattr(o, "mc_base", 0) else "")) \
            if o is not None and o is not False and getattr(o, "p2p", True) else "NCCL"
        config["exchange"] = {"bandit_updates": how(ex), "gradients": how(gx)}
    out = {"metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
           "repeats": REPEATS, "ms_per_step_all_regions": [m / args.steps for m in region_ms],
           "sampled_edges_per_s": world * edges / (sum(region_ms) / 1e3), "clocks": clock_info,
           "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": BATCH * 4 * world,
                   "d2h_bytes_per_step": (4 + 8 * dm.sampler._wsp.ctr_all.shape[1]) * world,
                   "ms_per_step_all_regions": [m / args.steps for m in e2e_ms],
                   "note": "seeds H2D from pinned memory every step (copy stream); loss + per-layer counters D2H every "
                           "step (copy nodes of the step graph), consumed by the host one step later"},
           "gpu_launches": launches, "graph_replays": tr.graph_replays,
           "pool_resizes_in_timed_regions": tr.pool_resizes - resizes0}
    out.update(rl)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        g_cpu = g.to("cpu")
        r = cpu_reference_steps(g_cpu, [b.clone() for b in host_batches[:23]], 3, 20, args.cpu_budget_s, threads)
        out["cpu_baseline"] = {"value": 1.0 / r["median_s"], "unit": "steps/s", "cores": threads, "kind": "port",
                               "ms_per_step_median": 1e3 * r["median_s"], "ms_per_step_p10": 1e3 * r["p10_s"],
                               "ms_per_step_p90": 1e3 * r["p90_s"],
                               "sample": f"{r['steps']} full training steps of the same workload after 3 warm-up steps "
                                         f"(oracle port, batch {BATCH}, {threads} threads), timed step by step; "
                                         "value = 1 / median step time"}
    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------
# roofline: SURVEY.md section 8(d)'s algorithmic (compulsory) bytes, verbatim
# ---------------------------------------------------------------------------------------------
#: kernel -> stage of the path (SURVEY.md section 8d names the stages)
STAGE_OF = {"k_plan_rows": "sampling", "k_plan_scan": "sampling", "k_plan_chunks": "sampling", "k_prob_pass1": "sampling",
            "k_prob_pass2": "sampling", "k_prob_pass3": "sampling", "k_collect_candidates": "sampling",
            "k_scale_search": "sampling", "k_poisson_scale": "sampling", "k_select_poisson": "sampling",
            "k_layer_front": "sampling",
            "k_block_count<true>": "block_build", "k_block_count<false>": "block_build", "k_block_index": "block_build",
            "k_block_fill": "block_build", "k_block_finish": "block_build",
            "k_spmm_seg": "spmm", "k_spmm_item": "spmm", "k_spmm_combine": "spmm", "k_spmm": "spmm", "k_spmm_tma": "spmm",
            "k_gather_rows": "gather", "k_reward_update": "bandit"}


def stage_bytes(stage, sizes, in_feats, layer_dims):
    """Algorithmic bytes of one stage over the profiled steps — SURVEY.md section 8(d), each distinct input read
    once and each output written once, s = 4 (fp32):
      sampling, per layer   B_samp  = 8 (n_s+1) + E_in (4 idx + 4 weight) + 4 N_c + n_src (4 nid + 4 P)
      block build           B_blk   = E_b (4 src + 4 eid + 4 W~ + 4 q) + 4 (n_s+1)
      SpMM, per layer       B_spmm  = E_b (4 idx + 4 W~) + 4 (n_s+1) + 4 D (n_src + n_s), forward and again backward
      gather                B_g     = 2 * 4 F n_src0 + 4 n_src0
      bandit                B_bandit= sum_l E_b (4 eid + 4 alpha + 4 q + 8 weight RMW) + 8 n_src + 4 n_s   (lazy norm)"""
    total = 0.0
    for step_sizes in sizes:
        for l, (n_s, e_in, n_c, n_src, e_b) in enumerate(step_sizes):
            if stage == "sampling":
                total += 8.0 * (n_s + 1) + 8.0 * e_in + 4.0 * n_c + 8.0 * n_src
            elif stage == "block_build":
                total += 16.0 * e_b + 4.0 * (n_s + 1)
            elif stage == "spmm":
                total += 2 * (8.0 * e_b + 4.0 * (n_s + 1) + 4.0 * layer_dims[l] * (n_src + n_s))
            elif stage == "gather" and l == 0:
                total += 2 * 4.0 * in_feats * n_src + 4.0 * n_src
            elif stage == "bandit":
                total += 20.0 * e_b + 8.0 * n_src + 4.0 * n_s
    return total


def roofline_tables(per_kernel, sizes, n_prof, in_feats, layer_dims, peak, peak_src, traffic_tab):
    """`roofline` = the dominant kernel (largest total CUDA-event time); `roofline_kernels` = every timed kernel;
    `roofline_stages` = the stages of SURVEY section 8(d) with the stage's algorithmic bytes over the summed time of
    its kernels.  A kernel that does its stage's whole work (SpMM, gather, reward update, a fused sampling kernel)
    carries the stage's bytes; kernels that share a stage (the three probability passes re-read the same weights)
    only have a stage-level fraction — dividing one compulsory read among them would be arbitrary."""
    whole_stage = {"k_spmm_seg": "spmm", "k_spmm_item": "spmm", "k_spmm_tma": "spmm", "k_spmm": "spmm", "k_gather_rows": "gather",
                   "k_reward_update": "bandit", "k_layer_front": "sampling"}
    stages, kernels = {}, []
    for name, (calls, ms) in sorted(per_kernel.items(), key=lambda kv: -kv[1][1]):
        st = STAGE_OF.get(name, "other")
        stages.setdefault(st, {"ms": 0.0, "kernels": []})
        stages[st]["ms"] += ms
        stages[st]["kernels"].append(name)
        row = {"kernel": name, "stage": st, "launches_timed": calls, "avg_launch_us": 1e3 * ms / max(calls, 1),
               "ms_per_step": ms / n_prof}
        if name in whole_stage and ms > 0:
            bts = stage_bytes(whole_stage[name], sizes, in_feats, layer_dims)
            row.update({"alg_bytes_per_launch": bts / max(calls, 1), "achieved": bts / 1e9 / (ms / 1e3),
                        "frac": bts / 1e9 / (ms / 1e3) / peak})
        t = traffic_tab.get("kernels", {}).get(name, {}).get("dram_bytes_per_launch")
        if t is not None:
            row["traffic"] = t
        kernels.append(row)
    stage_rows = {}
    for st, d in stages.items():
        if st == "other" or d["ms"] <= 0:
            continue
        bts = stage_bytes(st, sizes, in_feats, layer_dims)
        stage_rows[st] = {"alg_bytes_per_step": bts / n_prof, "ms_per_step": d["ms"] / n_prof,
                          "achieved": bts / 1e9 / (d["ms"] / 1e3), "frac": bts / 1e9 / (d["ms"] / 1e3) / peak,
                          "kernels": d["kernels"]}
    top = next((k for k in kernels if "achieved" in k), kernels[0] if kernels else {})
    dominant = kernels[0] if kernels else {}
    if "achieved" not in dominant and dominant:       # the dominant kernel shares its stage: report the stage's fraction
        st = stage_rows.get(dominant["stage"], {})
        dominant = dict(dominant, achieved=st.get("achieved", 0.0), frac=st.get("frac", 0.0),
                        alg_bytes_per_launch=None, note_bytes="stage-level bytes / stage-level time (kernels share inputs)")
    roofline = {"bound": "hbm", "kernel": dominant.get("kernel"), "achieved": dominant.get("achieved", 0.0), "peak": peak,
                "unit": "GB/s", "frac": dominant.get("frac", 0.0), "traffic": dominant.get("traffic"),
                "alg_bytes_per_launch": dominant.get("alg_bytes_per_launch"), "peak_source": peak_src,
                "launches_timed": dominant.get("launches_timed"), "avg_launch_us": dominant.get("avg_launch_us"),
                "note": "dominant kernel by total CUDA-event time over the profiled (eager) steps; algorithmic bytes per "
                        "SURVEY.md section 8(d).  The sampled blocks' source rows (<= 8 K x 1 KB) are L2-resident, so the "
                        "SpMM's DRAM traffic is about its compulsory bytes and its HBM fraction is low by construction; "
                        "l2_gather gives the bound that applies, measured in this run."}
    return {"roofline": roofline, "roofline_kernels": kernels, "roofline_stages": stage_rows,
            "roofline_top_whole_stage_kernel": top.get("kernel")}


def l2_gather_probe(N, device, per_kernel, sizes, layer_dims):
    """MEASURES the chip's L2 -> SM gather bandwidth (warp-per-row gathers of 1 KB rows from an 8 MB, L2-resident
    table, every SM busy) and puts the SpMM's gathered bytes per second beside it."""
    rows, dim, per_warp = 8192, 256, 512
    table = torch.randn(rows, dim, device=device)
    best = None
    for n_warps in (148 * 8 * 4, 148 * 8 * 8):
        outp = torch.empty(n_warps, dim, device=device)
        for _ in range(2):
            N.call("bliss_l2_gather_probe", N.ptr(table), rows, dim, per_warp, N.ptr(outp), n_warps, N.stream())
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            N.call("bliss_l2_gather_probe", N.ptr(table), rows, dim, per_warp, N.ptr(outp), n_warps, N.stream())
            e1.record()
            e1.synchronize()
            gbs = n_warps * per_warp * dim * 4 / 1e9 / (e0.elapsed_time(e1) / 1e3)
            best = gbs if best is None else max(best, gbs)
    res = {"peak_gbs": best, "peak_source": "measured in this run: bliss_l2_gather_probe (csrc/aggregate.cu), best of 10"}
    t = sum(per_kernel.get(k, (0, 0.0))[1] for k in ("k_spmm_seg", "k_spmm_item", "k_spmm_tma"))
    if t > 0:
        gathered = sum(2 * 4.0 * layer_dims[l] * e_b for st in sizes for l, (_, _, _, _, e_b) in enumerate(st))
        res.update({"achieved_gbs": gathered / 1e9 / (t / 1e3), "bytes_per_step": gathered / len(sizes),
                    "frac": gathered / 1e9 / (t / 1e3) / best})
    return res


if __name__ == "__main__":
    sys.exit(main())
