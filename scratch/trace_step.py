"""Timeline of one replayed training step (CUPTI through torch.profiler): per-kernel start, duration and
the idle gap before it.  usage: python scratch/trace_step.py [out.txt] [n_gpus_ignored]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bliss_gnn_b200 import _native as N  # noqa: E402
from bliss_gnn_b200.train import DataModule, Trainer, build_model  # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "trace_step.txt")
N.build()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}")
torch.cuda.set_device(dev)
pg = None
if world > 1:       # data-parallel timeline: python -m torch.distributed.run --nproc-per-node N scratch/trace_step.py out.txt
    torch.distributed.init_process_group("nccl", device_id=dev)
    pg = torch.distributed.group.WORLD
torch.set_float32_matmul_precision("medium")
g = bench.build_graph("reddit", 1.0, dev)
dm = DataModule("reddit", fan_out=bench.FANOUT, eta=bench.ETA, device=dev, batch_size=bench.BATCH,
                sampler="poisson-bandit", model="sage", seed=0, rank=rank, world_size=world, graph=g)
torch.manual_seed(3)
model = build_model("sage", dm.in_feats, bench.HIDDEN, dm.n_classes, 3, bench.DROPOUT).to(dev)
tr = Trainer(dm, model, bench.LR, pg, static_graph=True)
batches = [b.to(dev) for b in bench.seed_batches_for(g, rank, world, 64)]
for i in range(20):
    tr.training_step(batches[i], batches[i + 1])
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(20, 27):
        tr.training_step(batches[i], batches[i + 1])
    torch.cuda.synchronize()
ev = [e for e in prof.profiler.kineto_results.events() if "CUDA" in str(e.device_type()) or "cuda" in str(e.device_type()).lower()]
ev = [e for e in ev if e.duration_ns() > 0]
ev.sort(key=lambda e: e.start_ns())
streams = {}
# one step = from one feature gather (first kernel of the forward pass) to the next
marks = [i for i, e in enumerate(ev) if "k_gather_rows<4>" in e.name()]
with open(out_path if rank == 0 else os.devnull, "w") as f:
    if len(marks) >= 4:
        a, b = marks[-3], marks[-2]
        seg = ev[a:b]
        t0 = seg[0].start_ns()
        prev_end = t0
        busy = 0.0
        for e in seg:
            sid = streams.setdefault(e.device_resource_id(), len(streams))
            s_, d = (e.start_ns() - t0) / 1e3, e.duration_ns() / 1e3
            gap = (e.start_ns() - prev_end) / 1e3
            busy += d
            prev_end = max(prev_end, e.start_ns() + e.duration_ns())
            f.write(f"{s_:9.1f} {d:8.1f} gap {gap:6.1f}  s{sid} {e.name()[:96]}\n")
        f.write(f"# step span {(ev[b].start_ns() - t0) / 1e3:.1f} us, busy {busy:.1f} us, kernels {len(seg)}, streams {len(streams)}\n")
    else:
        f.write(f"# could not split steps: {len(ev)} device events, {len(marks)} marks\n")
        for e in ev[:400]:
            f.write(f"{e.start_ns() / 1e3:.1f} {e.duration_ns() / 1e3:.1f} {e.name()[:100]}\n")
if rank == 0:
    print(open(out_path).read()[-3000:])
if world > 1:
    torch.distributed.destroy_process_group()
