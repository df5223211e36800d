"""Full-neighbour layer-wise inference at the Reddit shape (model.py:335-383): 114.8 M edges x
256-wide rows per hidden layer.  Times `model.inference` and, per kernel (csrc/profile.cu), the whole-graph SpMM with
its own roofline line: algorithmic bytes = E (4 B index) + 4 (V+1) + 4 D (V + V) per layer against the measured HBM
peak, and the bytes actually gathered (4 D E) against the measured L2->SM gather peak.  Prints one JSON line."""
import json, os, sys, statistics, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from bliss_gnn_b200 import _native as N
from bliss_gnn_b200.train import DataModule, build_model
N.build()
dev = torch.device("cuda:0")
torch.set_float32_matmul_precision("medium")
shape = sys.argv[1] if len(sys.argv) > 1 else "reddit"
g = bench.build_graph(shape, 1.0, dev)
dm = DataModule(shape, fan_out=bench.FANOUT, eta=bench.ETA, device=dev, batch_size=bench.BATCH, sampler="poisson-bandit",
                model="sage", seed=0, graph=g)
torch.manual_seed(3)
model = build_model("sage", dm.in_feats, bench.HIDDEN, dm.n_classes, 3, bench.DROPOUT).to(dev)
for _ in range(2):
    model.inference(dm.g, dev, 128)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pred = model.inference(dm.g, dev, 128); e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1))
N.profile_enable(True)
model.inference(dm.g, dev, 128)
torch.cuda.synchronize()
per_kernel = N.profile_read()
N.profile_enable(False)
V, E = dm.g.num_nodes(), dm.g.num_edges()
dims = [bench.HIDDEN, bench.HIDDEN, dm.n_classes]        # SAGE projects before aggregating when in > out (all three layers here)
alg = sum(4.0 * E + 4.0 * (V + 1) + 4.0 * d * (V + V) for d in dims)
gathered = sum(4.0 * d * E for d in dims)
peak, src = bench._peaks()
calls, ms = 0, 0.0
for k in ("k_spmm_item", "k_spmm_seg", "k_spmm"):
    if k in per_kernel:
        calls, ms = calls + per_kernel[k][0], ms + per_kernel[k][1]
acc = bench.micro_f1(pred[dm.test_nid.long()], dm.g.ndata["labels"][dm.test_nid.long()], dm.multilabel) if hasattr(bench, "micro_f1") else None
out = {"workload": f"{shape}-shape full-neighbour inference, 3-layer SAGE hidden {bench.HIDDEN}", "nodes": V, "edges": E,
       "inference_ms_median": statistics.median(ts), "spmm_kernel_ms": ms, "spmm_launches": calls,
       "roofline": {"bound": "hbm", "kernel": "whole-graph SpMM", "alg_bytes": alg, "achieved": alg / 1e9 / (ms / 1e3),
                    "peak": peak, "unit": "GB/s", "frac": alg / 1e9 / (ms / 1e3) / peak, "peak_source": src},
       "gathered": {"bytes": gathered, "achieved_gbs": gathered / 1e9 / (ms / 1e3)},
       "per_kernel_ms": {k: v[1] for k, v in per_kernel.items()}}
print(json.dumps(out))
