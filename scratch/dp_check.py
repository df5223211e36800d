"""Data-parallel consistency under torchrun: the peer-memory exchanges (BLISS_P2P=1: bandit updates + gradients through
symmetric-memory windows) against the NCCL collectives (BLISS_P2P=0) — same seeds, same batches: losses of every rank
and the final EXP3 weights / parameters must agree; parameters must be bit-identical ACROSS ranks in both modes.
usage: python -m torch.distributed.run --nproc-per-node N scratch/dp_check.py"""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
from bliss_gnn_b200 import _native as N
from bliss_gnn_b200.graph import synthetic_graph
from bliss_gnn_b200.train import DataModule, Trainer, build_model
N.build()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device(f"cuda:{int(os.environ['LOCAL_RANK'])}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
torch.set_float32_matmul_precision("highest")
g = synthetic_graph("flickr", seed=0, device=dev, scale=0.3)
res = {}
for mode in ("1", "0"):
    os.environ["BLISS_P2P"] = mode
    dm = DataModule("flickr", fan_out=[512, 256, 128], eta=0.1, device=dev, batch_size=64, sampler="poisson-bandit",
                    model="sage", seed=0, rank=rank, world_size=world, graph=g)
    torch.manual_seed(3)
    model = build_model("sage", dm.in_feats, 128, dm.n_classes, 3, dropout=0.0).to(dev)
    tr = Trainer(dm, model, 0.002, dist.group.WORLD, static_graph=True, eager_warmup=3)
    batches = [b for _, b in zip(range(40), dm.train_batches())]
    losses = [float(tr.training_step(b, batches[i + 1] if i + 1 < len(batches) else None).item()) for i, b in enumerate(batches)]
    tr.flush()
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same_across_ranks = bool(torch.equal(ref, flat))
    res[mode] = (losses, flat.clone(), dm.sampler.exp3_weights.clone(), same_across_ranks,
                 bool(tr._exchange is not None and tr._exchange.p2p), bool(tr._gradx),
                 bool(tr._exchange is not None and getattr(tr._exchange, "mc_base", 0)),
                 bool(tr._gradx and getattr(tr._gradx, "mc_base", 0)))
l1, l0 = res["1"][0], res["0"][0]
worst = max(abs(a - b) / max(1.0, abs(b)) for a, b in zip(l1, l0))
pdiff = ((res["1"][1] - res["0"][1]).abs().max() / res["0"][1].abs().max()).item()
wdiff = ((res["1"][2] - res["0"][2]).abs() / res["0"][2]).max().item()
out = {"rank": rank, "world": world, "p2p_on": res["1"][4:], "p2p_off": res["0"][4:], "loss_max_rel": worst,
       "param_max_over_max": pdiff, "exp3_max_rel": wdiff, "params_identical_across_ranks": [res["1"][3], res["0"][3]],
       "last_losses": [l1[-1], l0[-1]]}
print(json.dumps(out), flush=True)
ok = worst <= 1e-5 and wdiff <= 1e-5 and res["1"][3] and res["0"][3] and res["1"][4] and res["1"][5]
dist.destroy_process_group()
sys.exit(0 if ok else 1)
