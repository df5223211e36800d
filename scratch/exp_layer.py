"""Profiling driver: a few sampled steps + SAGE fwd/bwd on the Reddit-shaped graph (short, for ncu)."""
import os, sys, torch
sys.path.insert(0, '/root/repo')
import bench
from bliss_gnn_b200.train import DataModule, Trainer, build_model
dev = torch.device('cuda:0')
torch.set_float32_matmul_precision("medium")
g = bench.build_graph('reddit', 1.0, dev)
dm = DataModule('reddit', fan_out=bench.FANOUT, eta=bench.ETA, device=dev, batch_size=bench.BATCH, sampler='poisson-bandit', model='sage', seed=0, graph=g)
dm.sampler.force_stage_path = True
torch.manual_seed(3)
model = build_model('sage', dm.in_feats, bench.HIDDEN, dm.n_classes, 3, bench.DROPOUT).to(dev)
tr = Trainer(dm, model, bench.LR)
batches = [b.to(dev) for b in bench.seed_batches_for(g, 0, 1, 16)]
for i in range(int(os.environ.get('STEPS', '4'))):
    tr.training_step(batches[i])
torch.cuda.synchronize()
print('ok', [ (b.num_src_nodes(), b.num_edges()) for b in tr.last_blocks])
