import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from bliss_gnn_b200.train import DataModule, Trainer, build_model
dev = torch.device('cuda:0')
g = bench.build_graph('reddit', 1.0, dev)
dm = DataModule('reddit', fan_out=bench.FANOUT, eta=bench.ETA, device=dev, batch_size=bench.BATCH, sampler='poisson-bandit', model='sage', seed=0, graph=g)
torch.manual_seed(3)
model = build_model('sage', dm.in_feats, bench.HIDDEN, dm.n_classes, 3, bench.DROPOUT).to(dev)
tr = Trainer(dm, model, bench.LR, None, static_graph=True, pipeline=False)
batches = [b.to(dev) for b in bench.seed_batches_for(g, 0, 1, 400)]
E = [[], [], []]; S = [[], [], []]
for i in range(400):
    tr.training_step(batches[i])
    for l, b in enumerate(tr.last_blocks):
        E[l].append(b.num_edges()); S[l].append(b.num_src_nodes())
import statistics as st
for l in range(3):
    print('layer', l, 'edges min/med/max', min(E[l]), st.median(E[l]), max(E[l]), 'first8 max', max(E[l][:8]), '| src min/med/max', min(S[l]), st.median(S[l]), max(S[l]), 'first8 max', max(S[l][:8]))
print('caps', [(p.cap_src, p.cap_edges) for p in tr._pools], 'replays', tr.graph_replays)
