"""BASELINE.json configs[2] and [4] at full size: a few whole-step-graph training steps each (steps/s, loss)."""
import sys, time, torch
sys.path.insert(0, '/root/repo')
from bliss_gnn_b200.graph import synthetic_graph
from bliss_gnn_b200.train import DataModule, Trainer, build_model
dev = torch.device('cuda:0')
torch.set_float32_matmul_precision("medium")
for shape, kind, sampler in [("flickr", "gat", "poisson-bandit"), ("yelp", "sage", "poisson-bandit"), ("yelp", "gat", "poisson-bandit"),
                             ("pubmed", "gcn", "poisson-ladies")]:
    g = synthetic_graph(shape, seed=0, device=dev)
    fan = [4096, 2048, 1024] if shape != "pubmed" else [512, 256, 128]
    bs = 256 if shape != "pubmed" else 32
    dm = DataModule(shape, fan_out=fan, eta=0.1, device=dev, batch_size=bs, sampler=sampler, model=kind, seed=0, graph=g)
    torch.manual_seed(3)
    model = build_model(kind, dm.in_feats, 256, dm.n_classes, 3, 0.1, faithful_gcn_quirk=False).to(dev)
    tr = Trainer(dm, model, 0.002, static_graph=True, eager_warmup=6)
    batches = []
    while len(batches) < 16 + 200 + 1:
        batches.extend(dm.train_batches())
    losses = []
    for i in range(16):
        losses.append(tr.training_step(batches[i], batches[i + 1]))
    torch.cuda.synchronize()
    n = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(16, 16 + n):
        loss = tr.training_step(batches[i], batches[i + 1])      # the next batch is announced (look-ahead sampling)
    tr.flush()
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) / 1e3
    print(f"{shape:7s} {kind:5s} {sampler:15s} V={g.num_nodes()} E={g.num_edges()} replays={tr.graph_replays} resizes={tr.pool_resizes} "
          f"steps/s={n/dt:8.1f} loss {float(losses[0]):.4f} -> {float(loss):.4f} blocks {[ (b.num_src_nodes(), b.num_edges()) for b in tr.last_blocks]}", flush=True)
    del tr, dm, model, g
    torch.cuda.empty_cache()
