import cProfile, pstats, sys, os, io, torch
sys.path.insert(0, '/root/repo')
import bench
from bliss_gnn_b200.train import DataModule, Trainer, build_model
dev = torch.device('cuda:0')
torch.set_float32_matmul_precision("medium")
g = bench.build_graph('reddit', 1.0, dev)
dm = DataModule('reddit', fan_out=bench.FANOUT, eta=bench.ETA, device=dev, batch_size=bench.BATCH, sampler='poisson-bandit', model='sage', seed=0, graph=g)
torch.manual_seed(3)
model = build_model('sage', dm.in_feats, bench.HIDDEN, dm.n_classes, 3, bench.DROPOUT).to(dev)
tr = Trainer(dm, model, bench.LR)
batches = [b.to(dev) for b in bench.seed_batches_for(g, 0, 1, 140)]
for i in range(20): tr.training_step(batches[i])
torch.cuda.synchronize()
import time
# phase timing (host wall, with sync at phase ends to attribute)
def phases(n=50):
    t = dict(sample=0., fwd=0., bwd=0., opt=0., exp3=0.)
    for i in range(n):
        seeds = batches[20+i]
        torch.cuda.synchronize(); t0=time.perf_counter()
        inp, out, mfgs = dm.sampler.sample_blocks(dm.g, seeds)
        torch.cuda.synchronize(); t1=time.perf_counter()
        x = mfgs[0].srcdata['features']; y = mfgs[-1].dstdata['labels']
        pred = tr.model(mfgs, x); loss = tr.loss_fn(pred, y)
        torch.cuda.synchronize(); t2=time.perf_counter()
        tr._flat_grad.zero_(); loss.backward()
        torch.cuda.synchronize(); t3=time.perf_counter()
        tr.optimizer.step()
        torch.cuda.synchronize(); t4=time.perf_counter()
        dm.sampler.exp3(mfgs, dm.g)
        torch.cuda.synchronize(); t5=time.perf_counter()
        for k,v in zip(t, (t1-t0,t2-t1,t3-t2,t4-t3,t5-t4)): t[k]+=v
    print({k: round(1e3*v/n,3) for k,v in t.items()}, 'ms per step (synced phases)')
phases()
pr = cProfile.Profile(); pr.enable()
for i in range(50): tr.training_step(batches[70+i])
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(28); print(s.getvalue()[:6000])
