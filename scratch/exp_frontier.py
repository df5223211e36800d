import os, sys, json, torch, ctypes as C
sys.path.insert(0, '/root/repo')
from bliss_gnn_b200 import _native as N
from bliss_gnn_b200.graph import synthetic_graph, normalized_edata
from bliss_gnn_b200.sampler import PoissonBanditLadiesSampler
dev = torch.device('cuda:0')
g = synthetic_graph('reddit', seed=0, device=dev, with_features=False)
g.edata['w'] = normalized_edata(g)
train = torch.nonzero(g.ndata['train_mask'], as_tuple=True)[0].to(torch.int32)
s = PoissonBanditLadiesSampler([4096, 2048, 1024], eta=0.1, rng_seed=2)
for i in range(3):
    inp, out, blocks = s.sample_blocks(g, train[i*256:(i+1)*256])
seeds0 = blocks[1].srcdata['_ID'].clone()          # layer-0 frontier (3.4K seeds)
print('seeds', seeds0.numel(), 'E_in', int(g.in_degrees(seeds0).sum()))
wsp = s._wsp
def run(flags, reps=int(os.environ.get('REPS','20'))):
    ts = []
    for r in range(reps):
        N.call("bliss_frontier_plan", C.byref(wsp.gview), N.ptr(seeds0), seeds0.numel(), C.byref(wsp.ws), N.stream())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        N.call("bliss_frontier_prob", C.byref(wsp.gview), N.ptr(seeds0), seeds0.numel(), N.ptr(s._w_csc[0]), 0.1, flags, C.byref(wsp.ws), N.stream())
        e1.record()
        # restore invariant
        out_ = N.BlockOut(cap_edges=0, cap_src=0)
        N.call("bliss_block_finish", seeds0.numel(), 0, C.byref(wsp.ws), C.byref(out_), N.stream())
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts)//2]
for name, fl in [("dense collect", 0), ("bitmap collect", 16)]:
    print(f"{name:22s} {run(fl):8.1f} us")
