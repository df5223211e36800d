import os, sys, torch, ctypes as C
sys.path.insert(0, '/root/repo')
from bliss_gnn_b200 import _native as N
from bliss_gnn_b200.graph import synthetic_graph, normalized_edata
from bliss_gnn_b200.sampler import PoissonBanditLadiesSampler
dev = torch.device('cuda:0')
g = synthetic_graph('reddit', seed=0, device=dev, with_features=False)
g.edata['w'] = normalized_edata(g)
s = PoissonBanditLadiesSampler([4096, 2048, 1024], eta=0.1, rng_seed=2)
s._bind(g)
wsp = s._wsp
deg = g.in_degrees()
order = torch.argsort(deg, descending=True)
def run(seeds, reps=10):
    seeds = seeds.to(torch.int32).contiguous()
    ts = []
    for r in range(reps):
        N.call("bliss_frontier_plan", C.byref(wsp.gview), N.ptr(seeds), seeds.numel(), C.byref(wsp.ws), N.stream())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        N.call("bliss_frontier_prob", C.byref(wsp.gview), N.ptr(seeds), seeds.numel(), N.ptr(s._w_csc[0]), 0.1, 0, C.byref(wsp.ws), N.stream())
        e1.record()
        out_ = N.BlockOut(cap_edges=0, cap_src=0)
        N.call("bliss_block_finish", seeds.numel(), 0, C.byref(wsp.ws), C.byref(out_), N.stream())
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts)//2], int(deg[seeds.long()].sum())
def pick(lo, hi, n):
    m = torch.nonzero((deg >= lo) & (deg < hi), as_tuple=True)[0]
    return m[torch.randperm(m.numel(), device=dev)[:n]]
for name, sd in [("1 row max deg", order[:1]), ("8 rows max deg", order[:8]), ("1 row ~6000", pick(5900, 6100, 1)),
                 ("1 row ~1400", pick(1350, 1450, 1)), ("888 rows ~1400", pick(1300, 1500, 888)),
                 ("3300 rows ~1400", pick(1200, 1600, 3300)), ("888 rows ~3000", pick(2800, 3200, 888)),
                 ("3300 rows 300-500", pick(300, 500, 3300)), ("3300 rows <=256 (light)", pick(100, 256, 3300))]:
    t, e = run(sd)
    print(f"{name:26s} rows={sd.numel():5d} edges={e:9d} {t:8.1f} us  {8*e/t/1e3:8.1f} GB/s")
