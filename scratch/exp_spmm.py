"""SpMM forward / transposed-SpMM on the three blocks of a Reddit-shape step: CUDA-event time per launch (median of 30),
gathered GB/s.  Run once with BLISS_SPMM_TMA=0 (LDG gather) and once with 1 (cp.async.bulk ring)."""
import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bliss_gnn_b200 import _native, ops
from bliss_gnn_b200.graph import synthetic_graph, normalized_edata
from bliss_gnn_b200.sampler import PoissonBanditLadiesSampler
_native.build()
dev = torch.device("cuda:0")
gd = synthetic_graph("reddit", seed=0, device=dev, with_features=False)
gd.edata["w"] = normalized_edata(gd)
train = torch.nonzero(gd.ndata["train_mask"], as_tuple=True)[0]
seeds = train[torch.randperm(train.numel(), generator=torch.Generator().manual_seed(1))[:256].to(dev)]
smp = PoissonBanditLadiesSampler([4096, 2048, 1024], eta=0.1, rng_seed=2)
_, _, blocks = smp.sample_blocks(gd, seeds)
flush = torch.empty(64 * 1024 * 1024, device=dev)
def timeit(fn, n=30):
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)
print("TMA", os.environ.get("BLISS_SPMM_TMA", "0"), "MODE", os.environ.get("BLISS_SPMM_MODE", "item"), "LB", os.environ.get("BLISS_SPMM_LB", "4"))
for l, blk in enumerate(blocks):
    for dim in (256,):
        x = torch.randn(blk.num_src_nodes(), dim, device=dev)
        gy = torch.randn(blk.num_dst_nodes(), dim, device=dev)
        w = blk.edata["edge_weights"]
        ds = ops.mean_scale(blk)
        t_indptr, t_dst, t_perm, t_seg = ops.block_transpose(blk)
        fwd = lambda: ops._spmm_raw(blk.indptr, blk.edge_src, None, w, None, ds, 0, x, blk.num_dst_nodes(), blk.seg_ptr)
        bwd = lambda: ops._spmm_raw(t_indptr, t_dst, t_perm, w, ds, None, 0, gy, blk.num_src_nodes(), t_seg)
        for _ in range(3): fwd(); bwd()
        tf, tb = timeit(fwd), timeit(bwd)
        E = blk.num_edges()
        print(f"block {l} E={E} D={dim}: fwd {tf:7.1f} us ({E*dim*4/tf/1e3:7.0f} GB/s gathered)   bwd {tb:7.1f} us ({E*dim*4/tb/1e3:7.0f} GB/s)")
