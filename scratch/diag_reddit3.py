"""Micro-checks: fp32 GEMM accuracy at the Reddit layer shapes, and SpMM / transposed SpMM on Reddit-size blocks."""
import sys, os, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bliss_gnn_b200 import _native, ops
from bliss_gnn_b200.graph import synthetic_graph, normalized_edata
from bliss_gnn_b200.sampler import PoissonBanditLadiesSampler
from tests.util import rel_to_max
_native.build()
torch.set_float32_matmul_precision("highest")
dev = torch.device("cuda:0")
print("allow_tf32", torch.backends.cuda.matmul.allow_tf32, torch.get_float32_matmul_precision())
for (m, k, n) in [(1299, 41, 256), (256, 41, 256), (3242, 256, 256), (7225, 604, 256), (1299, 256, 41)]:
    a, b = torch.randn(m, k, device=dev), torch.randn(k, n, device=dev)
    print("mm", (m, k, n), rel_to_max(a @ b, a.double() @ b.double()),
          "linear", rel_to_max(F.linear(a, b.t().contiguous()), a.double() @ b.double()))
gd = synthetic_graph("reddit", seed=0, device=dev, with_features=False)
gd.edata["w"] = normalized_edata(gd)
train = torch.nonzero(gd.ndata["train_mask"], as_tuple=True)[0]
seeds = train[torch.randperm(train.numel(), generator=torch.Generator().manual_seed(1))[:256].to(dev)]
smp = PoissonBanditLadiesSampler([4096, 2048, 1024], eta=0.1, rng_seed=2)
_, _, blocks = smp.sample_blocks(gd, seeds)
for l, blk in enumerate(blocks):
    for dim in (256, 41):
        x = torch.randn(blk.num_src_nodes(), dim, device=dev, requires_grad=True)
        w = blk.edata["edge_weights"]
        ds = ops.mean_scale(blk)
        y = ops.spmm(blk, x, w, dst_scale=ds)
        gy = torch.randn_like(y)
        y.backward(gy)
        src, dst = blk.edge_src.long(), blk.edge_dst.long()
        xr = x.detach().double().requires_grad_(True)
        yr = torch.zeros(blk.num_dst_nodes(), dim, device=dev, dtype=torch.float64).index_add(0, dst, xr[src] * w.double().unsqueeze(1)) * ds.double().unsqueeze(1)
        yr.backward(gy.double())
        od = torch.bincount(src, minlength=blk.num_src_nodes()).max().item()
        print(f"block {l} D={dim}: fwd {rel_to_max(y, yr):.2e} bwd {rel_to_max(x.grad, xr.grad):.2e}  (max in-deg {int(blk.in_degrees().max())}, max out-deg {od})")
