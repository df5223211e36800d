"""Layer-by-layer gradient comparison (device fp32 vs oracle fp64) at the Reddit shape."""
import sys, os, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bliss_gnn_b200 import _native, ops, model as M
from bliss_gnn_b200.graph import synthetic_graph, normalized_edata
from bliss_gnn_b200.sampler import PoissonBanditLadiesSampler
from oracle import model as omodel, samplers as osamp
from tests.util import copy_params, rel_to_max, philox_uniform_fn, blocks_as
_native.build()
torch.set_float32_matmul_precision("highest")
dev = torch.device("cuda:0")
shape = sys.argv[1] if len(sys.argv) > 1 else "reddit"
gd = synthetic_graph(shape, seed=0, device=dev)
gd.edata["w"] = normalized_edata(gd)
g = gd.to("cpu")
fan = [4096, 2048, 1024]
train = torch.nonzero(g.ndata["train_mask"], as_tuple=True)[0]
seeds = train[torch.randperm(train.numel(), generator=torch.Generator().manual_seed(1))[:256]]
ora = osamp.PoissonBanditLadiesSampler(fan, eta=0.1, accum="contract", uniform_fn=philox_uniform_fn(2, 0))
o_in, _, ob = ora.sample_blocks(g, seeds)
blocks_as(ob, torch.float64)
dsm = PoissonBanditLadiesSampler(fan, eta=0.1, rng_seed=2)
d_in, _, db = dsm.sample_blocks(gd, seeds)
labels = g.ndata["labels"][seeds.long()]
if labels.dim() > 1: labels = labels.argmax(1)
F_in, C = g.ndata["features"].shape[1], max(g.n_classes, 2)
torch.manual_seed(3)
dm = M.SAGE(F_in, 256, C, 3, F.relu, 0.0).to(dev)
om = omodel.SAGE(F_in, 256, C, 3, F.relu, 0.0)
copy_params(om, dm, torch.float64); om = om.double()
# device, layer by layer
x = gd.ndata["features"][d_in.long()]
hs_d, parts_d = [], []
h = x
for l, (layer, blk) in enumerate(zip(dm.layers, db)):
    if l < 2:
        a, b = layer.forward_parts(blk, h, edge_weight=blk.edata["edge_weights"])
        a.retain_grad(); b.retain_grad(); parts_d.append((a, b))
        h, _ = ops.sage_epilogue(a, b, layer.fc_self.bias, True, 0.0, 0, None, l, True)
        h.retain_grad(); hs_d.append(h)
    else:
        out_d = layer(blk, h, edge_weight=blk.edata["edge_weights"])
ld = F.cross_entropy(out_d, labels.to(dev)); ld.backward()
# oracle
xo = g.ndata["features"][o_in].double()
hs_o, parts_o = [], []
h = xo
for l, (layer, blk) in enumerate(zip(om.layers, ob)):
    if l < 2:
        feat_dst = h[: blk.num_dst_nodes()]
        hp = layer.fc_neigh(h)
        deg = blk.in_degrees().clamp(min=1).to(h.dtype)
        b = omodel._spmm(blk, hp, blk.edata["edge_weights"]) / deg.unsqueeze(-1)
        a = F.linear(feat_dst, layer.fc_self.weight)
        a.retain_grad(); b.retain_grad(); parts_o.append((a, b))
        h = F.relu(a + b + layer.fc_self.bias)
        h.retain_grad(); hs_o.append(h)
    else:
        out_o = layer(blk, h, edge_weight=blk.edata["edge_weights"])
lo = F.cross_entropy(out_o, labels); lo.backward()
print("loss", float(ld), float(lo))
print("logits", rel_to_max(out_d, out_o))
for l in (1, 0):
    print(f"layer {l}: h fwd {rel_to_max(hs_d[l], hs_o[l]):.2e}  dL/dh {rel_to_max(hs_d[l].grad, hs_o[l].grad):.2e}  "
          f"dL/da {rel_to_max(parts_d[l][0].grad, parts_o[l][0].grad):.2e} dL/db {rel_to_max(parts_d[l][1].grad, parts_o[l][1].grad):.2e}")
    gd_, go_ = hs_d[l].grad.cpu().double(), hs_o[l].grad
    diff = (gd_ - go_).abs()
    rows = diff.max(1).values
    top = torch.topk(rows, 5)
    print("   worst rows of dL/dh:", top.indices.tolist(), [f"{v:.2e}" for v in top.values.tolist()], "max|ref|", float(go_.abs().max()))
    gate_d, gate_o = (hs_d[l] > 0).cpu(), hs_o[l] > 0
    print("   relu gate mismatches:", int((gate_d != gate_o).sum()), "of", gate_o.numel())
for (n, p), (_, q) in zip(dm.named_parameters(), om.named_parameters()):
    print(f"   {n:28s} {rel_to_max(p.grad, q.grad):.2e}")
