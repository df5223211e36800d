"""Which piece of the Trainer path loses gradient accuracy at the Reddit shape?  One sampled step, four ways."""
import sys, os, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bliss_gnn_b200 import _native, ops, model as M
from bliss_gnn_b200.graph import synthetic_graph, normalized_edata
from bliss_gnn_b200.parallel import FlatGrads
from bliss_gnn_b200.sampler import PoissonBanditLadiesSampler
from oracle import model as omodel, samplers as osamp
from tests.util import copy_params, rel_to_max, philox_uniform_fn, blocks_as
_native.build()
torch.set_float32_matmul_precision("highest")
dev = torch.device("cuda:0")
gd = synthetic_graph("reddit", seed=0, device=dev)
gd.edata["w"] = normalized_edata(gd)
g = gd.to("cpu")
fan = [4096, 2048, 1024]
train = torch.nonzero(g.ndata["train_mask"], as_tuple=True)[0]
seeds = train[torch.randperm(train.numel(), generator=torch.Generator().manual_seed(1))[:256]]
ora = osamp.PoissonBanditLadiesSampler(fan, eta=0.1, accum="contract", uniform_fn=philox_uniform_fn(2, 0))
o_in, _, ob = ora.sample_blocks(g, seeds)
blocks_as(ob, torch.float64)
dsm = PoissonBanditLadiesSampler(fan, eta=0.1, rng_seed=2)
labels = g.ndata["labels"][seeds.long()]
torch.manual_seed(3)
base = M.SAGE(602, 256, 41, 3, F.relu, 0.0).to(dev)
om = omodel.SAGE(602, 256, 41, 3, F.relu, 0.0)
copy_params(om, base, torch.float64); om = om.double()
lo = F.cross_entropy(om(ob, g.ndata["features"][o_in].double()), labels); lo.backward()
og = {n: p.grad for n, p in om.named_parameters()}
import copy
def run(tag, flat=False, fused_xent=False, pad=False, side=True):
    dsm.step = 0
    d_in, _, db = dsm.sample_blocks(gd, seeds)
    mdl = copy.deepcopy(base)
    if not side:
        M._side_stream = lambda device: None
    if flat:
        fg = FlatGrads(mdl.parameters())
    feats = gd.ndata["features"]
    if pad:
        feats = F.pad(feats, (0, 2)).contiguous()
    x = ops.gather_rows(feats, d_in.to(torch.int32))
    y = mdl(db, x)
    loss = (ops.cross_entropy_mean if fused_xent else F.cross_entropy)(y, labels.to(dev))
    loss.backward()
    torch.cuda.synchronize()
    print(tag, "loss", float(loss), float(lo))
    for n, p in mdl.named_parameters():
        print(f"   {n:28s} {rel_to_max(p.grad, og[n]):.2e}")
run("A direct")
run("B fused xent", fused_xent=True)
run("C flat grads", flat=True)
run("D flat + pad + xent", flat=True, fused_xent=True, pad=True)
