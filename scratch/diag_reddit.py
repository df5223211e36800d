"""Bisect the Reddit-shape trajectory deviation: device eager step vs the float64 oracle loop, tensor by tensor."""
import sys, os, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bliss_gnn_b200 import _native
from bliss_gnn_b200.graph import synthetic_graph, normalized_edata
from bliss_gnn_b200.train import DataModule, Trainer, build_model
from oracle import model as omodel
from tests.util import OracleLoop, copy_params, rel_to_max
_native.build()
torch.set_float32_matmul_precision("highest")
shape = sys.argv[1] if len(sys.argv) > 1 else "reddit"
batch, fan, hidden = (256, [4096, 2048, 1024], 256)
dev = torch.device("cuda:0")
gd = synthetic_graph(shape, seed=0, device=dev)
gd.edata["w"] = normalized_edata(gd)
g = gd.to("cpu")
dm = DataModule(shape, fan_out=fan, eta=0.1, device=dev, batch_size=batch, sampler="poisson-bandit", model="sage", seed=0, graph=gd)
torch.manual_seed(3)
model = build_model("sage", dm.in_feats, hidden, dm.n_classes, 3, dropout=0.0).to(dev)
om = omodel.SAGE(g.ndata["features"].shape[1], hidden, g.n_classes, 3, F.relu, 0.0)
copy_params(om, model, torch.float64); om = om.double()
tr = Trainer(dm, model, 0.002, static_graph=False)
loop = OracleLoop(g, om, "PoissonBanditLadiesSampler", fan, rng_seed=dm.sampler.rng_seed, eta=0.1, lr=0.002)
batches = [b for _, b in zip(range(4), dm.train_batches())]
# hook: capture device grads before the optimizer step
grabbed = {}
orig = tr._optimizer_step
def hooked():
    grabbed["g"] = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    orig()
tr._optimizer_step = hooked
oorig = loop.opt.step
def ohook(*a, **k):
    grabbed["og"] = {n: p.grad.detach().clone() for n, p in om.named_parameters()}
    return oorig(*a, **k)
loop.opt.step = ohook
for s, seeds in enumerate(batches):
    ld = float(tr.training_step(seeds).item()); lo = loop.training_step(seeds)
    print(f"step {s}: loss dev {ld:.7f} ora {lo:.7f} rel {abs(ld-lo)/lo:.2e}")
    for n in grabbed["g"]:
        print(f"   grad {n:28s} max/max {rel_to_max(grabbed['g'][n], grabbed['og'][n]):.2e}   param {rel_to_max(dict(model.named_parameters())[n], dict(om.named_parameters())[n]):.2e}")
    for l, (a, b) in enumerate(zip(tr.last_blocks, loop.last_blocks)):
        ew = rel_to_max(a.canonical()["edge_weights"], b.canonical()["edge_weights"])
        same = torch.equal(a.srcdata["_ID"].cpu().long(), b.srcdata["_ID"])
        en = rel_to_max(a.srcdata["embed_norm"], b.srcdata["embed_norm"])
        print(f"   block {l}: same src {same} E {a.num_edges()} edge_w {ew:.2e} embed_norm {en:.2e}")
    w_dev, w_ora = dm.sampler.exp3_weights.cpu().double(), loop.smp.exp3_weights.double()
    print("   exp3 max rel", ((w_dev - w_ora).abs() / w_ora).max().item())
