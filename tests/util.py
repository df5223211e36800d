"""Shared helpers of the parity tests: small seeded graphs, Philox draws for the oracle, block comparison."""
import numpy as np
import torch

from bliss_gnn_b200.graph import Graph, add_self_loops_and_build, normalized_edata
from oracle import philox


def random_graph(num_nodes, num_edges, seed, hubs=0, hub_degree=0, symmetric=True):
    """Seeded random simple graph with self-loops (reference preprocessing, train_lightning.py:334-335);
    ``hubs`` nodes get ``hub_degree`` extra in-edges so heavy rows (> 256 / > 8192 edges) are exercised."""
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, num_nodes, (num_edges,), generator=g)
    dst = torch.randint(0, num_nodes, (num_edges,), generator=g)
    for h in range(hubs):
        extra = torch.randperm(num_nodes, generator=g)[:hub_degree]
        src = torch.cat([src, extra])
        dst = torch.cat([dst, torch.full((hub_degree,), h, dtype=torch.int64)])
    if symmetric:
        src, dst = torch.cat([src, dst]), torch.cat([dst, src])
    key = torch.unique(src * num_nodes + dst)
    perm = torch.randperm(key.numel(), generator=g)       # edge ids not sorted by endpoint
    key = key[perm]
    gr = add_self_loops_and_build(key // num_nodes, key % num_nodes, num_nodes)
    gr.edata["w"] = normalized_edata(gr)
    return gr


def philox_uniform_fn(seed, step):
    """uniform_fn for the oracle samplers reproducing the device's draws (csrc/common.cuh)."""
    def fn(layer, nids, prob=None):
        return torch.from_numpy(philox.uniform_for_nodes(seed, step, layer, nids.numpy()))
    return fn


class SafeDraws:
    """Philox draws nudged out of the tie band |u - P| <= band·P around the oracle's inclusion
    probabilities, recorded per layer as dense [|V|] arrays to inject into the device sampler
    (SURVEY.md §7 hard part 3: summation order differs, so ties are excluded by the generator)."""

    def __init__(self, num_nodes, seed, step, band=1e-4):
        self.num_nodes, self.seed, self.step, self.band = num_nodes, seed, step, band
        self.per_layer = {}

    def __call__(self, layer, nids, prob=None):
        u = torch.from_numpy(philox.uniform_for_nodes(self.seed, self.step, layer, nids.numpy())).clone()
        if prob is not None:
            p = prob.to(torch.float32)
            close = (u - p).abs() <= self.band * p.clamp(min=1e-30)
            u = torch.where(close & (p < 1), (p * (1 - 4 * self.band)).clamp(min=0), u)
        dense = torch.full((self.num_nodes,), 0.5, dtype=torch.float32)
        dense[nids.long()] = u
        self.per_layer[layer] = dense
        return u


def assert_blocks_equal(dev_block, ora_block, rtol=1e-5, exact_values=False, check=("edge_weights", "q_ij")):
    """Bit-exact structure (node order, canonical edge list, edge ids), values within rtol."""
    d_src = dev_block.srcdata["_ID"].cpu().long()
    o_src = ora_block.srcdata["_ID"].long()
    assert torch.equal(d_src, o_src), "block source node order differs"
    assert torch.equal(dev_block.dstdata["_ID"].cpu().long(), ora_block.dstdata["_ID"].long())
    dc, oc = dev_block.canonical(), ora_block.canonical()
    assert torch.equal(dc["src"].cpu().long(), oc["src"]), "edge sources differ"
    assert torch.equal(dc["dst"].cpu().long(), oc["dst"]), "edge destinations differ"
    assert torch.equal(dc["_ID"].cpu().long(), oc["_ID"].long()), "edge ids differ"
    stats = {}
    for k in check:
        if k not in oc:
            continue
        a, b = dc[k].cpu().double(), oc[k].double()
        if exact_values:
            assert torch.equal(dc[k].cpu(), oc[k].to(dc[k].dtype)), f"{k} not bit-exact"
        rel = ((a - b).abs() / b.abs().clamp(min=1e-30)).max().item() if a.numel() else 0.0
        assert rel <= rtol, f"{k}: max rel err {rel}"
        stats[k] = rel
    if "node_prob" in ora_block.srcdata:
        a, b = dev_block.srcdata["node_prob"].cpu().double(), ora_block.srcdata["node_prob"].double()
        rel = ((a - b).abs() / b.abs().clamp(min=1e-30)).max().item()
        assert rel <= rtol, f"node_prob: max rel err {rel}"
        if exact_values:
            assert torch.equal(dev_block.srcdata["node_prob"].cpu(), ora_block.srcdata["node_prob"].float())
        stats["node_prob"] = rel
    return stats


# ---------------------------------------------------------------------------------------------
# helpers of the configuration / trajectory parity tests
# ---------------------------------------------------------------------------------------------
import json
import os

_STATS = {}
_STATS_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_stats.json")


def record(name, value):
    """Measured parity errors, dumped to gpurun_out/parity_stats.json (copied into profiles/ by hand)."""
    _STATS[name] = value
    try:
        os.makedirs(os.path.dirname(_STATS_PATH), exist_ok=True)
        old = {}
        if os.path.exists(_STATS_PATH):
            try:
                old = json.load(open(_STATS_PATH))
            except Exception:
                old = {}
        old.update(_STATS)
        json.dump(old, open(_STATS_PATH, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


def rel_to_max(a, b):
    """max |a - b| / max |b| (cancelling sums do not produce false alarms)."""
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.numel() == 0:
        return 0.0
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()


def close(a, b, rtol, what):
    err = rel_to_max(a, b)
    record(what, err)
    assert err <= rtol, f"{what}: max err / max|ref| = {err:.3e} > {rtol}"
    return err


def close_grad(dev, ref64, ref32, rtol, what):
    """Gradient parity: the bar is ``rtol`` against the float64 oracle.  Where the chained fp32 backward pass cannot
    hold it — softmax / long signed sums cancel — the SAME comparison is made for the oracle evaluated in float32
    (plain torch fp32 autograd, the reference's own arithmetic) and the device must be within twice that error:
    the tolerance is stated per tensor by measurement, and both numbers are recorded."""
    e_dev, e_32 = rel_to_max(dev, ref64), rel_to_max(ref32, ref64)
    record(what, {"device_vs_fp64": e_dev, "torch_fp32_vs_fp64": e_32})
    assert e_dev <= max(rtol, 2.0 * e_32), (f"{what}: device {e_dev:.3e} vs fp64 oracle; torch fp32 itself is "
                                            f"{e_32:.3e} away; bar {rtol}")
    return e_dev


def blocks_clone(blocks, dtype):
    """Deep-enough copy of oracle blocks with float payloads cast to ``dtype`` (structure shared)."""
    import copy
    out = []
    for b in blocks:
        c = copy.copy(b)
        c.edata, c.srcdata, c.dstdata = dict(b.edata), dict(b.srcdata), dict(b.dstdata)
        for frame in (c.edata, c.srcdata, c.dstdata):
            for k, v in list(frame.items()):
                if torch.is_tensor(v) and v.is_floating_point():
                    frame[k] = v.detach().to(dtype)
        out.append(c)
    return out


def blocks_as(blocks, dtype):
    """The oracle's blocks with their float payloads cast to ``dtype`` (so an fp64 oracle model sees exactly
    the numbers the device model sees)."""
    for b in blocks:
        for frame in (b.edata, b.srcdata, b.dstdata):
            for k, v in list(frame.items()):
                if torch.is_tensor(v) and v.is_floating_point():
                    frame[k] = v.to(dtype)
    return blocks


def copy_params(dst_model, src_model, dtype=None):
    sd = {k: v.detach().cpu().clone() for k, v in src_model.state_dict().items()}
    if dtype is not None:
        sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    missing = dst_model.load_state_dict(sd, strict=False)
    assert not [k for k in missing.missing_keys if "fc_dst" not in k], missing


class OracleLoop:
    """The reference's training loop restated on the CPU oracle: ``train_lightning.py:100-168`` (sample → lazy
    feature / label fetch → forward → loss → backward → Adam) followed by ``BatchSizeCallback.on_train_batch_end``
    (``:463-471``: ``sampler.exp3``).  Draws are the device's Philox stream (seed, step = number of
    ``sample_blocks`` calls so far)."""

    def __init__(self, g_cpu, model, sampler_cls, fan, rng_seed, eta=0.1, lr=0.002, multilabel=False,
                 model_kind="sage", accum="contract", dtype=torch.float32):
        from oracle import samplers as osamp
        self.g, self.model, self.step, self.seed = g_cpu, model, 0, rng_seed
        kw = dict(eta=eta, model=model_kind) if "Bandit" in sampler_cls else {}
        self.smp = getattr(osamp, sampler_cls)(list(fan), accum=accum, dtype=dtype, uniform_fn=self._draw, **kw)
        self.opt = torch.optim.Adam(model.parameters(), lr=lr)
        self.multilabel = multilabel
        self.bandit = "Bandit" in sampler_cls
        self.mdtype = next(model.parameters()).dtype

    def _draw(self, layer, nids, prob=None):
        return torch.from_numpy(philox.uniform_for_nodes(self.seed, self.step, layer, nids.numpy()))

    def training_step(self, seeds):
        import torch.nn.functional as F
        g = self.g
        inp, _, blocks = self.smp.sample_blocks(g, seeds.cpu())
        self.step += 1
        if self.mdtype != torch.float32:
            blocks_as(blocks, self.mdtype)
        x = g.ndata["features"][inp].to(self.mdtype)
        y = g.ndata["labels"][seeds.cpu().long()]
        pred = self.model(blocks, x)
        loss = F.binary_cross_entropy_with_logits(pred, y.to(self.mdtype)) if self.multilabel else F.cross_entropy(pred, y)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        if self.bandit:
            if self.mdtype != torch.float32:          # the sampler state stays in its own dtype
                for b in blocks:
                    for k in ("embed_norm",):
                        b.srcdata[k] = b.srcdata[k].detach().to(self.smp.dtype)
                    for k in ("q_ij", "a_ij", "w"):
                        if k in b.edata:
                            b.edata[k] = b.edata[k].detach().to(self.smp.dtype)
                    b.srcdata["node_prob"] = b.srcdata["node_prob"].to(self.smp.dtype)
            self.smp.exp3(blocks, g)
        self.last_blocks = blocks
        return float(loss.item())


def oracle_fit(g_cpu, fan, batch, hidden, n_steps, seed, lr=0.002, dropout=0.1, eta=0.1):
    """The reference's whole run restated on the CPU oracle (``train_lightning.py:562-705``): training steps with the
    bandit update after each, StepLR(gamma=0.01, step_size=5) per EPOCH (``:205-216``), validation after every epoch
    with the same stochastic sampler (``:179-203,410-422``), best-val checkpoint (``:622-625``) reloaded before the
    layer-wise full-neighbour inference (``model.py:335-383``) and the final micro-F1 (``:686-705``)."""
    import copy
    import torch.nn.functional as F
    from bliss_gnn_b200.train import DataModule
    from oracle import dglops, model as omodel
    dm = DataModule("oracle", fan_out=fan, eta=eta, device=torch.device("cpu"), batch_size=batch, sampler="poisson-bandit",
                    model="sage", seed=seed, graph=g_cpu)
    torch.manual_seed(seed + 3)
    feats = g_cpu.ndata["features"]
    model = omodel.SAGE(feats.shape[1], hidden, g_cpu.n_classes, len(fan), F.relu, dropout)
    loop = OracleLoop(g_cpu, model, "PoissonBanditLadiesSampler", fan, rng_seed=dm.sampler.rng_seed & 0xFFFFFFFF, eta=eta,
                      lr=lr)
    sched = torch.optim.lr_scheduler.StepLR(loop.opt, gamma=0.01, step_size=5)
    labels = g_cpu.ndata["labels"]

    def validate():
        model.eval()
        hit = tot = 0
        with torch.no_grad():
            for seeds in dm.val_batches():
                inp, _, blocks = loop.smp.sample_blocks(g_cpu, seeds.cpu())
                loop.step += 1
                pred = model(blocks, feats[inp])
                hit += int((pred.argmax(1) == labels[seeds.long()]).sum())
                tot += seeds.numel()
        model.train()
        return hit / max(tot, 1)

    best, state, step, done = -1.0, None, 0, False
    while not done:
        for seeds in dm.train_batches():
            loop.training_step(seeds)
            step += 1
            if step >= n_steps:
                done = True
                break
        sched.step()
        val = validate()
        if val > best:
            best, state = val, copy.deepcopy(model.state_dict())
    model.load_state_dict(state)
    model.eval()
    src, dst = g_cpu.coo()
    order = torch.sort(dst, stable=True).indices
    full = dglops.OBlock(src[order], dst[order], g_cpu.num_nodes(), g_cpu.num_nodes())
    h = feats
    with torch.no_grad():
        for l, layer in enumerate(model.layers):
            h = layer(full, h)
            if l < len(model.layers) - 1:
                h = F.relu(h)
    test = torch.nonzero(g_cpu.ndata["test_mask"], as_tuple=True)[0]
    return float((h[test].argmax(1) == labels[test]).float().mean())
