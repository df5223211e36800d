"""GPU parity tests of the aggregation kernels and the drop-in models against the CPU oracle
(plain torch fp32, autograd for the backward).  Tolerance: 1e-5 relative (north-star; summation
order differs), measured against the largest magnitude of the compared tensor so cancelling
sums do not produce false alarms."""
import pytest
import torch
import torch.nn.functional as F

from oracle import model as omodel
from oracle import samplers as osamp
from tests.util import blocks_as, blocks_clone, close_grad, copy_params, philox_uniform_fn, random_graph

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _close(a, b, rtol=RTOL, what=""):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    scale = b.abs().max().clamp(min=1e-30)
    err = ((a - b).abs().max() / scale).item()
    assert err <= rtol, f"{what}: max err / max|ref| = {err}"


def _sample(model="sage", fan=(128, 64, 32), V=1500, E=9000, batch=24, hubs=3, hub_degree=600):
    from bliss_gnn_b200.sampler import PoissonBanditLadiesSampler
    g = random_graph(V, E, seed=13, hubs=hubs, hub_degree=hub_degree)
    seeds = torch.randperm(V, generator=torch.Generator().manual_seed(2))[:batch]
    ora = osamp.PoissonBanditLadiesSampler(list(fan), eta=0.1, model=model, accum="contract",
                                           uniform_fn=philox_uniform_fn(7, 0))
    _, _, ob = ora.sample_blocks(g, seeds)
    gd = g.to(_dev())
    dev = PoissonBanditLadiesSampler(list(fan), eta=0.1, model=model, rng_seed=7)
    _, _, db = dev.sample_blocks(gd, seeds)
    return g, gd, ob, db


@pytest.mark.parametrize("dim", [1, 7, 41, 64, 256, 602, 1433])
def test_gather_rows_and_norm(native_lib, dim):
    from bliss_gnn_b200 import ops
    dev = _dev()
    table = torch.randn(1000, dim, device=dev)
    nid = torch.randint(0, 1000, (333,), device=dev, dtype=torch.int32)
    out, norm = ops.gather_rows(table, nid, with_norm=True)
    assert torch.equal(out, table[nid.long()])                      # a copy: bit-exact
    _close(norm, torch.linalg.norm(table[nid.long()], dim=1), what="row norm")
    _close(ops.row_norm(table), torch.linalg.norm(table, dim=1), what="row_norm")


@pytest.mark.parametrize("dim", [1, 7, 41, 64, 256, 300, 1030])
def test_spmm_forward_backward(native_lib, dim):
    from bliss_gnn_b200 import ops
    g, gd, ob, db = _sample()
    for blk, oblk in zip(db, ob):
        x = torch.randn(blk.num_src_nodes(), dim, device=gd.device, requires_grad=True)
        w = blk.edata["edge_weights"]
        ss = torch.rand(blk.num_src_nodes(), device=gd.device) + 0.5
        ds = torch.rand(blk.num_dst_nodes(), device=gd.device) + 0.5
        y = ops.spmm(blk, x, w, src_scale=ss, dst_scale=ds)
        gy = torch.randn_like(y)
        y.backward(gy)
        xr = x.detach().cpu().clone().requires_grad_(True)
        src, dst = blk.edge_src.cpu().long(), blk.edge_dst.cpu().long()
        m = xr[src] * (w.cpu() * ss.cpu()[src]).unsqueeze(1)
        yr = torch.zeros(blk.num_dst_nodes(), dim).index_add(0, dst, m) * ds.cpu().unsqueeze(1)
        yr.backward(gy.cpu())
        _close(y, yr, what=f"spmm fwd D={dim}")
        _close(x.grad, xr.grad, what=f"spmm bwd D={dim}")


@pytest.mark.parametrize("env", [{"BLISS_SPMM_MODE": "group"}, {"BLISS_SPMM_TMA": "1"}, {"BLISS_SPMM_LB": "5"}])
def test_spmm_variants_agree(native_lib, env):
    """The SpMM variants kept for the ablation (profiles/r2_spmm_variants.md) — round-1 group kernel, cp.async.bulk
    ring, other register budget — are selected by environment variables read once per process, so each runs in a
    child process: forward and transposed SpMM over blocks with 600-edge rows against a float64 reference (1e-5)."""
    import os
    import subprocess
    import sys
    code = r"""
import sys, torch
sys.path.insert(0, %r)
from tests.test_gpu_ops import _sample, _close
from bliss_gnn_b200 import ops
g, gd, ob, db = _sample(fan=(512, 256, 64), V=3000, E=20000, batch=48, hubs=4, hub_degree=1500)
for blk in db:
    for dim in (256, 128, 64, 300):
        x = torch.randn(blk.num_src_nodes(), dim, device=gd.device, requires_grad=True)
        w = blk.edata["edge_weights"]
        y = ops.spmm(blk, x, w, dst_scale=ops.mean_scale(blk))
        gy = torch.randn_like(y)
        y.backward(gy)
        src, dst = blk.edge_src.long(), blk.edge_dst.long()
        xr = x.detach().double().requires_grad_(True)
        yr = torch.zeros(blk.num_dst_nodes(), dim, device=gd.device, dtype=torch.float64).index_add(
            0, dst, xr[src] * w.double().unsqueeze(1)) * ops.mean_scale(blk).double().unsqueeze(1)
        yr.backward(gy.double())
        _close(y, yr, what="fwd")
        _close(x.grad, xr.grad, what="bwd")
print("ok")
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),)
    r = subprocess.run([sys.executable, "-c", code], env={**os.environ, **env}, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-2000:]


def test_block_transpose_is_sorted(native_lib):
    from bliss_gnn_b200 import ops
    _, gd, _, db = _sample()
    for blk in db:
        t_indptr, t_dst, t_perm, _ = ops.block_transpose(blk)
        assert int(t_indptr[-1]) == blk.num_edges()
        src_of = blk.edge_src[t_perm.long()]
        rows = torch.repeat_interleave(torch.arange(blk.num_src_nodes(), device=gd.device),
                                       (t_indptr[1:] - t_indptr[:-1]).long())
        assert torch.equal(src_of.long(), rows)                     # grouped by source
        assert torch.equal(blk.edge_dst[t_perm.long()], t_dst)
        key = rows * (blk.num_edges() + 1) + t_perm.long()          # ascending edge id inside each row
        assert torch.all(key[1:] > key[:-1])


def _copy_params(dst_model, src_model):
    sd = {k: v.detach().cpu().clone() for k, v in src_model.state_dict().items()}
    missing = dst_model.load_state_dict(sd, strict=False)
    assert not [k for k in missing.missing_keys if "fc_dst" not in k], missing


@pytest.mark.parametrize("kind", ["sage", "gcn"])
def test_sage_gcn_models_forward_backward(native_lib, kind):
    from bliss_gnn_b200 import model as M
    g, gd, ob, db = _sample()
    in_f, hid, ncls = 37, 64, 5          # in < hidden -> aggregate first; hidden > classes -> project first
    torch.manual_seed(0)
    feats = torch.randn(g.num_nodes(), in_f)
    dmodel = (M.SAGE if kind == "sage" else M.GCN)(in_f, hid, ncls, 3, F.relu, 0.0).to(gd.device)
    # the oracle runs in float64 (exact-arithmetic reference): the error measured is the device's own fp32 rounding
    omod = (omodel.SAGE if kind == "sage" else omodel.GCN)(in_f, hid, ncls, 3, F.relu, 0.0)
    copy_params(omod, dmodel, torch.float64)
    omod32 = (omodel.SAGE if kind == "sage" else omodel.GCN)(in_f, hid, ncls, 3, F.relu, 0.0)
    copy_params(omod32, dmodel)
    omod = omod.double()
    ob32 = blocks_clone(ob, torch.float32)
    blocks_as(ob, torch.float64)
    xd = feats.to(gd.device)[db[0].srcdata["_ID"].long()]
    xo = feats[ob[0].srcdata["_ID"].long()].double()
    yd = dmodel(db, xd)
    yo = omod(ob, xo)
    _close(yd, yo, what=f"{kind} logits")
    for a, b in zip(db, ob):
        _close(a.srcdata["embed_norm"], b.srcdata["embed_norm"], what="embed_norm")
    labels = torch.randint(0, ncls, (yo.shape[0],))
    F.cross_entropy(yd, labels.to(gd.device)).backward()
    F.cross_entropy(yo, labels).backward()
    F.cross_entropy(omod32(ob32, feats[ob[0].srcdata["_ID"].long()]), labels).backward()
    g32 = dict(omod32.named_parameters())
    for (n, p), (_, q) in zip(dmodel.named_parameters(), omod.named_parameters()):
        close_grad(p.grad, q.grad, g32[n].grad, RTOL, f"ops/{kind} grad {n}")


@pytest.mark.parametrize("residual", [False, True])
def test_gatv2_model_forward_backward(native_lib, residual):
    from bliss_gnn_b200 import model as M
    g, gd, ob, db = _sample(model="gat")
    in_f, hid, ncls, heads = 19, 48, 6, [4, 4, 1]
    torch.manual_seed(1)
    feats = torch.randn(g.num_nodes(), in_f)
    args = (3, in_f, hid, ncls, heads, F.elu, 0.0, 0.0, 0.2, residual)
    dmodel = M.GATv2(*args).to(gd.device)
    omod = omodel.GATv2(*args)
    copy_params(omod, dmodel, torch.float64)
    omod32 = omodel.GATv2(*args)
    copy_params(omod32, dmodel)
    omod = omod.double()                       # float64 oracle: the exact-arithmetic reference
    ob32 = blocks_clone(ob, torch.float32)
    blocks_as(ob, torch.float64)
    yd = dmodel(db, feats.to(gd.device)[db[0].srcdata["_ID"].long()])
    yo = omod(ob, feats[ob[0].srcdata["_ID"].long()].double())
    _close(yd, yo, what="gat logits")
    for a, b in zip(db, ob):
        assert torch.equal(a.edge_src.cpu().long(), b.src)          # same native edge order
        _close(a.edata["a_ij"], b.edata["a_ij"], what="a_ij (head-mean logits)")
    labels = torch.randint(0, ncls, (yo.shape[0],))
    F.cross_entropy(yd, labels.to(gd.device)).backward()
    F.cross_entropy(yo, labels).backward()
    F.cross_entropy(omod32(ob32, feats[ob[0].srcdata["_ID"].long()]), labels).backward()
    dgrads, g32 = dict(dmodel.named_parameters()), dict(omod32.named_parameters())
    for n, q in omod.named_parameters():
        close_grad(dgrads[n].grad, q.grad, g32[n].grad, RTOL, f"ops/gat(residual={residual}) grad {n}")


def test_gatv2_attention_dropout_mask(native_lib):
    """attn_drop (model.py:88-90) enters the fused kernel as a pre-scaled keep mask."""
    from bliss_gnn_b200 import ops
    from oracle import dglops
    _, gd, ob, db = _sample(model="gat")
    blk, oblk = db[0], ob[0]
    H, D = 2, 40
    torch.manual_seed(3)
    feat = torch.randn(blk.num_src_nodes(), H, D)
    attn = torch.randn(1, H, D)
    mask = (torch.rand(blk.num_edges(), H) < 0.8).float() / 0.8
    fd = feat.to(gd.device).requires_grad_(True)
    ad = attn.to(gd.device).requires_grad_(True)
    out, logits = ops.gatv2_attention(blk, fd, ad, 0.2, mask.to(gd.device))
    fo, ao = feat.double().requires_grad_(True), attn.double().requires_grad_(True)     # float64 reference
    e = F.leaky_relu(fo[oblk.src] + fo[: oblk.num_dst_nodes()][oblk.dst], 0.2)
    e = (e * ao).sum(-1)
    a = dglops.edge_softmax(oblk, e) * mask.double()
    ref = torch.zeros(oblk.num_dst_nodes(), H, D, dtype=torch.float64).index_add(0, oblk.dst, fo[oblk.src] * a.unsqueeze(-1))
    _close(out, ref, what="gat out with dropout mask")
    _close(logits, e, what="gat logits")
    go = torch.randn(ref.shape)
    out.backward(go.to(gd.device))
    ref.backward(go.double())
    _close(fd.grad, fo.grad, rtol=RTOL, what="gat grad feat")
    _close(ad.grad, ao.grad, rtol=RTOL, what="gat grad attn")


@pytest.mark.parametrize("kind", ["sage", "gcn", "gat"])
def test_full_graph_inference(native_lib, kind):
    """``model.inference``: layer-wise full-neighbour pass without edge weights (SAGE model.py:335-383, GCN :441-488,
    GATv2 :236-289) == the oracle layers applied to the whole graph as one block.  The reference walks the nodes in
    batches of ``batch_size`` through ``MultiLayerFullNeighborSampler(1)``; a node's output depends only on its own
    in-edges and on the previous layer's full matrix, so one whole-graph block per layer is the same function (the
    batch size only bounds the reference's memory) — checked here for two batch sizes as well."""
    from bliss_gnn_b200 import model as M
    from oracle import dglops
    g = random_graph(700, 4000, seed=4, hubs=2, hub_degree=400)
    feats = torch.randn(700, 24, generator=torch.Generator().manual_seed(5))
    g.ndata["features"] = feats
    gd = g.to(_dev())
    torch.manual_seed(2)
    if kind == "gat":
        args = (3, 24, 16, 4, [2, 2, 1], F.elu, 0.0, 0.0, 0.2, True)
        dmodel, omod = M.GATv2(*args).to(gd.device), omodel.GATv2(*args)
    else:
        cls_d, cls_o = (M.SAGE, omodel.SAGE) if kind == "sage" else (M.GCN, omodel.GCN)
        dmodel, omod = cls_d(24, 32, 4, 3, F.relu, 0.0).to(gd.device), cls_o(24, 32, 4, 3, F.relu, 0.0)
    copy_params(omod, dmodel, torch.float64)
    omod = omod.double()
    pred = dmodel.inference(gd, gd.device, 128)
    assert torch.equal(pred, dmodel.inference(gd, gd.device, 7))       # the batch size does not change the function
    src, dst = g.coo()
    order = torch.sort(dst, stable=True).indices
    full = dglops.OBlock(src[order], dst[order], 700, 700)
    h = feats.double()
    with torch.no_grad():
        if kind == "gat":
            for l, layer in enumerate(omod.gatv2_layers):
                h = layer(full, h)
                h = h.flatten(1) if l < 2 else h.mean(1)
        else:
            for l, layer in enumerate(omod.layers):
                h = layer(full, h)
                if l < 2 and kind == "sage":
                    h = F.relu(h)                                     # (GraphConv applies its activation itself)
    _close(pred, h, what=f"{kind} full-graph inference")


@pytest.mark.parametrize("kind,sampler", [("sage", "poisson-bandit"), ("gcn", "poisson-bandit"), ("gat", "poisson-bandit"),
                                          ("sage", "bandit"), ("sage", "ladies"), ("sage", "poisson-ladies"),
                                          ("sage", "poisson-bandit/literal"), ("sage", "neighbor"), ("gcn", "full")])
def test_static_graph_step_matches_eager(native_lib, kind, sampler):
    """Trainer(static_graph=True) — the whole step (sampling with every sampler of the CLI, forward, backward,
    Adam, bandit update) as one replayed CUDA graph over capacity-padded blocks — follows the same loss
    trajectory as the eager step (same seeds, same Philox draws, dropout off)."""
    from bliss_gnn_b200.graph import synthetic_graph
    from bliss_gnn_b200.train import DataModule, Trainer, build_model
    dev = _dev()
    g = synthetic_graph("flickr", seed=0, scale=0.05).to(dev)      # 4.5 K nodes, half of them training nodes
    losses = {}
    sampler, _, normalize = sampler.partition("/")       # "<sampler>/literal": dense L1 renormalisation after every update
    for static in (False, True):
        dm = DataModule("flickr", fan_out=[128, 64, 32], eta=0.1, device=dev, batch_size=32, sampler=sampler,
                        model=kind, seed=0, graph=g, normalize=normalize or "lazy")
        torch.manual_seed(3)
        model = build_model(kind, dm.in_feats, 64, dm.n_classes, 3, dropout=0.0, attn_dropout=0.0,
                            faithful_gcn_quirk=False).to(dev)
        tr = Trainer(dm, model, 0.002, static_graph=static, eager_warmup=3)
        out = []
        for step, seeds in zip(range(12), dm.train_batches()):
            out.append(float(tr.training_step(seeds).item()))
        assert len(out) == 12
        if static:
            tr.flush()                       # pipelined counters: the last step's are still outstanding
        assert tr.num_steps == 12 and tr.total_sampled_edges > 0
        losses[static] = out
        bandit = "bandit" in sampler
        if static:
            assert tr.graph_replays >= 7
            w_static = dm.sampler.exp3_weights.clone() if bandit else None
        else:
            w_eager = dm.sampler.exp3_weights.clone() if bandit else None
    for a, b in zip(losses[False], losses[True]):
        assert abs(a - b) <= 2e-4 * max(1.0, abs(a)), (losses[False], losses[True])
    if bandit:
        # GAT's alpha divides by sums of signed logits: padded-vs-exact GEMM rounding is amplified there
        torch.testing.assert_close(w_static, w_eager, rtol=2e-3 if kind == "gat" else 1e-4, atol=0)


@pytest.mark.parametrize("kind,sampler", [("sage", "poisson-bandit"), ("gat", "poisson-bandit"), ("gcn", "ladies")])
def test_lookahead_sampling_keeps_the_trajectory(native_lib, kind, sampler):
    """``training_step(seeds, next_seeds)``: the next batch's blocks are sampled into the second pool set in the shadow
    of this step's backward pass (after this step's exp3) — same losses and bandit weights as the step that samples
    its own blocks first, including across an eager step in between (ragged batch) and an unannounced batch."""
    from bliss_gnn_b200.graph import synthetic_graph
    from bliss_gnn_b200.train import DataModule, Trainer, build_model
    dev = _dev()
    g = synthetic_graph("flickr", seed=0, scale=0.05).to(dev)
    res = {}
    for ahead in (False, True, "host"):
        dm = DataModule("flickr", fan_out=[128, 64, 32], eta=0.1, device=dev, batch_size=32, sampler=sampler,
                        model=kind, seed=0, graph=g)
        torch.manual_seed(3)
        model = build_model(kind, dm.in_feats, 64, dm.n_classes, 3, dropout=0.0, attn_dropout=0.0,
                            faithful_gcn_quirk=False).to(dev)
        tr = Trainer(dm, model, 0.002, static_graph=True, eager_warmup=3)
        batches = [b for _, b in zip(range(16), dm.train_batches())]
        batches[9] = batches[9][:20]                       # a ragged batch: runs eagerly, the prefetch must be dropped
        if ahead == "host":        # the data loader's case: seeds arrive as host tensors (pinned staging, copy stream)
            batches = [b.cpu() for b in batches]
        losses, late = [], []
        for i, b in enumerate(batches):
            nxt = batches[i + 1] if (ahead and i + 1 < len(batches) and i != 12) else None    # 12 -> 13 unannounced
            loss = tr.training_step(b, nxt)
            if ahead == "host":    # losses through the step graph's own copy to pinned memory, one step late
                if i:
                    late.append(tr.host_loss(back=1))
                if i == len(batches) - 1:
                    late.append(tr.host_loss())
            losses.append(float(loss.item()))
        tr.flush()
        assert tr.num_steps == len(batches), tr.num_steps
        assert not late or late == losses, (late, losses)
        res[ahead] = (losses, dm.sampler.exp3_weights.clone() if "bandit" in sampler else None, dm.sampler.step)
    assert res[True][2] == res[False][2] == res["host"][2] == 16, "every sampling must consume exactly one Philox step"
    for a, b, c in zip(res[False][0], res[True][0], res["host"][0]):
        assert abs(a - b) <= 1e-6 * max(1.0, abs(a)), (res[False][0], res[True][0])
        assert abs(a - c) <= 1e-6 * max(1.0, abs(a)), (res[False][0], res["host"][0])
    if res[True][1] is not None:
        # GAT: the attn gradient is accumulated with atomics (run-to-run last-bit differences) and alpha divides by
        # sums of signed logits, which amplifies them (see test_static_graph_step_matches_eager)
        torch.testing.assert_close(res[True][1], res[False][1], rtol=2e-3 if kind == "gat" else 1e-6, atol=0)


def test_pool_overflow_is_contained_and_reported(native_lib):
    """A block that outgrows its capacity pool inside a replayed step: the index kernel clamps the row extents and the
    counters to the capacity (the aggregation kernels of the same replay stay in bounds), the fill is skipped, and the
    host raises when it reads the step's counters — the device stays usable."""
    from bliss_gnn_b200.graph import synthetic_graph
    from bliss_gnn_b200.train import DataModule, Trainer, build_model
    dev = _dev()
    g = synthetic_graph("flickr", seed=0, scale=0.3).to(dev)
    dm = DataModule("flickr", fan_out=[512, 256, 128], eta=0.1, device=dev, batch_size=64, sampler="poisson-bandit",
                    model="sage", seed=0, graph=g)
    torch.manual_seed(3)
    model = build_model("sage", dm.in_feats, 64, dm.n_classes, 3, dropout=0.0).to(dev)
    tr = Trainer(dm, model, 0.002, static_graph=True, eager_warmup=3)
    batches = [b for _, b in zip(range(12), dm.train_batches())]
    for b in batches[:3]:
        tr.training_step(b)                      # eager steps that size the pools
    tr.POOL_EDGE_FACTOR, tr.POOL_EDGE_SLACK = 0.5, 0            # pools far too small for the blocks
    with pytest.raises(RuntimeError, match="capacity of layer"):
        for i, b in enumerate(batches[3:]):
            tr.training_step(b, batches[4 + i] if 4 + i < len(batches) else None)
        tr.flush()
    torch.cuda.synchronize()                     # no illegal access happened on the way
    for pset in tr._sets:
        for pool in pset.pools:
            assert int(pool.indptr.max()) <= pool.cap_edges and int(pool.src_nid.max()) < g.num_nodes()
    assert float((torch.ones(8, device=dev) * 2).sum()) == 16.0


def test_data_parallel_graph_path_on_one_rank(native_lib, monkeypatch):
    """The data-parallel step (graph A: sample+fwd+bwd+reward emit -> NCCL all-reduce / all-gather ->
    graph B: Adam + packed apply) forced on a 1-rank NCCL group follows the single-graph trajectory."""
    import torch.distributed as dist
    from bliss_gnn_b200.graph import synthetic_graph
    from bliss_gnn_b200.train import DataModule, Trainer, build_model
    dev = _dev()
    g = synthetic_graph("flickr", seed=0, scale=0.05).to(dev)
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29571", rank=0, world_size=1, device_id=dev)
    try:
        res = {}
        for dp in (False, True):
            if dp:
                monkeypatch.setenv("BLISS_FORCE_DP_PATH", "1")
            dm = DataModule("flickr", fan_out=[128, 64, 32], eta=0.1, device=dev, batch_size=32,
                            sampler="poisson-bandit", model="sage", seed=0, graph=g)
            torch.manual_seed(3)
            model = build_model("sage", dm.in_feats, 64, dm.n_classes, 3, dropout=0.0).to(dev)
            tr = Trainer(dm, model, 0.002, process_group=dist.group.WORLD if dp else None, static_graph=True,
                         eager_warmup=3)
            batches = [b for _, b in zip(range(12), dm.train_batches())]
            losses = [float(tr.training_step(b, batches[i + 1] if dp and i + 1 < 12 else None).item())
                      for i, b in enumerate(batches)]       # the data-parallel run also samples one batch ahead
            assert tr.graph_replays >= 7 and (tr._exchange is not None) == dp
            res[dp] = (losses, dm.sampler.exp3_weights.clone())
        for a, b in zip(res[False][0], res[True][0]):
            assert abs(a - b) <= 2e-4 * max(1.0, abs(a)), (res[False][0], res[True][0])
        torch.testing.assert_close(res[True][1], res[False][1], rtol=1e-4, atol=0)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["sage", "gcn", "gat"])
def test_padded_feature_rows_give_the_same_model_output(native_lib, kind):
    """DataModule pads the feature rows to a 16-byte multiple; the models zero-pad their first weight
    view — logits and gradients are unchanged."""
    from bliss_gnn_b200 import model as M
    g, gd, ob, db = _sample(model=kind if kind == "gat" else "sage")
    in_f, hid, ncls = 37, 64, 5
    torch.manual_seed(0)
    feats = torch.randn(g.num_nodes(), in_f, device=gd.device)
    if kind == "gat":
        mdl = M.GATv2(3, in_f, hid, ncls, [2, 2, 1], F.elu, 0.0, 0.0, 0.2, True).to(gd.device)
    else:
        mdl = (M.SAGE if kind == "sage" else M.GCN)(in_f, hid, ncls, 3, F.relu, 0.0).to(gd.device)
    nid = db[0].srcdata["_ID"].long()
    y0 = mdl(db, feats[nid])
    y0.sum().backward()
    g0 = [p.grad.clone() for p in mdl.parameters()]
    mdl.zero_grad()
    y1 = mdl(db, F.pad(feats, (0, 3))[nid])
    y1.sum().backward()
    _close(y1, y0, what="padded logits")
    for a, b in zip(mdl.parameters(), g0):
        _close(a.grad, b, rtol=2e-5, what="padded grads")


def test_flat_adam_matches_torch_adam(native_lib):
    """``parallel.FlatAdam`` (one ``bliss_adam_step`` launch over flat buffers) follows
    ``torch.optim.Adam`` — the reference's optimizer, ``train_lightning.py:206`` — step by step,
    including an lr change between steps and the cleared gradient buffer."""
    from bliss_gnn_b200.parallel import FlatAdam, FlatGrads
    dev = _dev()
    torch.manual_seed(0)
    shapes = [(37, 19), (19,), (5, 37), (1,), (64, 64)]
    ref_p = [torch.randn(s, device=dev, requires_grad=True) for s in shapes]
    my_p = [p.detach().clone().requires_grad_(True) for p in ref_p]
    ref_opt = torch.optim.Adam(ref_p, lr=0.002)
    fg = FlatGrads(my_p)
    my_opt = FlatAdam(fg, lr=0.002)
    for step in range(25):
        if step == 10:
            for o in (ref_opt, my_opt):
                o.param_groups[0]["lr"] = 0.0007
        grads = [torch.randn(s, device=dev) * (0.1 + step) for s in shapes]
        for p, q, g in zip(ref_p, my_p, grads):
            p.grad = g.clone()
            q.grad.add_(g)                      # accumulate into the (cleared) flat buffer like autograd does
        ref_opt.step()
        my_opt.step()
        assert float(fg.flat.abs().max()) == 0.0, "the step must leave the gradient buffer cleared"
        for p, q in zip(ref_p, my_p):
            _close(q.detach(), p.detach(), rtol=2e-6, what=f"adam step {step}")
    assert int(my_opt.step_dev.item()) == 25


@pytest.mark.parametrize("dim,p", [(64, 0.0), (256, 0.0), (256, 0.3), (1024, 0.1), (36, 0.5)])
def test_sage_epilogue_matches_torch(native_lib, dim, p):
    """``ops.sage_epilogue`` = dropout(relu(a + b + bias)) with row norms (``model.py:321-332,318``): exact
    against torch without dropout; with dropout the kept elements, the drop rate and the backward pass
    (through the realised mask) are checked."""
    from bliss_gnn_b200 import ops
    dev = _dev()
    torch.manual_seed(dim)
    n = 777
    a = torch.randn(n, dim, device=dev, requires_grad=True)
    b = torch.randn(n, dim, device=dev, requires_grad=True)
    bias = torch.randn(dim, device=dev, requires_grad=True)
    step = torch.full((1,), 5, dtype=torch.int64, device=dev)
    y, norm = ops.sage_epilogue(a, b, bias, True, p, seed=11, step_dev=step if p > 0 else None, layer=1)
    z = torch.relu(a.detach() + b.detach() + bias.detach())
    if p == 0.0:
        assert torch.equal(y, z)
        mask = torch.ones_like(z)
    else:
        mask = ((y != 0) | (z == 0)).float()
        kept = mask.bool() & (z > 0)
        _close(y[kept], z[kept] / (1 - p), rtol=2e-7, what="kept elements")
        rate = 1.0 - mask[z > 0].mean().item()
        assert abs(rate - p) < 0.02, rate
        y2, _ = ops.sage_epilogue(a, b, bias, True, p, seed=11, step_dev=step, layer=1)
        assert torch.equal(y, y2), "same (seed, step, layer) must give the same mask"
        step.add_(1)
        y3, _ = ops.sage_epilogue(a, b, bias, True, p, seed=11, step_dev=step, layer=1)
        assert not torch.equal(y, y3)
    _close(norm, y.detach().norm(dim=1), rtol=2e-6, what="row norms")
    gy = torch.randn_like(y)
    y.backward(gy)
    a2 = a.detach().clone().requires_grad_(True)
    b2 = b.detach().clone().requires_grad_(True)
    bias2 = bias.detach().clone().requires_grad_(True)
    ref = torch.relu(a2 + b2 + bias2) * mask / (1 - p)
    ref.backward(gy)
    _close(a.grad, a2.grad, rtol=2e-7, what="grad a")
    _close(b.grad, b2.grad, rtol=2e-7, what="grad b")
    _close(bias.grad, bias2.grad, rtol=2e-5, what="grad bias")


@pytest.mark.parametrize("n,c", [(256, 41), (32, 7), (1000, 100), (5, 3)])
def test_cross_entropy_mean_matches_torch(native_lib, n, c):
    """``ops.cross_entropy_mean`` == ``nn.CrossEntropyLoss()`` (``train_lightning.py:77-79``), value and gradient."""
    from bliss_gnn_b200 import ops
    dev = _dev()
    torch.manual_seed(n + c)
    x = (torch.randn(n, c, device=dev) * 3).requires_grad_(True)
    y = torch.randint(0, c, (n,), device=dev)
    loss = ops.cross_entropy_mean(x, y)
    (loss * 1.7).backward()
    x2 = x.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(x2, y)
    (ref * 1.7).backward()
    _close(loss.reshape(1), ref.reshape(1), rtol=2e-6, what="loss")
    _close(x.grad, x2.grad, rtol=2e-6, what="grad")


@pytest.mark.parametrize("rows,in_f,pad", [(4096, 602, 2), (3001, 256, 0), (2048, 37, 3)])
def test_linear_splitk_matches_plain(native_lib, rows, in_f, pad):
    """``model._linear`` on >= 2048 rows: the split-K weight gradient (batched GEMM over row chunks +
    ``bliss_splitk_accumulate`` into an existing ``weight.grad``, or a returned sum when there is none) equals
    the plain ``F.linear`` autograd result, with and without zero-padded input columns."""
    from bliss_gnn_b200 import model as M
    dev = _dev()
    torch.manual_seed(rows)
    lin = torch.nn.Linear(in_f, 64).to(dev)
    x = torch.randn(rows, in_f, device=dev)
    xp = F.pad(x, (0, pad)) if pad else x
    gy = torch.randn(rows, 64, device=dev)
    ref = F.linear(x, lin.weight, lin.bias)
    ref.backward(gy)
    gw_ref, gb_ref = lin.weight.grad.clone(), lin.bias.grad.clone()
    for preset in (False, True):
        lin.weight.grad = torch.full_like(lin.weight, 0.5) if preset else None     # accumulate into an existing buffer
        lin.bias.grad = None
        xin = xp.clone().requires_grad_(True)
        y = M._linear(xin, lin)
        _close(y, ref, rtol=2e-6, what="forward")
        y.backward(gy)
        _close(lin.weight.grad - (0.5 if preset else 0.0), gw_ref, rtol=3e-5, what=f"weight grad (preset={preset})")
        _close(lin.bias.grad, gb_ref, rtol=2e-5, what="bias grad")
        _close(xin.grad[:, :in_f], gy @ lin.weight.detach(), rtol=2e-5, what="input grad")


def test_trainer_checkpoint_resume_continues_the_trajectory(native_lib):
    """``Trainer.state_dict`` / ``load_state_dict`` (model, flat Adam, lr schedule, EXP3 weights + L1 norms, Philox
    step): restoring into the SAME trainer (its step graph already captured: everything is restored in place)
    and into a FRESH one reproduces the losses of the original run."""
    from bliss_gnn_b200.graph import synthetic_graph
    from bliss_gnn_b200.train import DataModule, Trainer, build_model
    dev = _dev()
    g = synthetic_graph("flickr", seed=0, scale=0.05).to(dev)

    def make():
        dm = DataModule("flickr", fan_out=[128, 64, 32], eta=0.1, device=dev, batch_size=32, sampler="poisson-bandit",
                        model="sage", seed=0, graph=g)
        torch.manual_seed(3)
        model = build_model("sage", dm.in_feats, 64, dm.n_classes, 3, dropout=0.0).to(dev)
        return dm, Trainer(dm, model, 0.002, static_graph=True, eager_warmup=3)

    dm, tr = make()
    batches = [b for _, b in zip(range(14), dm.train_batches())]
    for b in batches[:8]:
        tr.training_step(b)
    assert tr.graph_replays >= 3
    sd = tr.state_dict()
    ref = [float(tr.training_step(b).item()) for b in batches[8:]]
    tr.load_state_dict(sd)                                   # back in time, graph stays captured
    again = [float(tr.training_step(b).item()) for b in batches[8:]]
    for a, b in zip(ref, again):
        assert abs(a - b) <= 1e-6 * max(1.0, abs(a)), (ref, again)
    _, tr2 = make()
    tr2.load_state_dict(sd)                                  # fresh process: eager sizing steps first
    fresh = [float(tr2.training_step(b).item()) for b in batches[8:]]
    for a, b in zip(ref, fresh):
        assert abs(a - b) <= 2e-4 * max(1.0, abs(a)), (ref, fresh)
