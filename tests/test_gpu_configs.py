"""Device vs CPU oracle on every configuration BASELINE.json names, at the configuration's full graph size,
batch and fan-out (the Reddit shape lives in tests/test_gpu_fullsize.py next to its graph fixture):

* Pubmed shape, GCN, poisson-bandit vs ladies          (configs[1])
* Flickr shape, GATv2 — the GAT alpha branch of the bandit (configs[2])
* Yelp shape, SAGE, 100 multi-label classes, BCE loss    (configs[4])
* Cora shape, SAGE: 12-step trajectory of the whole-step CUDA graph against the oracle loop (configs[0])

One full step per configuration, in the reference's order (``train_lightning.py:100-168,463-471``): sampled blocks
(structure bit-exact, values 1e-5) → model forward (logits, embed_norm, a_ij) → loss → backward (every parameter
gradient) → ``exp3`` (EXP3 weights) → the NEXT step's blocks drawn from the updated weights.  Floating-point
values are held to 1e-5 against the oracle evaluated in float64 — the exact-arithmetic reference, so the error
measured is the device's own fp32 rounding (north-star: "within 1e-5 relative in fp32").  Measured errors are
written to gpurun_out/parity_stats.json.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import model as omodel
from oracle import samplers as osamp
from tests.util import (OracleLoop, SafeDraws, assert_blocks_equal, blocks_as, blocks_clone, close, close_grad,
                        copy_params, philox_uniform_fn, record, rel_to_max)

pytestmark = pytest.mark.gpu
RTOL = 1e-5      # north-star tolerance


def _dev():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch.device("cuda:0")


SAMPLERS = {"poisson-bandit": "PoissonBanditLadiesSampler", "bandit": "BanditLadiesSampler",
            "ladies": "LadiesSampler", "poisson-ladies": "PoissonLadiesSampler"}


def _graph(shape):
    """The synthetic graph of a dataset shape: generated on the GPU (seconds), mirrored to the CPU for the oracle."""
    from bliss_gnn_b200.graph import normalized_edata, synthetic_graph
    gd = synthetic_graph(shape, seed=0, device=_dev())
    gd.edata["w"] = normalized_edata(gd)
    return gd.to("cpu"), gd


def _models(kind, in_f, hidden, n_classes, dev):
    from bliss_gnn_b200 import model as M
    torch.manual_seed(3)
    if kind == "gat":
        args = (3, in_f, hidden, n_classes, [4, 4, 1], F.elu, 0.0, 0.0, 0.2, False)     # train_lightning.py:247,581-596
        dm, om = M.GATv2(*args).to(dev), omodel.GATv2(*args)
    else:
        cls_d, cls_o = (M.SAGE, omodel.SAGE) if kind == "sage" else (M.GCN, omodel.GCN)
        dm, om = cls_d(in_f, hidden, n_classes, 3, F.relu, 0.0).to(dev), cls_o(in_f, hidden, n_classes, 3, F.relu, 0.0)
    import copy
    om32 = copy.deepcopy(om)
    copy_params(om32, dm)                                  # the same oracle in float32: torch's own fp32 arithmetic
    copy_params(om, dm, torch.float64)
    return dm, om.double(), om32


def _gat_condition(ob):
    """Condition number of the GAT alpha (bandit_sampler.py:148-154) alpha_e = a_e / Σ_row a · Σ_row q with respect to
    the fp32 rounding of a_ij (head mean of SIGNED pre-softmax logits, accurate to ~1e-7 of max|a|, as asserted on
    a_ij itself): the numerator contributes max|a| / |a_e| (a logit that cancels to ~0 has no relative accuracy
    left), the row sum contributes Σ|a| / |Σ a|.  The reward is ∝ alpha², so its relative error is twice that."""
    a = ob.edata["a_ij"].double()
    n = ob.num_dst_nodes()
    s = torch.zeros(n, dtype=torch.float64).index_add_(0, ob.dst, a)
    sa = torch.zeros(n, dtype=torch.float64).index_add_(0, ob.dst, a.abs())
    return (sa / s.abs().clamp(min=1e-300))[ob.dst] + a.abs().max() / a.abs().clamp(min=1e-300)


@pytest.mark.parametrize("shape,kind,sampler,batch,fan,hidden", [
    ("pubmed", "gcn", "poisson-bandit", 32, [512, 256, 128], 256),
    ("pubmed", "gcn", "ladies", 32, [512, 256, 128], 256),
    ("pubmed", "gcn", "poisson-ladies", 32, [512, 256, 128], 256),
    ("flickr", "gat", "poisson-bandit", 256, [4096, 2048, 1024], 256),
    ("yelp", "sage", "poisson-bandit", 256, [4096, 2048, 1024], 256),
    ("cora", "sage", "poisson-bandit", 32, [512, 256, 128], 256),
])
def test_config_full_step_matches_oracle(native_lib, shape, kind, sampler, batch, fan, hidden):
    from bliss_gnn_b200 import sampler as S
    tag = f"{shape}-{kind}-{sampler}"
    g, gd = _graph(shape)
    dev, V = gd.device, g.num_nodes()
    cls = SAMPLERS[sampler]
    bandit, poisson = "Bandit" in cls, "Poisson" in cls
    train = torch.nonzero(g.ndata["train_mask"], as_tuple=True)[0]
    perm = train[torch.randperm(train.numel(), generator=torch.Generator().manual_seed(1))]
    seed = 11
    okw = dict(eta=0.1, model=kind) if bandit else {}
    # Poisson samplers: float64 torch-order oracle (exact-arithmetic reference) with the tie band |u - P| <= 1e-4 P
    # excluded by the draw generator; top-k samplers: the device's numeric contract (keys must order identically)
    ora = getattr(osamp, cls)(fan, accum="native" if poisson else "contract",
                              dtype=torch.float64 if poisson else torch.float32, **okw)
    if poisson:
        g.edata["w"] = osamp.normalized_edata(g, torch.float64)
    dsm = getattr(S, cls)(fan, rng_seed=seed, **okw)
    dmodel, omod, omod32 = _models(kind, g.ndata["features"].shape[1], hidden, g.n_classes, dev)
    multilabel = bool(g.multilabel)

    for step in range(2):
        seeds = perm[step * batch:(step + 1) * batch]
        draws = SafeDraws(V, seed, step) if poisson else philox_uniform_fn(seed, step)
        ora.uniform_fn = draws
        o_in, _, ob = ora.sample_blocks(g, seeds)
        dsm.step = step
        if poisson:
            dsm.inject_uniforms = {l: u.to(dev) for l, u in draws.per_layer.items()}
        d_in, _, db = dsm.sample_blocks(gd, seeds)
        assert torch.equal(d_in.cpu().long(), o_in), f"{tag}: input nodes of step {step} differ"
        # step 1 is drawn from the UPDATED bandit weights: structure stays bit-exact; for GAT the values inherit the
        # conditioned error of the reference's alpha (a sum of signed logits in the denominator, see below): the few
        # ill-conditioned weights are ~1e-4 apart, and so are the q_ij / W~ of the edges that use them
        rtol_b = RTOL if (step == 0 or kind != "gat") else 1e-3
        for l, (a, b) in enumerate(zip(db, ob)):
            st = assert_blocks_equal(a, b, rtol=rtol_b)
            for k, v in st.items():
                record(f"{tag}/step{step}/block{l}/{k}", v)
            record(f"{tag}/step{step}/block{l}/sizes", [a.num_dst_nodes(), a.num_src_nodes(), a.num_edges()])
        if step == 1:
            break
        # ---- model forward / backward on the same blocks ----
        ob32 = blocks_clone(ob, torch.float32)
        blocks_as(ob, torch.float64)
        xd = gd.ndata["features"][d_in.long()]
        xo = g.ndata["features"][o_in].double()
        yd, yo = dmodel(db, xd), omod(ob, xo)
        yo32 = omod32(ob32, g.ndata["features"][o_in])
        close(yd, yo, RTOL, f"{tag}/logits")
        for l, (a, b) in enumerate(zip(db, ob)):
            close(a.srcdata["embed_norm"], b.srcdata["embed_norm"], RTOL, f"{tag}/embed_norm{l}")
            if kind == "gat":
                assert torch.equal(a.edge_src.cpu().long(), b.src)          # same native edge order
                close(a.edata["a_ij"], b.edata["a_ij"], RTOL, f"{tag}/a_ij{l}")
        labels = g.ndata["labels"][seeds.long()]
        if multilabel:                                                       # train_lightning.py:77-79
            ld = F.binary_cross_entropy_with_logits(yd, labels.to(dev))
            lo = F.binary_cross_entropy_with_logits(yo, labels.double())
            lo32 = F.binary_cross_entropy_with_logits(yo32, labels)
        else:
            ld, lo, lo32 = F.cross_entropy(yd, labels.to(dev)), F.cross_entropy(yo, labels), F.cross_entropy(yo32, labels)
        close(ld.reshape(1), lo.reshape(1), RTOL, f"{tag}/loss")
        ld.backward()
        lo.backward()
        lo32.backward()
        dgrads, g32 = dict(dmodel.named_parameters()), dict(omod32.named_parameters())
        for n, q in omod.named_parameters():
            close_grad(dgrads[n].grad, q.grad, g32[n].grad, RTOL, f"{tag}/grad/{n}")
        if not bandit:
            continue
        # ---- bandit update from this step's forward pass (bandit_sampler.py:251-267) ----
        for l, (a, b) in enumerate(zip(db, ob)):
            b.srcdata["embed_norm"] = b.srcdata["embed_norm"].detach()
            if kind == "gat":
                b.edata["a_ij"] = b.edata["a_ij"].detach()
            al = ora.calculate_alpha(b)
            ora.calculate_rewards(l, b, g, al)
            dsm.calculate_rewards(l, a, gd, dsm.calculate_alpha(a))
            r_d, r_o = a.edata["rewards"].cpu().double(), b.edata["rewards"].double()
            tol = RTOL * (2.0 * _gat_condition(b).clamp(min=1.0) if kind == "gat" else 1.0)
            err = ((r_d - r_o).abs() / (tol * r_o.abs()).clamp(min=1e-300))
            err = err[r_o.abs() > 1e-30]
            record(f"{tag}/rewards{l}/max_err_over_tol", float(err.max()) if err.numel() else 0.0)
            assert err.numel() == 0 or float(err.max()) <= 1.0, f"{tag}: rewards of layer {l}: {float(err.max())} x tolerance"
        conds = [_gat_condition(b) for b in ob] if kind == "gat" else None
        ora.exp3(ob, g)
        dsm.exp3(db, gd)
        w_dev, w_ora = dsm.exp3_weights.cpu().double(), ora.exp3_weights.double()
        rel_e = (w_dev - w_ora).abs() / w_ora
        tol = torch.full_like(rel_e, RTOL)
        if kind == "gat":
            # w *= exp(x), x = min(1, reward term) (bandit_sampler.py:240-248): the weight inherits the reward's relative
            # error scaled by x, and the GAT reward's relative error is 2 cond_e x 1e-5 (see _gat_condition) — the
            # reference's alpha divides by a sum of signed logits; rows where that sum cancels get the LARGEST updates
            for l, b in enumerate(ob):
                x = ora.trace["delta_reward"][l].double().abs()
                tol[l, b.edata["_ID"].long()] = RTOL * (1.0 + 2.0 * conds[l] * x)
        rel = (rel_e / tol).max().item()
        record(f"{tag}/exp3_weights/max_rel", rel_e.max().item())
        record(f"{tag}/exp3_weights/max_rel_over_tol", rel)
        assert rel <= 1.0, f"{tag}: EXP3 weights: {rel} x tolerance (max rel err {rel_e.max().item()})"
        torch.testing.assert_close(w_dev.sum(dim=1), torch.ones(len(fan), dtype=torch.float64), rtol=1e-6, atol=0)


def _force_oracle_state(loop, omod, tr, model):
    """Teacher forcing: the oracle continues from the DEVICE's parameters and Adam moments (cast to the oracle's
    dtype), so every step is compared on its own instead of through a free-running trajectory."""
    dt = next(omod.parameters()).dtype
    off = 0
    opt = tr.optimizer
    step = float(opt.step_dev.item())
    with torch.no_grad():
        for q, p in zip(omod.parameters(), tr.grads.params):
            n = p.numel()
            q.copy_(p.detach().cpu().to(dt))
            st = loop.opt.state[q]
            st["step"] = torch.tensor(step)
            st["exp_avg"] = opt.exp_avg[off:off + n].view_as(p).cpu().to(dt)
            st["exp_avg_sq"] = opt.exp_avg_sq[off:off + n].view_as(p).cpu().to(dt)
            off += n


def _trajectory(shape, n_steps, hidden, batch, fan, eager_warmup, tag, loss_rtol=1e-4, param_rtol=1e-4,
                teacher_forced=False):
    """Trainer(static_graph=True) (eager sizing steps, then the whole step as replayed CUDA graphs) against the
    oracle loop: same seed batches, same Philox stream, dropout 0, fp32 GEMMs (``--precision highest``).

    The oracle loop runs twice: with the model in float64 (the exact-arithmetic reference the device is held to)
    and in float32 (torch's own fp32 arithmetic).  Adam divides every gradient entry by its own running magnitude,
    so the few entries that are cancelling sums (relative error ~1 in ANY fp32 implementation) move by +-lr with a
    noise-determined sign: a max-norm bound on the parameters cannot hold for fp32, whoever computes it.  The
    device is therefore held to: losses within ``loss_rtol`` and parameters within ``param_rtol`` in relative L2 norm
    of the float64 trajectory — or within twice the float32 oracle's own distance from it, when that is larger
    (both recorded); sampled sizes identical at every step; EXP3 weights within 1e-5.

    ``teacher_forced``: after every step the oracles continue from the device's parameters and Adam moments, so each
    step (eager or replayed) is checked on its own.  Needed at the Reddit shape, where a free-running comparison is
    not meaningful in fp32: with ~10^6 hidden activations per step one relu input lands within rounding of zero
    every few steps, that gate differs between ANY two fp32 implementations (measured: 1 of 328,960 gates at step 0,
    scratch/diag_reddit4.py), the flipped entry perturbs the weight gradients by ~1e-4 of their maximum, and Adam's
    first steps (update = +-lr whatever the gradient's size) turn that into a different trajectory (losses 5e-4
    apart after two steps while every tensor of each single step agrees to 1e-6)."""
    from bliss_gnn_b200.train import DataModule, Trainer, build_model
    torch.set_float32_matmul_precision("highest")
    g, gd = _graph(shape)
    dm = DataModule(shape, fan_out=fan, eta=0.1, device=gd.device, batch_size=batch, sampler="poisson-bandit",
                    model="sage", seed=0, graph=gd)
    torch.manual_seed(3)
    model = build_model("sage", dm.in_feats, hidden, dm.n_classes, 3, dropout=0.0).to(gd.device)
    in_f = g.ndata["features"].shape[1]
    omod64, omod32 = (omodel.SAGE(in_f, hidden, g.n_classes, 3, F.relu, 0.0) for _ in range(2))
    copy_params(omod64, model, torch.float64)
    omod64 = omod64.double()
    copy_params(omod32, model)
    tr = Trainer(dm, model, 0.002, static_graph=True, eager_warmup=eager_warmup, pipeline=False)
    loop64, loop32 = (OracleLoop(g, m, "PoissonBanditLadiesSampler", fan, rng_seed=dm.sampler.rng_seed, eta=0.1, lr=0.002)
                      for m in (omod64, omod32))
    batches = []
    while len(batches) < n_steps:
        batches.extend(dm.train_batches())
    d_loss, o_loss, o32_loss = [], [], []
    for seeds in batches[:n_steps]:
        d_loss.append(float(tr.training_step(seeds).item()))
        o_loss.append(loop64.training_step(seeds))
        o32_loss.append(loop32.training_step(seeds))
        if teacher_forced:
            tr.flush()
            torch.cuda.synchronize()
            _force_oracle_state(loop64, omod64, tr, model)
            _force_oracle_state(loop32, omod32, tr, model)
        sizes_d = [(int(c.n_src), int(c.n_edges)) for c in dm.sampler.last_counters]
        for loop in (loop64, loop32):
            sizes_o = [(b.num_src_nodes(), b.num_edges()) for b in loop.last_blocks]
            assert sizes_d == sizes_o, f"{tag}: step {len(d_loss) - 1}: sampled sizes {sizes_d} vs oracle {sizes_o}"
    tr.flush()
    assert tr.graph_replays >= n_steps - eager_warmup - 1
    rel = lambda xs, ys: max(abs(a - b) / max(1.0, abs(b)) for a, b in zip(xs, ys))
    worst, floor = rel(d_loss, o_loss), rel(o32_loss, o_loss)
    record(f"{tag}/loss_max_rel", {"device_vs_fp64": worst, "torch_fp32_vs_fp64": floor})
    record(f"{tag}/losses", {"device": d_loss, "oracle_fp64": o_loss, "oracle_fp32": o32_loss})
    assert worst <= max(loss_rtol, 2.0 * floor), (d_loss, o_loss, o32_loss)
    dparams, p32 = dict(model.named_parameters()), dict(omod32.named_parameters())
    for n, q in ([] if teacher_forced else omod64.named_parameters()):
        ref = q.detach().double()
        l2 = lambda t: ((t.detach().cpu().double() - ref).norm() / ref.norm().clamp(min=1e-30)).item()
        e_dev, e_32 = l2(dparams[n]), l2(p32[n])
        record(f"{tag}/param/{n}", {"device_rel_l2": e_dev, "torch_fp32_rel_l2": e_32,
                                    "device_max_over_max": rel_to_max(dparams[n], ref),
                                    "torch_fp32_max_over_max": rel_to_max(p32[n], ref)})
        assert e_dev <= max(param_rtol, 2.0 * e_32), f"{tag}: {n}: rel L2 {e_dev:.3e} (torch fp32: {e_32:.3e})"
    w_dev, w_ora = dm.sampler.exp3_weights.cpu().double(), loop64.smp.exp3_weights.double()
    rel_w = ((w_dev - w_ora).abs() / w_ora).max().item()
    record(f"{tag}/exp3_weights/max_rel", rel_w)
    assert rel_w <= RTOL, f"{tag}: EXP3 weights after {n_steps} steps: max rel err {rel_w}"


def test_cora_shape_trajectory_matches_oracle_loop(native_lib):
    """configs[0]: 12 steps (3 eager + 9 graph replays) of sample → fwd → bwd → Adam → exp3 against the oracle."""
    _trajectory("cora", 12, 256, 32, [512, 256, 128], 3, "trajectory-cora")


def test_reddit_shape_trajectory_matches_oracle_loop(native_lib):
    """configs[3], the bench workload (232,965 nodes, ~115 M edges, batch 256, fan-out 4096/2048/1024, hidden 256):
    6 steps (2 eager + 4 replays of the whole-step CUDA graphs) against the oracle loop, teacher-forced (see
    ``_trajectory``): the sampled sizes of every layer identical at every step (the sets are drawn from bandit weights
    both sides updated), every step's loss within 1e-5 of the float64 oracle, EXP3 weights within 1e-5 after 6 updates."""
    _trajectory("reddit", 6, 256, 256, [4096, 2048, 1024], 2, "trajectory-reddit", loss_rtol=1e-5, teacher_forced=True)
