"""GPU parity tests of the sampling path (C ABI kernels) against the CPU oracle.

Bar (north-star): sampled node sets, relabelling, block CSR and edge ids bit-exact given identical
uniform draws and bandit state; probabilities / weights within 1e-5 relative in fp32 (tolerance
written at each assert).  Against the oracle's ``accum='contract'`` mode (the device's numeric
contract) values are expected bit-exact up to 1 ulp in a few entries; against ``accum='native'``
(torch-order sums) the tie band around P is excluded by the draw generator (tests/util.SafeDraws).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import samplers as osamp
from oracle import philox
from tests.util import SafeDraws, assert_blocks_equal, philox_uniform_fn, random_graph

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "toy_kat.json")
RTOL = 1e-5   # north-star tolerance for fp32 values


def _dev():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch.device("cuda:0")


def _device_sampler(cls_name, fanouts, **kw):
    from bliss_gnn_b200 import sampler as S
    return getattr(S, cls_name)(fanouts, **kw)


def test_philox_matches_oracle(native_lib):
    from bliss_gnn_b200 import _native as N
    dev = _dev()
    nids = torch.arange(0, 5000, 7, dtype=torch.int32, device=dev)
    out = torch.empty(nids.numel(), dtype=torch.float32, device=dev)
    seed, step, layer = 0x1234_5678_9ABC_DEF0, (1 << 33) + 5, 2
    N.check(native_lib.bliss_philox_fill(seed, step, layer, N.ptr(nids), nids.numel(), N.ptr(out), N.stream()), "philox")
    ref = philox.uniform_for_nodes(seed, step, layer, nids.cpu().numpy())
    assert np.array_equal(out.cpu().numpy(), ref)          # bit-exact
    assert ref.min() >= 0.0 and ref.max() < 1.0


def test_toy_known_answer(native_lib):
    """The hand-derived toy vector (tests/golden/toy_kat.json) through the device path."""
    from bliss_gnn_b200.graph import toy_graph, normalized_edata
    kat = json.load(open(GOLDEN))
    g = toy_graph()
    g.edata["w"] = normalized_edata(g)
    g = g.to(_dev())
    s = _device_sampler("PoissonBanditLadiesSampler", [kat["fanout"]], eta=kat["eta"])
    s.inject_uniforms = torch.tensor(kat["u_inject"], dtype=torch.float32, device=g.device)
    inp, outn, blocks = s.sample_blocks(g, torch.tensor(kat["seeds"]))
    b = blocks[0]
    kb = kat["block"]
    assert b.srcdata["_ID"].tolist() == kb["src_nid"] and inp.tolist() == kb["src_nid"]
    assert b.dstdata["_ID"].tolist() == kb["dst_nid"] and outn.tolist() == kb["dst_nid"]
    assert b.edge_src.tolist() == kb["edge_src_local"]      # insg-filtered order == native order
    assert b.edge_dst.tolist() == kb["edge_dst_local"]
    assert b.edata["_ID"].tolist() == kb["eid"]
    ctr = s.last_counters[0]
    assert abs(ctr.c - kat["c"]) <= 1e-6 * kat["c"] and ctr.iters == kat["iters"]
    np.testing.assert_allclose(b.edata["q_ij"].cpu().numpy(), kb["q_ij"], rtol=RTOL)
    np.testing.assert_allclose(b.edata["edge_weights"].cpu().numpy(), kb["edge_weights"], rtol=RTOL)
    np.testing.assert_allclose(b.srcdata["node_prob"].cpu().numpy(), kb["node_prob"], rtol=RTOL)
    np.testing.assert_allclose(b.edata["w"].cpu().numpy(), [kat["w_static"][e] for e in kb["eid"]], rtol=1e-6)
    b.srcdata["embed_norm"] = torch.ones(b.num_src_nodes(), device=g.device)
    s.calculate_rewards(0, b, g, s.calculate_alpha(b))
    np.testing.assert_allclose(b.edata["rewards"].cpu().numpy(), kat["rewards"], rtol=RTOL)
    s.exp3(blocks, g)
    np.testing.assert_allclose(s.exp3_weights[0].cpu().numpy(), kat["exp3_after"], rtol=RTOL)


GRAPHS = {
    # name: (nodes, edges, hubs, hub_degree, batch, fanouts)
    "light": (400, 1500, 0, 0, 16, [64, 32, 16]),                 # every row handled by a warp
    "heavy": (3000, 20000, 6, 1500, 48, [512, 256, 128]),         # CTA rows staged in shared memory
    "huge_row": (12000, 30000, 2, 9500, 32, [1024, 512, 64]),     # rows beyond the 8192-float stage
}


def _oracle_and_device(gname, cls_name, accum, step=3, seed=11, eta=0.1, draws="philox", ora_kw=None, **kw):
    V, E, hubs, hdeg, batch, fan = GRAPHS[gname]
    g = random_graph(V, E, seed=5, hubs=hubs, hub_degree=hdeg)
    seeds = torch.randperm(V, generator=torch.Generator().manual_seed(1))[:batch]
    if hubs:
        seeds[:hubs] = torch.arange(hubs)       # make sure the hub rows are in the first frontier
    is_bandit = "Bandit" in cls_name
    okw = dict(eta=eta) if is_bandit else {}
    safe = SafeDraws(V, seed, step) if draws == "safe" else None
    ora = getattr(osamp, cls_name)(fan, accum=accum, uniform_fn=safe or philox_uniform_fn(seed, step), **okw,
                                   **(ora_kw or {}), **kw)
    if ora_kw and "dtype" in ora_kw:
        g.edata["w"] = g.edata["w"].to(ora_kw["dtype"])
    o_in, o_out, o_blocks = ora.sample_blocks(g, seeds)
    g.edata["w"] = g.edata["w"].float()
    gd = g.to(_dev())
    dev = _device_sampler(cls_name, fan, rng_seed=seed, **okw, **kw)
    dev.step = step
    if safe is not None:
        dev.inject_uniforms = {l: u.to(gd.device) for l, u in safe.per_layer.items()}
    d_in, d_out, d_blocks = dev.sample_blocks(gd, seeds)
    return g, gd, ora, dev, (o_in, o_out, o_blocks), (d_in, d_out, d_blocks)


@pytest.mark.parametrize("gname", list(GRAPHS))
def test_poisson_bandit_contract_parity(native_lib, gname):
    """Device vs oracle under the same numeric contract: structure bit-exact, values ~1 ulp."""
    g, gd, ora, dev, (o_in, _, o_blocks), (d_in, d_out, d_blocks) = _oracle_and_device(
        gname, "PoissonBanditLadiesSampler", "contract")
    assert torch.equal(d_in.cpu().long(), o_in)
    for l, (db, ob) in enumerate(zip(d_blocks, o_blocks)):
        st = assert_blocks_equal(db, ob, rtol=1e-6)
        ctr = dev.last_counters[l]
        assert ctr.n_cand == ora.trace["prob"][l][0].numel()
        if l in ora.trace.get("c", {}):
            c, it = ora.trace["c"][l]
            # fp64 row sums are order-dependent in their last bits: a few p differ by 1 ulp -> c by ~1e-9
            assert ctr.iters == it and abs(ctr.c - c) <= 1e-7 * c
        else:
            assert ctr.take_all == 1
        # row sums of the normalised block weights equal the kept in-degree (bandit_sampler.py:316-320)
        rs = torch.zeros(db.num_dst_nodes(), dtype=torch.float64, device=gd.device).index_add_(
            0, db.edge_dst.long(), db.edata["edge_weights"].double())
        torch.testing.assert_close(rs, db.in_degrees().double(), rtol=1e-5, atol=0)


@pytest.mark.parametrize("gname,dtype,rtol", [("light", torch.float32, RTOL), ("heavy", torch.float64, RTOL),
                                              ("huge_row", torch.float64, RTOL)])
def test_poisson_bandit_native_parity_safe_draws(native_lib, gname, dtype, rtol):
    """Device vs the torch-order oracle: sets bit-exact once ties are excluded by the draw generator.
    Values: within 1e-5 of the float64 oracle (the exact-arithmetic reference) on the graphs with 1,500- and
    9,500-edge rows; the float32 torch-order oracle is used on the light graph only (its own sequential fp32
    ``index_add_`` is good to ~2e-5 on long rows, i.e. less accurate than the device)."""
    *_, (o_in, _, o_blocks), (d_in, _, d_blocks) = _oracle_and_device(
        gname, "PoissonBanditLadiesSampler", "native", draws="safe", ora_kw=dict(dtype=dtype))
    assert torch.equal(d_in.cpu().long(), o_in)
    for db, ob in zip(d_blocks, o_blocks):
        assert_blocks_equal(db, ob, rtol=rtol)


@pytest.mark.parametrize("cls_name", ["PoissonLadiesSampler", "LadiesSampler", "BanditLadiesSampler"])
@pytest.mark.parametrize("gname", ["light", "heavy"])
def test_other_samplers_parity(native_lib, cls_name, gname):
    *_, (o_in, _, o_blocks), (d_in, _, d_blocks) = _oracle_and_device(gname, cls_name, "contract")
    assert torch.equal(d_in.cpu().long(), o_in)
    for db, ob in zip(d_blocks, o_blocks):
        assert_blocks_equal(db, ob, rtol=RTOL, check=("edge_weights", "q_ij"))
        if "Bandit" not in cls_name:
            assert "q_ij" not in dict.keys(db.edata)        # ladies_sampler.py:99-106 attaches neither
            assert "node_prob" not in dict.keys(db.srcdata)


def test_bitmap_candidate_collection_matches_dense(native_lib):
    """Both candidate-collection modes (dense accumulator scan / RED.OR bitmap) yield the same blocks."""
    g = random_graph(3000, 20000, seed=5, hubs=4, hub_degree=1200).to(_dev())
    seeds = torch.arange(0, 64)
    out = {}
    for mode in ("dense", "bitmap"):
        dev = _device_sampler("PoissonBanditLadiesSampler", [256, 128, 64], eta=0.1, rng_seed=3)
        dev.collect = mode
        _, _, out[mode] = dev.sample_blocks(g, seeds)
        assert int(dev._wsp.cand_bits.count_nonzero()) == 0 and int(dev._wsp.acc.count_nonzero()) == 0
    for x, y in zip(out["dense"], out["bitmap"]):
        assert torch.equal(x.srcdata["_ID"], y.srcdata["_ID"]) and torch.equal(x.edge_src, y.edge_src)
        assert torch.equal(x.edata["edge_weights"], y.edata["edge_weights"])


@pytest.mark.parametrize("cls_name", ["PoissonBanditLadiesSampler", "BanditLadiesSampler", "PoissonLadiesSampler"])
def test_overridden_stage_method_takes_the_stage_path(native_lib, cls_name):
    """The reference's stage methods stay extension points: a subclass override switches from the fused
    two-call fast path to the per-stage entry points, with bit-identical blocks."""
    from bliss_gnn_b200 import sampler as S
    g = random_graph(3000, 20000, seed=5, hubs=4, hub_degree=1200).to(_dev())
    seeds = torch.arange(0, 64)
    base = getattr(S, cls_name)
    calls = []

    class Sub(base):
        def select_neighbors(self, prob, num):
            calls.append(num)
            return super().select_neighbors(prob, num)

    kw = dict(eta=0.1) if "Bandit" in cls_name else {}
    a, b = base([256, 128, 64], rng_seed=3, **kw), Sub([256, 128, 64], rng_seed=3, **kw)
    assert a._stages_not_overridden() and not b._stages_not_overridden()
    _, _, ba = a.sample_blocks(g, seeds)
    _, _, bb = b.sample_blocks(g, seeds)
    assert calls == [64, 128, 256]
    for x, y in zip(ba, bb):
        assert torch.equal(x.srcdata["_ID"], y.srcdata["_ID"]) and torch.equal(x.edge_src, y.edge_src)
        assert torch.equal(x.edata["edge_weights"], y.edata["edge_weights"]) and torch.equal(x.indptr, y.indptr)


def test_take_all_branch(native_lib):
    """N_c <= fanout: P = 1 for every candidate, block = full in-neighbourhood (bandit_sampler.py:392-393)."""
    g = random_graph(300, 900, seed=2)
    seeds = torch.arange(0, 40)
    dev = _device_sampler("PoissonBanditLadiesSampler", [100000], eta=0.1)
    gd = g.to(_dev())
    _, _, blocks = dev.sample_blocks(gd, seeds)
    b = blocks[0]
    assert dev.last_counters[0].take_all == 1
    assert b.num_edges() == int(g.in_degrees(seeds).sum())
    assert torch.all(b.srcdata["node_prob"] == 1)
    ora = osamp.PoissonBanditLadiesSampler([100000], eta=0.1, accum="contract", uniform_fn=philox_uniform_fn(0, 0))
    _, _, ob = ora.sample_blocks(g, seeds)
    assert_blocks_equal(b, ob[0], rtol=1e-6)


def test_importance_sampling_off(native_lib):
    g = random_graph(500, 3000, seed=8)
    seeds = torch.arange(10, 40)
    kw = dict(eta=0.2, importance_sampling=False)
    ora = osamp.PoissonBanditLadiesSampler([64, 32], accum="contract", uniform_fn=philox_uniform_fn(4, 0), **kw)
    _, _, ob = ora.sample_blocks(g, seeds)
    dev = _device_sampler("PoissonBanditLadiesSampler", [64, 32], rng_seed=4, **kw)
    _, _, db = dev.sample_blocks(g.to(_dev()), seeds)
    for a, b in zip(db, ob):
        assert_blocks_equal(a, b, rtol=RTOL)


def test_workspace_invariant_and_determinism(native_lib):
    g = random_graph(3000, 20000, seed=5, hubs=4, hub_degree=1200).to(_dev())
    seeds = torch.arange(0, 64)
    dev = _device_sampler("PoissonBanditLadiesSampler", [256, 128, 64], eta=0.1, rng_seed=3)
    _, _, b1 = dev.sample_blocks(g, seeds)
    w = dev._wsp
    assert int(w.acc.count_nonzero()) == 0 and int((w.first_pos != -1).sum()) == 0
    assert int((w.node_info[0::2] != -1).sum()) == 0 and int(w.sel_bits.count_nonzero()) == 0
    assert int(w.cand_bits.count_nonzero()) == 0
    dev.step = 0                       # same Philox counters -> identical blocks, bit for bit
    _, _, b2 = dev.sample_blocks(g, seeds)
    for x, y in zip(b1, b2):
        assert torch.equal(x.edge_src, y.edge_src) and torch.equal(x.indptr, y.indptr)
        assert torch.equal(x.edata["edge_weights"], y.edata["edge_weights"])
        assert torch.equal(x.srcdata["node_prob"], y.srcdata["node_prob"])
    dev.step = 1
    _, _, b3 = dev.sample_blocks(g, seeds)
    assert not torch.equal(b1[0].srcdata["_ID"], b3[0].srcdata["_ID"]) or b1[0].num_src_nodes() != b3[0].num_src_nodes()


@pytest.mark.parametrize("normalize", ["lazy", "literal"])
def test_exp3_update_parity_three_steps(native_lib, normalize):
    """sample → (fake forward: embed_norm) → exp3, three times; bandit weights within 1e-5 of the oracle."""
    V, E, hubs, hdeg, batch, fan = GRAPHS["heavy"]
    g = random_graph(V, E, seed=5, hubs=hubs, hub_degree=hdeg)
    gd = g.to(_dev())
    seed = 21
    ora = osamp.PoissonBanditLadiesSampler(fan, eta=0.1, accum="contract")
    dev = _device_sampler("PoissonBanditLadiesSampler", fan, eta=0.1, rng_seed=seed, normalize=normalize)
    gen = torch.Generator().manual_seed(0)
    for step in range(3):
        seeds = torch.randperm(V, generator=gen)[:batch]
        ora.uniform_fn = philox_uniform_fn(seed, step)
        _, _, ob = ora.sample_blocks(g, seeds)
        _, _, db = dev.sample_blocks(gd, seeds)
        for a, b in zip(db, ob):
            assert_blocks_equal(a, b, rtol=RTOL)
            emb = torch.rand(b.num_src_nodes(), generator=gen) * 3 + 0.1
            b.srcdata["embed_norm"] = emb
            a.srcdata["embed_norm"] = emb.to(gd.device)
        ora.exp3(ob, g)
        dev.exp3(db, gd)
        w_dev = dev.exp3_weights.cpu().double()
        w_ora = ora.exp3_weights.double()
        rel = ((w_dev - w_ora).abs() / w_ora).max().item()
        assert rel <= RTOL, f"step {step}: exp3 weights max rel err {rel}"
        torch.testing.assert_close(w_dev.sum(dim=1), torch.ones(3, dtype=torch.float64), rtol=1e-6, atol=0)


def test_lazy_renorm_is_driven_by_the_weight_maximum(native_lib):
    """Lazy normalisation: the update kernels keep the running maximum of a layer's weights; the physical
    re-normalisation (a pass over L x |E| weights) happens when that maximum nears the top of the fp32 range, not on
    a schedule — and leaves the normalised weights (``exp3_weights``) unchanged."""
    V, E, hubs, hdeg, batch, fan = GRAPHS["heavy"]
    gd = random_graph(V, E, seed=5, hubs=hubs, hub_degree=hdeg).to(_dev())
    dev = _device_sampler("PoissonBanditLadiesSampler", fan, eta=0.1, rng_seed=3, normalize="lazy")
    gen = torch.Generator().manual_seed(0)

    def step():
        _, _, db = dev.sample_blocks(gd, torch.randperm(V, generator=gen)[:batch])
        for a in db:
            a.srcdata["embed_norm"] = (torch.rand(a.num_src_nodes(), generator=gen) * 3 + 0.1).to(gd.device)
        dev.exp3(db, gd)

    step()
    L = len(fan)
    true_max = torch.stack([w.max() for w in dev._w_csc])
    assert torch.all(dev._wmax >= true_max) and torch.all(dev._wmax <= true_max * 1.000001), (dev._wmax, true_max)
    # the schedule's horizon passes with small weights: no re-normalisation (L1 stays ~|E|)
    dev._updates_since_renorm = dev.renorm_every
    dev.tick_renorm(L)
    assert float(dev._l1.min()) > 0.5 * gd.num_edges() and dev._updates_since_renorm == 0
    before = dev.exp3_weights.clone()
    # a weight near the top of the range: re-normalised at the next look (device value, then the pinned copy)
    for mirrored in (False, True):
        dev._wmax.fill_(1e37)
        dev._wmax_host.fill_(1e37 if mirrored else 1.0)
        dev._updates_since_renorm = dev.renorm_every
        dev.tick_renorm(L, mirrored=mirrored)
        torch.testing.assert_close(dev._l1, torch.ones(L, dtype=torch.float64, device=gd.device))
        assert float(dev._wmax.max()) == 1.0 and float(dev._wmax_host.max()) == 1.0
        torch.testing.assert_close(dev.exp3_weights, before, rtol=1e-6, atol=0)
        step()                                    # updates after the re-scale keep the bound
        true_max = torch.stack([w.max() for w in dev._w_csc])
        assert torch.all(dev._wmax >= true_max)
        before = dev.exp3_weights.clone()


def test_gat_alpha_rewards(native_lib):
    """GAT alpha path of the bandit (bandit_sampler.py:146-154) with random a_ij."""
    g = random_graph(800, 5000, seed=9)
    gd = g.to(_dev())
    seeds = torch.arange(0, 32)
    ora = osamp.PoissonBanditLadiesSampler([128, 64], eta=0.1, model="gat", accum="contract",
                                           uniform_fn=philox_uniform_fn(2, 0))
    dev = _device_sampler("PoissonBanditLadiesSampler", [128, 64], eta=0.1, model="gat", rng_seed=2)
    _, _, ob = ora.sample_blocks(g, seeds)
    _, _, db = dev.sample_blocks(gd, seeds)
    gen = torch.Generator().manual_seed(3)
    for a, b in zip(db, ob):
        assert_blocks_equal(a, b, rtol=RTOL)
        emb = torch.rand(b.num_src_nodes(), generator=gen) + 0.5
        att = torch.randn(b.num_edges(), generator=gen)
        b.srcdata["embed_norm"], b.edata["a_ij"] = emb, att
        # the device block's native edge order equals the oracle's (insg order filtered)
        assert torch.equal(a.edge_src.cpu().long(), b.src)
        a.srcdata["embed_norm"], a.edata["a_ij"] = emb.to(gd.device), att.to(gd.device)
    for l, (a, b) in enumerate(zip(db, ob)):
        al = ora.calculate_alpha(b)
        ora.calculate_rewards(l, b, g, al)
        dev.calculate_rewards(l, a, gd, dev.calculate_alpha(a))
        r_d, r_o = a.edata["rewards"].cpu().double(), b.edata["rewards"].double()
        # alpha divides by a sum of SIGNED logits (bandit_sampler.py:148-154): fp32 rounding of the sum is amplified by
        # the row's condition number kappa = sum|a| / |sum a| in alpha and 2 kappa in the reward (alpha^2), so the
        # 1e-5 bar is scaled by it; rows whose sum does not cancel (kappa ~ 1) are held to 1e-5 itself
        a64 = b.edata["a_ij"].double()
        s = torch.zeros(b.num_dst_nodes(), dtype=torch.float64).index_add_(0, b.dst, a64)
        sa = torch.zeros(b.num_dst_nodes(), dtype=torch.float64).index_add_(0, b.dst, a64.abs())
        kappa = (sa / s.abs().clamp(min=1e-300))[b.dst].clamp(min=1.0)
        assert ((r_d - r_o).abs() <= RTOL * 2.0 * kappa * r_o.abs() + 1e-30).all()
    ora.exp3(ob, g)
    dev.exp3(db, gd)
    w_dev, w_ora = dev.exp3_weights.cpu().double(), ora.exp3_weights.double()
    # the weight moves by exp(x), x = min(1, reward-term) (:240-246): the reward error enters damped by x <= 1
    assert ((w_dev - w_ora).abs() / w_ora).max().item() <= RTOL


def test_packed_exchange_apply_matches_sequential_updates(native_lib):
    """The data-parallel apply kernel on a hand-built 2-rank exchange buffer (no NCCL needed): every
    rank's (position, exponent) list multiplies into the weights, shared positions included."""
    from bliss_gnn_b200 import _native as N
    from bliss_gnn_b200.parallel import BanditExchange
    dev = _dev()
    E, caps, world = 5000, [700, 300], 2
    ex = BanditExchange(caps, world, dev, group=None)
    gen = torch.Generator().manual_seed(0)
    w = [torch.rand(E, generator=gen) + 0.5 for _ in caps]
    w_dev = [x.to(dev).clone() for x in w]
    l1 = torch.zeros(len(caps), dtype=torch.float64, device=dev)
    expect = [x.double().clone() for x in w]
    first_pos = {}
    for r in range(world):
        base = ex.recv[r * ex.stride:(r + 1) * ex.stride]
        for l, cap in enumerate(caps):
            n = cap - 50 * (r + 1)
            pos = torch.randperm(E, generator=gen)
            if r == 1:                                # make sure both ranks hit some common positions
                pos = torch.cat([first_pos[l][:20], pos[~torch.isin(pos, first_pos[l][:20])]])
            pos = pos[:n]                             # unique within a rank (a block has no duplicate edges)
            first_pos.setdefault(l, pos)
            xs = torch.rand(n, generator=gen) * 0.01
            base[:64].view(torch.int64)[l] = n
            base[ex.pos_off[l]:ex.pos_off[l] + 4 * n].view(torch.int32).copy_(pos.to(torch.int32))
            base[ex.x_off[l]:ex.x_off[l] + 4 * n].view(torch.float32).copy_(xs)
            expect[l][pos] = expect[l][pos] * torch.exp(xs.double())
    wmax = torch.zeros(len(caps), dtype=torch.float32, device=w_dev[0].device)
    for l in range(len(caps)):
        N.call("bliss_apply_updates_packed", N.ptr(ex.recv), ex.stride, world, 8 * l, ex.pos_off[l], ex.x_off[l],
               caps[l], N.ptr(w_dev[l]), N.ptr(l1[l:l + 1]), N.ptr(wmax[l:l + 1]), N.stream())
        torch.testing.assert_close(w_dev[l].cpu().double(), expect[l], rtol=1e-6, atol=0)
        # the running maximum of the UPDATED weights (the range guard of the lazy normalisation)
        assert 0.5 < float(wmax[l]) <= float(w_dev[l].max())
        delta = float(l1[l].item())
        assert abs(delta - float((expect[l] - w[l].double()).sum())) <= 1e-5 * abs(delta)


@pytest.mark.parametrize("fanouts", [[10, 5, 3], [300, 40, -1], [-1, -1]])
def test_neighbor_and_full_samplers_match_oracle(native_lib, fanouts):
    """``--sampler neighbor | full`` (train_lightning.py:349-357): per-seed uniform k-subsets of the in-edges (all of them
    for -1) and ``to_block`` relabelling — block structure bit-exact against the oracle restatement under the same
    per-edge Philox keys; rows of 1,500 in-edges exercise the radix select, fan-outs above the degree the take-all
    branch; the blocks carry no edge weights."""
    from bliss_gnn_b200.sampler import MultiLayerFullNeighborSampler, NeighborSampler
    V, E, hubs, hdeg, batch, _ = GRAPHS["heavy"]
    g = random_graph(V, E, seed=5, hubs=hubs, hub_degree=hdeg)
    seeds = torch.randperm(V, generator=torch.Generator().manual_seed(1))[:batch]
    seeds[:hubs] = torch.arange(hubs)
    gd = g.to(_dev())
    seed = 13
    full = all(f < 0 for f in fanouts)
    dev = MultiLayerFullNeighborSampler(len(fanouts), rng_seed=seed) if full else NeighborSampler(fanouts, rng_seed=seed)
    for step in range(2):
        ora = osamp.NeighborSampler(fanouts, key_fn=lambda l, pos, st=step: torch.from_numpy(
            philox.edge_keys(seed, st, l, pos.numpy()).astype(np.int64)))
        o_in, _, ob = ora.sample_blocks(g, seeds)
        d_in, d_out, db = dev.sample_blocks(gd, seeds)
        assert torch.equal(d_in.cpu().long(), o_in) and torch.equal(d_out.cpu().long(), seeds)
        for l, (a, b) in enumerate(zip(db, ob)):
            assert_blocks_equal(a, b, check=())
            assert "edge_weights" not in dict.keys(a.edata)
            deg = g.in_degrees(b.dstdata["_ID"])
            want = deg if fanouts[l] < 0 else deg.clamp(max=fanouts[l])
            assert torch.equal(a.in_degrees().cpu().long(), want), "every seed keeps min(fanout, in-degree) in-edges"
    w = dev._wsp      # workspace invariant restored (nothing |V|-sized is cleared per step)
    assert int(w.acc.count_nonzero()) == 0 and int((w.node_info[0::2] != -1).sum()) == 0
    assert int(w.sel_bits.count_nonzero()) == 0


def test_neighbor_sampler_is_uniform(native_lib):
    """Every in-edge of a 1,500-edge row is kept with probability fanout / degree: 600 steps x fan-out 50, per-edge
    counts against the binomial expectation (5 sigma) and a chi-square of the whole row."""
    from bliss_gnn_b200.sampler import NeighborSampler
    V, E, hubs, hdeg, _, _ = GRAPHS["heavy"]
    g = random_graph(V, E, seed=5, hubs=hubs, hub_degree=hdeg).to(_dev())
    dev = NeighborSampler([50], rng_seed=3)
    seeds = torch.tensor([0])
    d = int(g.in_degrees(seeds.to(g.device))[0])
    counts = torch.zeros(g.num_edges(), device=g.device)
    steps = 600
    for _ in range(steps):
        _, _, (b,) = dev.sample_blocks(g, seeds)
        assert b.num_edges() == 50
        counts[b.csc_pos.long()] += 1
    a = int(g.indptr[0])
    c = counts[a:a + d].cpu().double()
    p = 50.0 / d
    mean, sd = steps * p, (steps * p * (1 - p)) ** 0.5
    assert float((c - mean).abs().max()) < 5.5 * sd, (float(c.min()), float(c.max()), mean, sd)
    chi2 = float(((c - mean) ** 2 / (steps * p * (1 - p))).sum())
    assert abs(chi2 - d) < 6 * (2 * d) ** 0.5, (chi2, d)
