"""CPU tests pinning the oracle: the hand-derived toy known-answer vector, Philox4x32-10 Random123
vectors, and the algebraic invariants of SURVEY.md §8(c).  (The reference ships no tests and DGL is
not installable: parity is otherwise unpinned.)"""
import json
import os

import numpy as np
import pytest
import torch

from bliss_gnn_b200.graph import normalized_edata, toy_graph
from oracle import dglops, philox
from oracle import samplers as osamp
from tests.util import philox_uniform_fn, random_graph

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "toy_kat.json")


def test_philox_random123_known_answers():
    h = lambda r: [int(x) for x in r]
    assert h(philox.philox4x32_10((0, 0, 0, 0), (0, 0))) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert h(philox.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2)) == [
        0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert h(philox.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0))) == [
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize("dtype,accum,rtol", [(torch.float64, "native", 1e-12), (torch.float32, "native", 1e-6),
                                              (torch.float32, "contract", 1e-6)])
def test_oracle_reproduces_toy_known_answer(dtype, accum, rtol):
    kat = json.load(open(GOLDEN))
    g = toy_graph()
    g.edata["w"] = osamp.normalized_edata(g, dtype)
    u = torch.tensor(kat["u_inject"], dtype=torch.float32)
    s = osamp.PoissonBanditLadiesSampler([kat["fanout"]], eta=kat["eta"], dtype=dtype, accum=accum,
                                         uniform_fn=lambda l, nid, p=None: u[nid.long()])
    inp, out, blocks = s.sample_blocks(g, torch.tensor(kat["seeds"]))
    b, kb = blocks[0], kat["block"]
    assert b.srcdata["_ID"].tolist() == kb["src_nid"] == inp.tolist()
    assert b.dstdata["_ID"].tolist() == kb["dst_nid"]
    assert b.src.tolist() == kb["edge_src_local"] and b.dst.tolist() == kb["edge_dst_local"]
    assert b.edata["_ID"].tolist() == kb["eid"]
    c, it = s.trace["c"][0]
    assert it == kat["iters"] and abs(c - kat["c"]) <= max(rtol, 1e-7) * kat["c"]
    np.testing.assert_allclose(s.trace["prob"][0][1].double().numpy(), kat["P"], rtol=max(rtol, 1e-7))
    np.testing.assert_allclose(b.edata["q_ij"].double().numpy(), kb["q_ij"], rtol=rtol)
    np.testing.assert_allclose(b.edata["edge_weights"].double().numpy(), kb["edge_weights"], rtol=rtol)
    np.testing.assert_allclose(b.srcdata["node_prob"].double().numpy(), kb["node_prob"], rtol=rtol)
    b.srcdata["embed_norm"] = torch.ones(b.num_src_nodes(), dtype=dtype)
    s.exp3(blocks, g)
    np.testing.assert_allclose(b.edata["rewards"].double().numpy(), kat["rewards"], rtol=rtol)
    np.testing.assert_allclose(s.trace["delta_reward"][0].double().numpy(), kat["x"], rtol=rtol)
    np.testing.assert_allclose(s.exp3_weights[0].double().numpy(), kat["exp3_after"], rtol=rtol)


@pytest.mark.parametrize("cls", ["PoissonBanditLadiesSampler", "BanditLadiesSampler", "PoissonLadiesSampler",
                                 "LadiesSampler"])
def test_oracle_invariants_random_graph(cls):
    g = random_graph(600, 4000, seed=3, hubs=2, hub_degree=300)
    seeds = torch.arange(5, 45)
    kw = dict(eta=0.1) if "Bandit" in cls else {}
    s = getattr(osamp, cls)([96, 48, 24], uniform_fn=philox_uniform_fn(1, 0), **kw)
    inp, out, blocks = s.sample_blocks(g, seeds)
    assert torch.equal(out, seeds) and len(blocks) == 3
    nxt = inp
    for l, b in enumerate(blocks):
        n_dst = b.num_dst_nodes()
        assert torch.equal(b.srcdata["_ID"][:n_dst], b.dstdata["_ID"])             # invariant (i)
        assert torch.equal(b.srcdata["_ID"], nxt if l == 0 else blocks[l - 1].dstdata["_ID"]) or l == 0
        assert b.srcdata["_ID"].unique().numel() == b.num_src_nodes()
        src_g, dst_g = g.coo()
        e = b.edata["_ID"].long()                                                  # block edges are graph edges
        assert torch.equal(src_g[e], b.srcdata["_ID"][b.src]) and torch.equal(dst_g[e], b.dstdata["_ID"][b.dst])
        d = b.in_degrees().to(torch.float32)
        rs = torch.zeros(n_dst).index_add_(0, b.dst, b.edata["edge_weights"])
        if "Bandit" in cls:
            torch.testing.assert_close(rs, d, rtol=1e-5, atol=0)                   # Σ W~ = d_i (:316-320)
            qs = torch.zeros(n_dst).index_add_(0, b.dst, b.edata["q_ij"])
            assert (qs <= 1 + 1e-5).all()
        if "Poisson" in cls and "Bandit" in cls:
            P = b.srcdata["node_prob"]
            assert (P[:n_dst] == 1).all() and (P > 0).all() and (P <= 1).all()     # invariant (iii)
    for l in range(2):
        assert torch.equal(blocks[l].dstdata["_ID"], blocks[l + 1].srcdata["_ID"])


def test_oracle_contract_vs_native_are_close():
    g = random_graph(800, 6000, seed=7, hubs=2, hub_degree=500)
    seeds = torch.arange(0, 40)
    probs = {}
    for accum in ("native", "contract"):
        s = osamp.PoissonBanditLadiesSampler([64], eta=0.1, accum=accum, uniform_fn=philox_uniform_fn(1, 0))
        s.sample_blocks(g, seeds)
        probs[accum] = s.trace["prob"][0][1].double()
    rel = ((probs["native"] - probs["contract"]).abs() / probs["native"]).max().item()
    assert rel < 1e-5, rel      # north-star tolerance; fp32 index_add_ over ~500-edge rows is the looser side


def test_take_all_and_exp3_norm():
    g = random_graph(200, 600, seed=1)
    s = osamp.PoissonBanditLadiesSampler([10 ** 6], eta=0.1, uniform_fn=philox_uniform_fn(0, 0))
    seeds = torch.arange(0, 20)
    _, _, blocks = s.sample_blocks(g, seeds)
    b = blocks[0]
    assert b.num_edges() == int(g.in_degrees(seeds).sum())                          # full in-neighbourhood
    assert (b.srcdata["node_prob"] == 1).all()
    b.srcdata["embed_norm"] = torch.rand(b.num_src_nodes()) + 0.1
    s.exp3(blocks, g)
    assert abs(float(s.exp3_weights[0].double().sum()) - 1.0) < 1e-5               # ‖w‖₁ = 1 after an update


def test_dgl_op_restatements():
    g = toy_graph()
    sg = dglops.in_subgraph(g, torch.tensor([1, 0]))
    assert sg.edata["_ID"].tolist() == [2, 3, 5, 0, 1, 4]                           # seed order, CSC order, self-loop last
    cg = dglops.compact_graphs(sg, torch.tensor([1, 0]))
    assert cg.ndata["_ID"].tolist() == [1, 0, 3, 4, 2]                              # seeds first, then first occurrence
    blk = dglops.to_block(cg, torch.tensor([0, 1]))
    assert blk.srcdata["_ID"].tolist() == [0, 1, 2, 3, 4] and blk.num_dst_nodes() == 2
    e = torch.tensor([1.0, 2.0, 3.0, 1.0, 1.0, 1.0])
    sm = dglops.edge_softmax(blk, e)
    assert abs(float(sm[:3].sum()) - 1) < 1e-6 and abs(float(sm[3:].sum()) - 1) < 1e-6
