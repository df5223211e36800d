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
