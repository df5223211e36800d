"""ORACLE (test infrastructure, never shipped): CPU restatement of the reference samplers.

Follows ``/root/reference/bandit_sampler.py`` and ``ladies_sampler.py`` method by method
(citations on each function) with the DGL calls replaced by ``oracle.dglops``.

PARITY UNPINNED (see ``oracle/dglops.py``): the reference cannot run here (no DGL) and ships no
golden vectors; the only pin is the hand-derived toy-graph vector ``tests/golden/toy_kat.json``.

Differences from the reference that are deliberate and visible:
* dtype is a parameter (reference: bfloat16 everywhere; north-star parity is fp32).
* random draws are injectable: ``uniform_fn(block_id, global_nids, prob) -> float32 u in [0,1)``.
  Selection rule ``u < P`` equals CPU ``torch.bernoulli``; non-Poisson selection
  is ``topk(p / Exp(1))`` with ``Exp(1) = -log1p(-u)``, which is how ``torch.multinomial`` without
  replacement is defined.
* ``accum='contract'`` switches the segmented sums to the B200 path's order-independent numeric
  contract (fp64 row sums, 64-bit fixed-point column sums and scale-search sums).
  ``accum='native'`` adds in the working dtype the way torch/DGL would.
"""
from __future__ import annotations

import math

import torch

from . import dglops as ops
from .dglops import EID, NID


def find_indices_in(a, b):
    """bandit_sampler.py:5-14."""
    b_sorted, indices = torch.sort(b)
    sorted_indices = torch.searchsorted(b_sorted, a)
    sorted_indices[sorted_indices >= indices.shape[0]] = 0
    return indices[sorted_indices]


def union(*arrays):
    """bandit_sampler.py:16-18."""
    return torch.unique(torch.cat(arrays))


def normalized_edata(g, dtype=torch.float32):
    """bandit_sampler.py:20-27 / ladies_sampler.py:15-22: w_ij = 1 / in_deg(i), edge-id order."""
    deg = (g.indptr[1:] - g.indptr[:-1])
    cdt = torch.float64 if dtype == torch.float64 else torch.float32
    w_csc = torch.repeat_interleave((1.0 / deg.to(cdt)), deg)
    w = torch.empty_like(w_csc)
    w[g.eid.long()] = w_csc
    return w.to(dtype)


class BanditLadiesSampler:
    """bandit_sampler.py:29-367."""

    def __init__(self, nodes_per_layer, importance_sampling=True, weight="w", out_weight="edge_weights",
                 node_embedding="nfeat", node_prob="node_prob", replace=False, eta=0.4, num_steps=5000,
                 model="sage", dtype=torch.float32, accum="native", uniform_fn=None):
        self.nodes_per_layer = nodes_per_layer
        self.importance_sampling = importance_sampling
        self.edge_weight = weight
        self.output_weight = out_weight
        self.node_prob = node_prob
        self.node_embedding = node_embedding
        self.replace = replace
        self.eta = eta
        self.T = num_steps
        self.exp3_weights = None
        self.model = model
        self.dtype = dtype
        self.accum = accum
        self.uniform_fn = uniform_fn
        self._layer = None
        self.trace = {}

    # -- accumulation helpers (native vs contract) --
    def _rowsum(self, g, x):
        return ops.copy_e_sum(g, x, accum="contract" if self.accum == "contract" else "native")

    def _colsum(self, g, x, n_seeds):
        if self.accum == "contract":
            return ops.copy_e_sum(g, x, accum="fixed", frac_bits=ops.fx_bits_for(n_seeds))
        return ops.copy_e_sum(g, x)

    def compute_prob(self, insg, seed_nodes, edge_prob, num):
        """bandit_sampler.py:47-82."""
        if self.importance_sampling:
            edge_prob_sum = self._rowsum(insg, edge_prob)                       # :67
            out_frontier = ops.reverse(insg)                                    # :69
            edge_prob_div_sum = ops.e_div_u(out_frontier, edge_prob, edge_prob_sum)   # :71
            prob = self._colsum(out_frontier, edge_prob_div_sum ** 2, seed_nodes.numel())  # :73
            prob = torch.sqrt(prob)                                             # :75
        else:
            prob = torch.ones(insg.num_nodes()).to(self.dtype)                  # :79
            prob[insg.out_degrees() == 0] = 0                                   # :81
        return prob

    def _uniform(self, insg, prob=None):
        """uniform_fn(block_id, global_nids, prob) -> float32 draws, one per candidate."""
        return torch.as_tensor(self.uniform_fn(self._layer, insg.ndata[NID], prob), dtype=torch.float32)

    def select_neighbors(self, prob, num, insg=None):
        """bandit_sampler.py:84-99: ``torch.multinomial(prob, min(num, N), replacement=False)``
        == the ``min(num, N)`` largest ``prob / Exp(1)``."""
        u = self._uniform(insg, prob)
        expo = -torch.log1p(-u)
        key = prob.to(torch.float32) / expo
        k = min(num, prob.shape[0])
        return torch.topk(key, k).indices

    def exp3_probabilities(self, idx, g, seed_nodes):
        """bandit_sampler.py:101-138."""
        insg = ops.in_subgraph(g, seed_nodes)                                   # :123
        insg = ops.compact_graphs(insg, seed_nodes)                             # :125
        exp_weights = self.exp3_weights[idx][insg.edata[EID].long()]            # :127
        exp3_weights_sum = self._rowsum(insg, exp_weights)                      # :129
        exp_weights_divided = ops.e_div_v(insg, exp_weights, exp3_weights_sum)  # :131
        n_i = (g.indptr[1:] - g.indptr[:-1])[insg.srcdata[NID]]                 # :133
        cdt = torch.float64 if self.dtype == torch.float64 else torch.float32   # reference: fp32 then cast
        edge_prob = ops.v_add_e(insg, (self.eta / n_i.to(cdt)).to(self.dtype),
                                (1 - self.eta) * exp_weights_divided)          # :137
        return edge_prob, insg

    def calculate_alpha(self, mfg):
        """bandit_sampler.py:140-158."""
        if self.model == "gat":
            q_ij = mfg.edata["q_ij"]
            attention = mfg.edata["a_ij"]
            q_ij_sum = self._rowsum(mfg, q_ij)                                  # :150
            attention_sum = self._rowsum(mfg, attention)                        # :151
            a_div = ops.e_div_v(mfg, attention, attention_sum)                  # :152
            a_div = torch.nan_to_num(a_div)                                     # :153
            alpha = ops.e_dot_v(mfg, a_div, q_ij_sum)                           # :154
        else:
            alpha = mfg.edata[self.edge_weight]                                 # :157
        return alpha

    def calculate_rewards(self, idx, mfg, g, alpha):
        """bandit_sampler.py:160-193."""
        k_i = mfg.in_degrees()[:len(mfg.dstdata[NID])].to(self.dtype)           # :180
        h_j_norm = mfg.srcdata["embed_norm"].detach()                           # :182
        q_ij = mfg.edata["q_ij"]                                                # :184
        alpha_div_k_i = ops.e_div_v(mfg, alpha ** 2, k_i)                       # :186
        alpha_div_k_i = torch.nan_to_num(alpha_div_k_i, posinf=0)               # :187
        h_j_norm_div_q_j = ops.u_div_e(mfg, h_j_norm ** 2, q_ij ** 2)           # :189
        mfg.edata["rewards"] = alpha_div_k_i * h_j_norm_div_q_j                 # :191-193

    def update_exp3_weights(self, idx, mfg, g):
        """bandit_sampler.py:195-249."""
        n_i = (g.indptr[1:] - g.indptr[:-1])[mfg.dstdata[NID].long()].to(self.dtype)   # :223
        delta = 0.01                                                            # :233
        rewards = mfg.edata["rewards"].clone().detach()                         # :236
        prob = mfg.srcdata[self.node_prob].clone().detach()                     # :238
        rewards_hat = ops.e_div_u(mfg, rewards, prob)                           # :240
        delta_reward = ops.e_mul_v(mfg, rewards_hat, delta / n_i)               # :242
        delta_reward[delta_reward > 1] = 1                                      # :244
        exp_rewards = torch.exp(delta_reward)                                   # :246
        self.trace.setdefault("delta_reward", {})[idx] = delta_reward
        w = self.exp3_weights[idx]
        w[mfg.edata[EID].long()] *= exp_rewards                                 # :248
        if self.accum == "contract":
            norm = w.to(torch.float64).abs().sum().clamp_min(1e-12)
            self.exp3_weights[idx] = (w / norm.to(self.dtype))
        else:
            self.exp3_weights[idx] = torch.nn.functional.normalize(w, p=1, dim=0)   # :249

    def exp3(self, mfgs, g):
        """bandit_sampler.py:251-267."""
        for idx, mfg in enumerate(mfgs):
            alpha = self.calculate_alpha(mfg)
            self.calculate_rewards(idx, mfg, g, alpha)
            self.update_exp3_weights(idx, mfg, g)

    # ladies overrides the final normalisation (ladies_sampler.py:94-97)
    def _normalise_block_weights(self, sg, W_tilde):
        W_tilde_sum = self._rowsum(sg, W_tilde)                                 # :316
        d = sg.in_degrees()                                                     # :318
        return ops.e_mul_v(sg, W_tilde, d / W_tilde_sum)                        # :320

    def generate_block(self, insg, neighbor_nodes_idx, seed_nodes, P_sg, W_sg, g=None):
        """bandit_sampler.py:269-339 (ladies_sampler.py:71-107)."""
        seed_nodes_idx = find_indices_in(seed_nodes.long(), insg.ndata[NID])    # :285
        u_nodes = union(neighbor_nodes_idx.long(), seed_nodes_idx)              # :287
        sg = ops.node_subgraph(insg, u_nodes)                                   # :289
        u, v = sg.edges()                                                       # :291
        lu = sg.ndata[NID][u.long()]                                            # :293
        nb = neighbor_nodes_idx.long()
        s = find_indices_in(lu, nb)                                             # :295
        eg = ops.edge_subgraph(sg, lu == nb[s])                                 # :298
        eg.ndata[NID] = sg.ndata[NID][:eg.num_nodes()]                          # :300
        eg.edata[EID] = sg.edata[EID][eg.edata[EID].long()]                     # :302
        sg = eg                                                                 # :304
        nids = insg.ndata[NID][sg.ndata[NID].long()]                            # :306
        P = P_sg[u_nodes.long()]                                                # :309
        W = W_sg[sg.edata[EID].long()]                                          # :311
        W_tilde = ops.e_div_u(sg, W, P)                                         # :314
        W_tilde = self._normalise_block_weights(sg, W_tilde)                    # :316-320
        block = ops.to_block(sg, seed_nodes_idx)                                # :322
        block.edata[self.output_weight] = W_tilde                               # :324
        self._attach(block, W, P)                                               # :326-328
        block.srcdata[NID] = nids[block.srcdata[NID].long()]                    # :331
        block.dstdata[NID] = nids[block.dstdata[NID].long()]                    # :333
        ins_pos = sg.edata[EID].long()
        sg_eids = insg.edata[EID][ins_pos]                                      # :335
        block.edata[EID] = sg_eids[block.edata[EID].long()]                     # :337
        if g is not None:   # frames DGL carries through every sub-graph op (calculate_alpha :157)
            for name, val in g.edata.items():
                block.edata[name] = val[block.edata[EID].long()]
        return block

    def _attach(self, block, W, P):
        block.edata["q_ij"] = W                                                 # :326
        block.srcdata[self.node_prob] = P                                       # :328

    def sample_blocks(self, g, seed_nodes, exclude_eids=None):
        """bandit_sampler.py:341-367."""
        if self.exp3_weights is None:
            self.exp3_weights = torch.ones(len(self.nodes_per_layer), g.num_edges()).to(self.dtype)  # :343
        seed_nodes = seed_nodes.long()
        output_nodes = seed_nodes
        blocks = []
        for block_id in reversed(range(len(self.nodes_per_layer))):             # :350
            self._layer = block_id
            num = self.nodes_per_layer[block_id]
            edge_prob, insg = self.exp3_probabilities(block_id, g, seed_nodes)  # :354
            node_prob = self.compute_prob(insg, seed_nodes, edge_prob, num)     # :356
            W = edge_prob                                                       # :358
            chosen = self.select_neighbors(node_prob, num, insg)                # :360
            self.trace.setdefault("prob", {})[block_id] = (insg.ndata[NID].clone(), node_prob.clone())
            block = self.generate_block(insg, chosen, seed_nodes, node_prob, W, g)   # :362
            seed_nodes = block.srcdata[NID]                                     # :364
            blocks.insert(0, block)                                             # :366
        return seed_nodes, output_nodes, blocks


def poisson_scale(prob, num, eps, accum):
    """The scale search of bandit_sampler.py:395-401 (ladies_sampler.py:154-160).
    Returns (c, iterations)."""
    one = torch.ones_like(prob)
    c = 1.0
    it = 0
    for i in range(50):
        it = i + 1
        t = torch.minimum(prob * c, one)
        if accum == "contract":
            fx = torch.round(t.to(torch.float64) * float(2 ** ops.S_FIX_BITS)).to(torch.int64)
            S = float(int(fx.sum())) * float(2.0 ** -ops.S_FIX_BITS)
        else:
            S = torch.sum(t.to(torch.float64)).item()
        if min(S, num) / max(S, num) >= eps:
            break
        c *= num / S
    return c, it


class PoissonBanditLadiesSampler(BanditLadiesSampler):
    """bandit_sampler.py:369-424."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.eps = 0.9999                                                       # :379

    def compute_prob(self, insg, seed_nodes, edge_prob, num):
        """bandit_sampler.py:381-406."""
        prob = super().compute_prob(insg, seed_nodes, edge_prob, num)
        one = torch.ones_like(prob)
        if prob.shape[0] <= num:                                                # :392
            return one
        c, it = poisson_scale(prob, num, self.eps, self.accum)                  # :395-401
        self.trace.setdefault("c", {})[self._layer] = (c, it)
        skip_nodes = find_indices_in(seed_nodes.long(), insg.ndata[NID])        # :403
        prob[skip_nodes] = float("inf")                                         # :404
        return torch.minimum(prob * c, one)                                     # :406

    def select_neighbors(self, prob, num, insg=None):
        """bandit_sampler.py:408-425: ``bernoulli(prob) == 1`` ⇔ ``u < prob``."""
        u = self._uniform(insg, prob)
        return torch.arange(prob.shape[0])[u < prob.to(torch.float32)]


class LadiesSampler(BanditLadiesSampler):
    """ladies_sampler.py:24-123 — static edge weights, no bandit state, no q_ij/node_prob."""

    def __init__(self, nodes_per_layer, importance_sampling=True, weight="w", out_weight="edge_weights",
                 replace=False, allow_zero_in_degree=False, dtype=torch.float32, accum="native",
                 uniform_fn=None):
        super().__init__(nodes_per_layer, importance_sampling, weight, out_weight, replace=replace,
                         dtype=dtype, accum=accum, uniform_fn=uniform_fn)
        self.allow_zero_in_degree = allow_zero_in_degree

    def ladies_compute_prob(self, g, seed_nodes, weight, num):
        """ladies_sampler.py:34-52."""
        insg = ops.in_subgraph(g, seed_nodes)                                   # :42
        insg = ops.compact_graphs(insg, seed_nodes)                             # :43
        if self.importance_sampling:
            out_frontier = ops.reverse(insg)                                    # :45
            weight = weight[out_frontier.edata[EID].long()]                     # :46
            prob = self._colsum(out_frontier, weight ** 2, seed_nodes.numel())  # :47
            prob = torch.sqrt(prob)                                             # :48
        else:
            prob = torch.ones(insg.num_nodes())
            prob[insg.out_degrees() == 0] = 0
        return prob, insg

    def _normalise_block_weights(self, sg, W_tilde):
        d = sg.in_degrees()
        return ops.e_mul_v(sg, W_tilde, (d / 1.0).to(self.dtype))               # ladies :97

    def _attach(self, block, W, P):
        pass                                                                    # ladies :99-106

    def _finish_prob(self, prob, insg, seed_nodes, num):
        return prob

    def sample_blocks(self, g, seed_nodes, exclude_eids=None):
        """ladies_sampler.py:109-123."""
        seed_nodes = seed_nodes.long()
        output_nodes = seed_nodes
        blocks = []
        for block_id in reversed(range(len(self.nodes_per_layer))):
            self._layer = block_id
            num = self.nodes_per_layer[block_id]
            W = g.edata[self.edge_weight].to(self.dtype)                        # :114
            prob, insg = self.ladies_compute_prob(g, seed_nodes, W, num)        # :115
            prob = self._finish_prob(prob, insg, seed_nodes, num)
            self.trace.setdefault("prob", {})[block_id] = (insg.ndata[NID].clone(), prob.clone())
            chosen = self.select_neighbors(prob, num, insg)                     # :117
            block = self.generate_block(insg, chosen, seed_nodes, prob,
                                        W[insg.edata[EID].long()], g)           # :118-120
            seed_nodes = block.srcdata[NID]
            blocks.insert(0, block)
        return seed_nodes, output_nodes, blocks

    def exp3(self, mfgs, g):
        raise AttributeError("LadiesSampler has no bandit state")


class PoissonLadiesSampler(LadiesSampler):
    """ladies_sampler.py:125-183."""

    def __init__(self, nodes_per_layer, importance_sampling=True, weight="w", out_weight="edge_weights",
                 allow_zero_in_degree=False, dtype=torch.float32, accum="native", uniform_fn=None):
        # the reference passes allow_zero_in_degree into the ``replace`` slot (:134-136); harmless here
        super().__init__(nodes_per_layer, importance_sampling, weight, out_weight,
                         replace=allow_zero_in_degree, dtype=dtype, accum=accum, uniform_fn=uniform_fn)
        self.eps = 0.9999

    def _finish_prob(self, prob, insg, seed_nodes, num):
        """ladies_sampler.py:150-164."""
        one = torch.ones_like(prob)
        if prob.shape[0] <= num:
            return one
        c, it = poisson_scale(prob, num, self.eps, self.accum)
        self.trace.setdefault("c", {})[self._layer] = (c, it)
        skip_nodes = find_indices_in(seed_nodes.long(), insg.ndata[NID])
        prob[skip_nodes] = float("inf")
        return torch.minimum(prob * c, one)

    def select_neighbors(self, prob, num, insg=None):
        u = self._uniform(insg, prob)
        return torch.arange(prob.shape[0])[u < prob.to(torch.float32)]


class NeighborSampler:
    """``dgl.dataloading.NeighborSampler(fanouts)`` / ``MultiLayerFullNeighborSampler`` (``train_lightning.py:349-357``)
    restated (DGL, recalled): ``for fanout in reversed(fanouts): frontier = g.sample_neighbors(seeds, fanout);
    block = to_block(frontier, seeds); seeds = block.srcdata[NID]``.  ``sample_neighbors`` keeps, per seed,
    ``min(fanout, in-degree)`` in-edges uniformly without replacement (``-1``: all).  DGL draws from its own generator;
    here the draw is injectable: ``key_fn(block_id, csc_positions) -> uint32 keys`` and a seed keeps its ``fanout``
    smallest keys (ties in CSC order) — a uniform subset for i.i.d. keys."""

    def __init__(self, fanouts, key_fn=None):
        self.fanouts, self.key_fn = [int(f) for f in fanouts], key_fn

    def sample_blocks(self, g, seed_nodes, exclude_eids=None):
        seed_nodes = seed_nodes.long()
        output_nodes = seed_nodes
        blocks = []
        for block_id in reversed(range(len(self.fanouts))):
            fanout = self.fanouts[block_id]
            insg = ops.in_subgraph(g, seed_nodes)                  # all in-edges, seed order, CSC order within a seed
            pos = insg.edata["_csc_pos"]
            keep = torch.ones(insg.num_edges(), dtype=torch.bool)
            if fanout > 0:
                keys = torch.as_tensor(self.key_fn(block_id, pos)).long()
                start = g.indptr[seed_nodes]
                deg = g.indptr[seed_nodes + 1] - start
                off = 0
                for d in deg.tolist():
                    if d > fanout:
                        k = keys[off:off + d]
                        order = torch.sort(k, stable=True).indices       # ties: CSC order
                        m = torch.zeros(d, dtype=torch.bool)
                        m[order[:fanout]] = True
                        keep[off:off + d] = m
                    off += d
            frontier = ops.edge_subgraph(insg, keep)
            frontier.edata[EID] = insg.edata[EID][frontier.edata[EID].long()]
            block = ops.to_block(frontier, seed_nodes)
            block.edata[EID] = frontier.edata[EID][block.edata[EID].long()]
            seed_nodes = block.srcdata[NID]
            blocks.insert(0, block)
        return seed_nodes, output_nodes, blocks
