"""ORACLE (test infrastructure, never shipped): the DGL 2.2.1 graph-op semantics the
reference samplers rely on, restated with plain CPU torch index ops.

PARITY UNPINNED: DGL is an un-vendored pip dependency of the reference (README.md:21,
``dgl==2.2.1``) and is not installable in this sandbox; the reference ships no tests or golden
vectors.  These restatements follow the documented/recalled DGL semantics listed in
SURVEY.md §8(c) and are pinned only by the hand-derived known-answer vector in
``tests/golden/toy_kat.json``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
"""
from __future__ import annotations

import torch

NID = "_ID"
EID = "_ID"

#: fixed-point scale of the Poisson scale search's sum (numeric contract)
S_FIX_BITS = 40


class OGraph:
    """Homogeneous graph in COO (edge order is meaningful) with node / edge frames."""

    def __init__(self, src, dst, num_nodes):
        self.src = src.long()
        self.dst = dst.long()
        self._n = int(num_nodes)
        self.ndata = {}
        self.edata = {}
        self.srcdata = self.ndata
        self.dstdata = self.ndata
        self.idtype = torch.int32
        self.is_block = False

    def num_nodes(self):
        return self._n

    def num_edges(self):
        return int(self.src.numel())

    def edges(self):
        return self.src, self.dst

    def in_degrees(self):
        return torch.bincount(self.dst, minlength=self._n)

    def out_degrees(self):
        return torch.bincount(self.src, minlength=self._n)

    @property
    def device(self):
        return self.src.device


class OBlock:
    """Bipartite block as produced by ``dgl.to_block`` (src ids include the dst ids first)."""

    is_block = True

    def __init__(self, src, dst, n_src, n_dst):
        self.src = src.long()
        self.dst = dst.long()
        self._n_src = int(n_src)
        self._n_dst = int(n_dst)
        self.srcdata = {}
        self.dstdata = {}
        self.edata = {}

    def num_src_nodes(self):
        return self._n_src

    def num_dst_nodes(self):
        return self._n_dst

    number_of_dst_nodes = num_dst_nodes

    def num_edges(self):
        return int(self.src.numel())

    def in_degrees(self):
        return torch.bincount(self.dst, minlength=self._n_dst)

    def out_degrees(self):
        return torch.bincount(self.src, minlength=self._n_src)

    def int(self):
        return self

    def to(self, _device):
        return self

    def canonical(self):
        key = self.dst * max(self._n_src, 1) + self.src
        perm = torch.argsort(key)
        out = {"src": self.src[perm], "dst": self.dst[perm], "perm": perm}
        for k, v in self.edata.items():
            out[k] = v[perm]
        return out


# ---- segment sums ---------------------------------------------------------------------

def seg_sum(x, seg, n, accum="native"):
    """Σ of ``x`` per segment id.  ``native`` adds in the working dtype like DGL's scalar
    g-SpMM; ``contract`` is the B200 path's numeric contract for row sums (accumulate in
    fp64, round once to the working dtype — DESIGN.md §4)."""
    if accum == "contract":
        out = torch.zeros(n, dtype=torch.float64).index_add_(0, seg, x.to(torch.float64))
        return out.to(x.dtype)
    return torch.zeros(n, dtype=x.dtype).index_add_(0, seg, x)


def fx_bits_for(n_seeds: int) -> int:
    """Fixed-point fraction bits of the column accumulator for a layer with ``n_seeds`` rows
    (contract: every term ≤ ~1, at most n_seeds terms per column, sum must stay < 2^63)."""
    return 62 - max(1, int(n_seeds).bit_length())


def seg_sum_fixed(x, seg, n, frac_bits):
    """Order-independent column sum: each term is quantised to ``max(1, rint(x·2^bits))``
    and added as an integer (what the device does with 64-bit integer atomics)."""
    q = torch.round(x.to(torch.float64) * float(2 ** frac_bits)).to(torch.int64).clamp_(min=1)
    acc = torch.zeros(n, dtype=torch.int64).index_add_(0, seg, q)
    return (acc.to(torch.float64) * float(2.0 ** -frac_bits)).to(x.dtype)


# ---- dgl.ops (scalar edge/node broadcasts) -------------------------------------------

def copy_e_sum(g, x, accum="native", frac_bits=None):
    """``dgl.ops.copy_e_sum``: out[v] = Σ_{e: dst(e)=v} x_e."""
    n = g.num_dst_nodes() if g.is_block else g.num_nodes()
    if accum == "fixed":
        return seg_sum_fixed(x, g.dst, n, frac_bits)
    return seg_sum(x, g.dst, n, accum)


def e_div_v(g, e, v):
    return e / v[g.dst]


def e_div_u(g, e, u):
    return e / u[g.src]


def e_mul_v(g, e, v):
    return e * v[g.dst]


def e_dot_v(g, e, v):
    return e * v[g.dst]


def v_add_e(g, v, e):
    return v[g.dst] + e


def u_div_e(g, u, e):
    return u[g.src] / e


def reverse(g):
    """``dgl.reverse(copy_edata=True)``: endpoints swapped, edge order and ids unchanged."""
    r = OGraph(g.dst, g.src, g.num_nodes())
    r.ndata.update(g.ndata)
    r.edata.update(g.edata)
    return r


# ---- structural transforms --------------------------------------------------------------

def in_subgraph(g, seeds):
    """``dgl.in_subgraph``: for each seed in the given order its in-edges in CSC order; all |V|
    nodes kept; ``edata[EID]`` = original edge ids.  ``g`` is a CSC container with
    ``indptr/indices/eid`` (bliss_gnn_b200.graph.Graph or any look-alike)."""
    seeds = seeds.long()
    start = g.indptr[seeds]
    deg = g.indptr[seeds + 1] - start
    total = int(deg.sum())
    row = torch.repeat_interleave(torch.arange(seeds.numel()), deg)
    off = torch.arange(total) - torch.repeat_interleave(torch.cumsum(deg, 0) - deg, deg)
    pos = start[row] + off
    sg = OGraph(g.indices[pos].long(), seeds[row], g.num_nodes())
    sg.edata[EID] = g.eid[pos].long()
    sg.edata["_csc_pos"] = pos
    return sg


def compact_graphs(g, always_preserve):
    """``dgl.compact_graphs``: new ids by first occurrence in concat(preserve, src, dst)."""
    allv = torch.cat([always_preserve.long(), g.src, g.dst])
    uniq, inv = torch.unique(allv, return_inverse=True)
    first = torch.full((uniq.numel(),), allv.numel(), dtype=torch.int64)
    first.scatter_reduce_(0, inv, torch.arange(allv.numel()), reduce="amin")
    order = torch.argsort(first)                 # unique ids ordered by first occurrence
    new_of_uniq = torch.empty_like(order)
    new_of_uniq[order] = torch.arange(order.numel())
    relabel = new_of_uniq[inv]
    k = always_preserve.numel()
    E = g.num_edges()
    out = OGraph(relabel[k:k + E], relabel[k + E:], order.numel())
    out.ndata[NID] = uniq[order]
    out.edata.update(g.edata)
    return out


def node_subgraph(g, nodes):
    """``g.subgraph(nodes)``: new id = position in ``nodes``; edges with both endpoints kept, in
    the parent's edge order (canonical choice); ``edata[EID]`` = parent positions."""
    nodes = nodes.long()
    pos = torch.full((g.num_nodes(),), -1, dtype=torch.int64)
    pos[nodes] = torch.arange(nodes.numel())
    keep = (pos[g.src] >= 0) & (pos[g.dst] >= 0)
    idx = torch.nonzero(keep, as_tuple=True)[0]
    sg = OGraph(pos[g.src[idx]], pos[g.dst[idx]], nodes.numel())
    sg.ndata[NID] = nodes
    sg.edata[EID] = idx
    return sg


def edge_subgraph(g, mask):
    """``dgl.edge_subgraph(relabel_nodes=False)``: order-preserving edge filter."""
    idx = torch.nonzero(mask, as_tuple=True)[0]
    sg = OGraph(g.src[idx], g.dst[idx], g.num_nodes())
    sg.edata[EID] = idx
    return sg


def to_block(g, dst_nodes):
    """``dgl.to_block``: dst ids = position in ``dst_nodes``; src ids = dst nodes first then the
    other sources by first occurrence in the edge list; edge order preserved."""
    dst_nodes = dst_nodes.long()
    allv = torch.cat([dst_nodes, g.src])
    uniq, inv = torch.unique(allv, return_inverse=True)
    first = torch.full((uniq.numel(),), allv.numel(), dtype=torch.int64)
    first.scatter_reduce_(0, inv, torch.arange(allv.numel()), reduce="amin")
    order = torch.argsort(first)
    new_of_uniq = torch.empty_like(order)
    new_of_uniq[order] = torch.arange(order.numel())
    src_local = new_of_uniq[inv][dst_nodes.numel():]
    dpos = torch.full((g.num_nodes(),), -1, dtype=torch.int64)
    dpos[dst_nodes] = torch.arange(dst_nodes.numel())
    blk = OBlock(src_local, dpos[g.dst], order.numel(), dst_nodes.numel())
    blk.srcdata[NID] = uniq[order]
    blk.dstdata[NID] = dst_nodes
    blk.edata[EID] = torch.arange(g.num_edges())
    return blk


def edge_softmax(g, e):
    """``dgl.nn.functional.edge_softmax``: softmax over the in-edges of each dst (any trailing dims)."""
    n = g.num_dst_nodes() if g.is_block else g.num_nodes()
    shp = (n,) + tuple(e.shape[1:])
    idx = g.dst.view(-1, *([1] * (e.dim() - 1))).expand_as(e)
    mx = torch.full(shp, float("-inf"), dtype=e.dtype).scatter_reduce_(0, idx, e, reduce="amax")
    ex = torch.exp(e - mx[g.dst])
    den = torch.zeros(shp, dtype=e.dtype).index_add_(0, g.dst, ex)
    return ex / den[g.dst]
