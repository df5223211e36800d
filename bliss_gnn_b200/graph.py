"""Graph and Block containers (no DGL) + synthetic graphs of the reference's dataset shapes.

The reference keeps its graph in a ``dgl.DGLGraph`` restricted to the CSC format
(``train_lightning.py:373``) and hands ``dgl`` *blocks* (MFGs) from the sampler to the
model.  This module holds the two plain-tensor containers that replace them:

* :class:`Graph`  — CSC of the whole (replicated) graph: ``indptr[int64 |V|+1]``,
  ``indices[int32 |E|]`` (source ids, column = destination), ``eid[int32 |E|]``
  (CSC position -> original edge id).  Column order follows DGL's contract
  (SURVEY.md §8c): stable sort of the COO by destination, COO = non-self edges
  followed by the |V| self-loops of ``add_self_loop`` (``train_lightning.py:334-335``)
  so the self-loop is the last entry of every column.
* :class:`Block`  — one sampled bipartite layer in destination-major CSR
  (``indptr[n_dst+1]``, ``edge_src[E_b]`` local source ids) that duck-types the part of
  the DGL block surface the reference touches (``srcdata/dstdata/edata``,
  ``num_src_nodes()``, ``in_degrees()`` ...; ``bandit_sampler.py:322-337``,
  ``model.py:318-329``, ``train_lightning.py:104-139``).

Nothing here is on the GPU hot path; it is containers and one-off graph preparation.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

NID = "_ID"  # dgl.NID
EID = "_ID"  # dgl.EID

#: |V|, |E| (directed, before self-loops), feature width, classes, multilabel,
#: train/val/test split, power-law exponent, max degree cap.  SURVEY.md §8(d).
DATASET_SHAPES = {
    "cora": dict(nodes=2708, edges=10556, feats=1433, classes=7, multilabel=False,
                 split=(140, 500, 1000), gamma=2.5, max_deg=168),
    "citeseer": dict(nodes=3327, edges=9228, feats=3703, classes=6, multilabel=False,
                     split=(120, 500, 1000), gamma=2.5, max_deg=99),
    "pubmed": dict(nodes=19717, edges=88648, feats=500, classes=3, multilabel=False,
                   split=(60, 500, 1000), gamma=2.5, max_deg=171),
    "flickr": dict(nodes=89250, edges=899756, feats=500, classes=7, multilabel=False,
                   split=(0.50, 0.25, 0.25), gamma=2.5, max_deg=5425),
    "reddit": dict(nodes=232965, edges=114615892, feats=602, classes=41, multilabel=False,
                   split=(153431, 23831, 55703), gamma=2.3, max_deg=21657),
    "yelp": dict(nodes=716847, edges=13954819, feats=300, classes=100, multilabel=True,
                 split=(0.75, 0.10, 0.15), gamma=2.3, max_deg=4500),
}


class _Frame(dict):
    """A ``dict`` of per-node / per-edge tensors (``g.ndata`` / ``block.edata``)."""

    def update(self, other=(), **kw):  # keep dict semantics, accept DGL-style dict update
        super().update(other, **kw)


class Graph:
    """Whole-graph CSC container (replaces the ``g.formats(['csc'])`` DGLGraph).

    ``edata`` tensors are stored in **original edge-id order** (the reference's public
    view, ``bandit_sampler.py:127``); :meth:`csc_edata` returns the same data permuted into
    CSC order, which is the layout every kernel reads (coalesced with ``indices``).
    """

    def __init__(self, indptr: torch.Tensor, indices: torch.Tensor, eid: torch.Tensor,
                 num_nodes: int):
        assert indptr.dtype == torch.int64 and indptr.numel() == num_nodes + 1
        assert indices.dtype == torch.int32 and eid.dtype == torch.int32
        self.indptr = indptr.contiguous()
        self.indices = indices.contiguous()
        self.eid = eid.contiguous()
        self._num_nodes = int(num_nodes)
        self.ndata: Dict[str, torch.Tensor] = _Frame()
        self.edata: Dict[str, torch.Tensor] = _Frame()
        self._csc_cache: Dict[str, torch.Tensor] = {}
        self._in_deg: Optional[torch.Tensor] = None
        self.idtype = torch.int32

    # ---- construction -------------------------------------------------------------
    @staticmethod
    def from_coo(src: torch.Tensor, dst: torch.Tensor, num_nodes: int) -> "Graph":
        """Build the CSC by a *stable* sort of the COO by destination (DGL contract)."""
        src = src.to(torch.int64)
        dst = dst.to(torch.int64)
        order = torch.sort(dst, stable=True).indices
        counts = torch.bincount(dst, minlength=num_nodes)
        indptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=src.device)
        torch.cumsum(counts, 0, out=indptr[1:])
        return Graph(indptr, src[order].to(torch.int32), order.to(torch.int32), num_nodes)

    # ---- DGL-like surface ---------------------------------------------------------
    def num_nodes(self) -> int:
        return self._num_nodes

    number_of_nodes = num_nodes

    def num_edges(self) -> int:
        return int(self.indices.numel())

    number_of_edges = num_edges

    @property
    def device(self) -> torch.device:
        return self.indptr.device

    def in_degrees(self, v: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self._in_deg is None:
            self._in_deg = (self.indptr[1:] - self.indptr[:-1])
        return self._in_deg if v is None else self._in_deg[v.long()]

    def csc_edata(self, key: str) -> torch.Tensor:
        """``edata[key]`` permuted into CSC order (cached)."""
        t = self.edata[key]
        hit = self._csc_cache.get(key)
        if hit is None or hit[0] is not t:
            hit = (t, t[self.eid.long()].contiguous())
            self._csc_cache[key] = hit
        return hit[1]

    def to(self, device) -> "Graph":
        device = torch.device(device)
        if device == self.device:
            return self
        g = Graph(self.indptr.to(device), self.indices.to(device), self.eid.to(device),
                  self._num_nodes)
        for k, v in self.ndata.items():
            g.ndata[k] = v.to(device)
        for k, v in self.edata.items():
            g.edata[k] = v.to(device)
        for attr in ("n_classes", "multilabel"):
            if hasattr(self, attr):
                setattr(g, attr, getattr(self, attr))
        return g

    def int(self) -> "Graph":
        return self

    def formats(self, _fmt=None) -> "Graph":
        return self

    def coo(self):
        """(src, dst) in original edge-id order."""
        E = self.num_edges()
        dst_csc = torch.repeat_interleave(
            torch.arange(self._num_nodes, device=self.device), self.in_degrees())
        src = torch.empty(E, dtype=torch.int64, device=self.device)
        dst = torch.empty(E, dtype=torch.int64, device=self.device)
        src[self.eid.long()] = self.indices.long()
        dst[self.eid.long()] = dst_csc
        return src, dst


class Block:
    """One sampled layer (MFG).  Destination-major CSR + the fields the reference reads.

    ``srcdata[NID]`` / ``dstdata[NID]`` are global node ids, ``edata[EID]`` the original
    edge ids (``bandit_sampler.py:331-337``); ``edata['edge_weights'|'q_ij'|'w']`` and
    ``srcdata['node_prob']`` as attached by ``generate_block`` (``:324-328``).  ``features`` /
    ``labels`` are gathered lazily from the parent graph on first access, like DGL's lazy
    frames (``train_lightning.py:138-139``).
    """

    is_block = True

    def __init__(self, indptr, edge_src, edge_dst, src_nid, dst_nid, graph: Optional[Graph] = None,
                 csc_pos: Optional[torch.Tensor] = None):
        self.indptr = indptr          # int32 [n_dst + 1]
        self.edge_src = edge_src      # int32 [E_b], local source id
        self.edge_dst = edge_dst      # int32 [E_b], local destination id
        self.csc_pos = csc_pos        # int64/int32 [E_b], position of the edge in the parent CSC
        self._graph = graph
        self.srcdata = _LazyFrame(self, "src")
        self.dstdata = _LazyFrame(self, "dst")
        self.edata = _LazyFrame(self, "edge")
        self.srcdata[NID] = src_nid
        self.dstdata[NID] = dst_nid
        self._n_src = int(src_nid.numel())
        self._n_dst = int(dst_nid.numel())
        self._n_edges = int(edge_src.numel())
        self._transpose = None        # filled by ops.block_transpose (backward SpMM)

    def num_src_nodes(self) -> int:
        return self._n_src

    number_of_src_nodes = num_src_nodes

    def num_dst_nodes(self) -> int:
        return self._n_dst

    number_of_dst_nodes = num_dst_nodes

    def num_edges(self) -> int:
        return self._n_edges

    number_of_edges = num_edges

    @property
    def device(self):
        return self.indptr.device

    def in_degrees(self) -> torch.Tensor:
        return (self.indptr[1:] - self.indptr[:-1])

    def out_degrees(self) -> torch.Tensor:
        return torch.bincount(self.edge_src.long(), minlength=self._n_src)

    def int(self) -> "Block":
        return self

    def to(self, device) -> "Block":
        if torch.device(device) == self.device:
            return self   # same object, so model side effects reach exp3() (train_lightning.py:102)
        raise RuntimeError("Block.to(other device) is not supported: blocks live where they are sampled")

    def canonical(self):
        """Edges sorted by (dst_local, src_local) with payloads permuted alike — the form
        parity tests compare in (DGL's intra-block edge order is unspecified)."""
        key = self.edge_dst.long() * max(self._n_src, 1) + self.edge_src.long()
        perm = torch.argsort(key)
        out = {"src": self.edge_src[perm], "dst": self.edge_dst[perm], "perm": perm}
        for k in list(self.edata.keys()):
            out[k] = self.edata[k][perm]
        return out


class _LazyFrame(_Frame):
    """Frame that falls back to gathering rows of the parent graph's ``ndata``/``edata``.
    ``on_set[key]`` (optional callables) fire after ``frame[key] = value``: the data-parallel step starts a
    layer's bandit exchange as soon as the model has stored that layer's ``embed_norm``."""

    def __init__(self, block: Block, kind: str):
        super().__init__()
        self._block = block
        self._kind = kind
        self.on_set = {}

    def __setitem__(self, key, value):
        dict.__setitem__(self, key, value)
        hook = self.on_set.get(key) if self.on_set else None
        if hook is not None:
            hook()

    def __missing__(self, key):
        g = self._block._graph
        if g is None:
            raise KeyError(key)
        if self._kind == "edge":
            if key not in g.edata or self._block.csc_pos is None:
                raise KeyError(key)
            val = g.csc_edata(key)[self._block.csc_pos.long()]
        else:
            if key not in g.ndata:
                raise KeyError(key)
            from . import ops  # late import: needs the native library
            val = ops.gather_rows(g.ndata[key], self[NID])
        self[key] = val
        return val

    def __contains__(self, key):
        if dict.__contains__(self, key):
            return True
        g = self._block._graph
        if g is None:
            return False
        return key in (g.edata if self._kind == "edge" else g.ndata)


# ------------------------------------------------------------------------------------
# graph preparation (reference: train_lightning.py:331-373, load_graph.py:91-119)
# ------------------------------------------------------------------------------------

def add_self_loops_and_build(src: torch.Tensor, dst: torch.Tensor, num_nodes: int) -> Graph:
    """``remove_self_loop`` + ``add_self_loop`` (``train_lightning.py:334-335``) then CSC."""
    keep = src != dst
    src, dst = src[keep], dst[keep]
    loops = torch.arange(num_nodes, dtype=src.dtype, device=src.device)
    return Graph.from_coo(torch.cat([src, loops]), torch.cat([dst, loops]), num_nodes)


def normalized_edata(g: Graph, weight: Optional[str] = None) -> torch.Tensor:
    """``w_ij = 1 / in_deg(i)`` per edge, in edge-id order (``bandit_sampler.py:20-27``)."""
    deg = g.in_degrees().to(torch.float32)
    w_csc = torch.repeat_interleave(1.0 / deg, g.in_degrees())
    w = torch.empty_like(w_csc)
    w[g.eid.long()] = w_csc
    if weight is not None:
        w = w * g.edata[weight]
    return w


def toy_graph() -> Graph:
    """The reference's only fixture: ``ToyDataset`` (``load_graph.py:91-119``)."""
    src = torch.tensor([2, 3, 3, 4])
    dst = torch.tensor([0, 0, 1, 1])
    g = add_self_loops_and_build(src, dst, 5)
    g.ndata["features"] = torch.tensor(
        [[0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 1, 0], [0, 0, 0, 1], [1, 0, 0, 0]], dtype=torch.float32)
    g.ndata["labels"] = torch.tensor([0, 0, 1, 1, 1], dtype=torch.int64)
    g.ndata["train_mask"] = torch.ones(5, dtype=torch.bool)
    g.ndata["val_mask"] = torch.zeros(5, dtype=torch.bool)
    g.ndata["test_mask"] = torch.zeros(5, dtype=torch.bool)
    return g


def synthetic_graph(name: str, seed: int = 0, device="cpu", scale: float = 1.0,
                    feat_dtype=torch.float32, with_features: bool = True, planted: bool = False,
                    homophily: float = 0.7, signal: float = 0.35, label_noise: float = 0.1) -> Graph:
    """Chung-Lu power-law graph of a named dataset shape.

    Expected degree ∝ ``(rank+10)^(-1/(γ-1))`` capped at the real maximum degree,
    symmetrised, de-duplicated, self-loops removed then one added per node, node ids
    randomly permuted so degree is not correlated with id.  ``scale`` < 1 shrinks |V| and
    |E| together (tests).  Everything is seeded; generation runs on ``device``.

    ``planted=False``: N(0,1) features and uniform labels — throughput only, nothing to learn.
    ``planted=True``: a planted partition with the same degree sequence, for accuracy runs (the reference measures
    test micro-F1, ``train_lightning.py:686-705``): every node belongs to one of C communities (C = classes;
    16 for multi-label shapes), a fraction ``homophily`` of the edges is re-pointed to an equal-degree-rank node of
    the source's own community, features are ``signal * mu[community] + N(0,1)`` (a weak per-node signal that
    neighbourhood aggregation denoises) and labels are the community (multi-label: the community's random 10 %
    pattern) with ``label_noise`` of them re-drawn at random.
    """
    shape = DATASET_SHAPES[name]
    dev = torch.device(device)
    V = max(8, int(round(shape["nodes"] * scale)))
    E_target = max(8, int(round(shape["edges"] * scale)))
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)

    rank = torch.arange(V, dtype=torch.float64, device=dev)
    wts = (rank + 10.0).pow(-1.0 / (shape["gamma"] - 1.0))
    wts = wts * (E_target / wts.sum())
    cap = float(min(shape["max_deg"], V - 1))
    for _ in range(8):  # cap and redistribute so the expected total stays on target
        wts = wts.clamp(max=cap)
        free = wts < cap
        deficit = E_target - wts.sum()
        if deficit <= 1e-6 * E_target or not bool(free.any()):
            break
        wts = torch.where(free, wts * (1.0 + deficit / wts[free].sum()), wts)
    cdf = torch.cumsum(wts, 0)
    cdf = cdf / cdf[-1]

    n_comm = (16 if shape["multilabel"] else shape["classes"])      # communities of the planted partition, by degree rank
    n_pairs_target = E_target // 2
    keys = torch.empty(0, dtype=torch.int64, device=dev)
    draw = int(n_pairs_target * 1.08) + 16
    for _ in range(6):  # top up until enough distinct undirected pairs survive de-dup
        u = torch.searchsorted(cdf, torch.rand(draw, generator=gen, device=dev, dtype=torch.float64))
        v = torch.searchsorted(cdf, torch.rand(draw, generator=gen, device=dev, dtype=torch.float64))
        u.clamp_(max=V - 1)
        v.clamp_(max=V - 1)
        if planted:     # re-point to the node of u's community with (nearly) the same degree rank
            same = torch.rand(draw, generator=gen, device=dev) < homophily
            v = torch.where(same, ((v // n_comm) * n_comm + (u % n_comm)).clamp_(max=V - 1), v)
        ok = u != v
        lo, hi = torch.minimum(u, v)[ok], torch.maximum(u, v)[ok]
        keys = torch.unique(torch.cat([keys, lo * V + hi]))
        if keys.numel() >= n_pairs_target:
            break
        draw = int((n_pairs_target - keys.numel()) * 1.5) + 16
    if keys.numel() > n_pairs_target:
        sel = torch.randperm(keys.numel(), generator=gen, device=dev)[:n_pairs_target]
        keys = keys[torch.sort(sel).values]
    perm = torch.randperm(V, generator=gen, device=dev)
    a, b = perm[keys // V], perm[keys % V]
    del keys
    src = torch.cat([a, b])
    dst = torch.cat([b, a])
    del a, b
    g = add_self_loops_and_build(src, dst, V)
    del src, dst

    C = shape["classes"]
    comm = torch.empty(V, dtype=torch.int64, device=dev)
    comm[perm] = torch.arange(V, device=dev) % n_comm                # community of every node id (rank -> id by perm)
    if with_features:
        F = shape["feats"]
        feats = torch.randn(V, F, generator=gen, device=dev, dtype=torch.float32)
        if planted:
            mu = torch.randn(n_comm, F, generator=gen, device=dev, dtype=torch.float32)
            feats += signal * mu[comm]
        g.ndata["features"] = feats.to(feat_dtype)
    if shape["multilabel"]:
        if planted:
            pattern = (torch.rand(n_comm, C, generator=gen, device=dev) < 0.1)
            flip = torch.rand(V, C, generator=gen, device=dev) < label_noise * 0.1
            g.ndata["labels"] = (pattern[comm] ^ flip).to(torch.float32)
        else:
            g.ndata["labels"] = (torch.rand(V, C, generator=gen, device=dev) < 0.1).to(torch.float32)
    elif planted:
        noisy = torch.rand(V, generator=gen, device=dev) < label_noise
        g.ndata["labels"] = torch.where(noisy, torch.randint(0, C, (V,), generator=gen, device=dev), comm)
    else:
        g.ndata["labels"] = torch.randint(0, C, (V,), generator=gen, device=dev)
    sp = shape["split"]
    if isinstance(sp[0], float):
        n_tr, n_va = int(V * sp[0]), int(V * sp[1])
        n_te = V - n_tr - n_va
    else:
        f = V / shape["nodes"]
        n_tr, n_va, n_te = (max(1, int(round(s * f))) for s in sp)
        n_te = min(n_te, V - n_tr - n_va)
    order = torch.randperm(V, generator=gen, device=dev)
    for nm, lo, hi in (("train_mask", 0, n_tr), ("val_mask", n_tr, n_tr + n_va),
                       ("test_mask", n_tr + n_va, n_tr + n_va + n_te)):
        m = torch.zeros(V, dtype=torch.bool, device=dev)
        m[order[lo:hi]] = True
        g.ndata[nm] = m
    g.n_classes = C
    g.multilabel = shape["multilabel"]
    return g


def load_dataset(dataset_name: str, device="cpu", seed: int = 0, root: Optional[str] = None):
    """``load_graph.load_dataset`` (``load_graph.py:65-80``): ``(g, n_classes, multilabel)``.

    * ``toy`` — the reference's fixture (``load_graph.py:91-119``).
    * ``cora | citeseer | pubmed | reddit | yelp | flickr | ogbn-*`` — the REAL dataset, read from local raw files
      under ``root`` / ``$BLISS_DATA`` by :mod:`bliss_gnn_b200.datasets` (the reference downloads them through
      DGL / OGB; there is no network here).  Missing files raise: a real name is never silently replaced by random data.
    * ``synthetic:<name>[:scale][:planted]`` — a Chung-Lu graph of that dataset's shape, stated
      explicitly; ``planted`` gives it learnable labels (see :func:`synthetic_graph`).
    """
    name = dataset_name
    if name.startswith("synthetic:"):
        parts = name.split(":")
        name, scale, planted = parts[1], 1.0, False
        for extra in parts[2:]:
            if extra == "planted":
                planted = True
            else:
                scale = float(extra)
        if name not in DATASET_SHAPES:
            raise ValueError(f"unknown dataset shape '{name}'")
        g = synthetic_graph(name, seed=seed, device=device, scale=scale, planted=planted)
        return g, g.n_classes, g.multilabel
    if name == "toy":
        return toy_graph().to(device), 2, False
    from .datasets import REAL_DATASETS, load_real_dataset
    if name not in REAL_DATASETS:
        raise ValueError("unknown dataset")                                      # load_graph.py:77-78
    g, n_classes, multilabel = load_real_dataset(name, root)
    return g.to(device), n_classes, multilabel
