"""On-disk readers for the datasets the reference loads through DGL / OGB (``load_graph.py:5-80``) — without DGL.

The reference calls ``dgl.data.{Cora,Citeseer,Pubmed}GraphDataset``, ``RedditDataset``, ``FlickrDataset``,
``YelpDataset`` and ``ogb.nodeproppred.DglNodePropPredDataset``; those classes download an archive and parse the raw
files inside.  There is no network here, so :func:`load_real_dataset` reads the SAME raw files from a local directory
(``$BLISS_DATA`` or the ``root`` argument) and returns the plain :class:`~bliss_gnn_b200.graph.Graph` container:

=============  ==============================================================  =================================
dataset        files looked for under ``<root>/<name>/`` (or ``<root>/``)      format (what DGL / OGB ship)
=============  ==============================================================  =================================
any            ``<name>.npz``                                                  generic: ``src, dst`` or ``indptr,
                                                                               indices``; ``features``; ``labels``;
                                                                               ``train_mask, val_mask, test_mask``
cora,          ``ind.<name>.{x,tx,allx,y,ty,ally,graph,test.index}``           Planetoid pickles (DGL citation
citeseer,                                                                      datasets, ``dgl/data/citation_graph.py``)
pubmed
reddit         ``reddit_data.npz`` + ``reddit_graph.npz``                      DGL ``RedditDataset`` raw files
                                                                               (``feature, label, node_types``;
                                                                               scipy COO via ``sp.save_npz``)
flickr, yelp   ``adj_full.npz, feats.npy, class_map.json, role.json``          GraphSAINT layout (DGL ``FlickrDataset``
                                                                               / ``YelpDataset``)
ogbn-*         ``raw/edge.csv[.gz], raw/node-feat.csv[.gz],                   OGB node-property-prediction raw layout
               raw/node-label.csv[.gz], split/*/{train,valid,test}.csv[.gz]``
=============  ==============================================================  =================================

Labels / flags follow ``load_graph.load_dataset``: ``multilabel`` only for yelp (float labels, ``:69-71``); OGB labels
are column 0 as int64 and the class count is the number of distinct non-NaN labels (``:42-46``).  Features are kept in
float32 (the reference casts to bfloat16, ``:7,45``; the north-star parity contract is fp32).  Self-loop handling,
``--undirected`` and the static edge weights stay in ``train.DataModule`` like in the reference
(``train_lightning.py:334-362``).
"""
from __future__ import annotations

import gzip
import json
import os
import pickle
from typing import Optional, Tuple

import numpy as np
import torch

from .graph import Graph

REAL_DATASETS = ("cora", "citeseer", "pubmed", "reddit", "yelp", "flickr", "actor",
                 "ogbn-products", "ogbn-arxiv", "ogbn-papers100M")


class DatasetNotFound(FileNotFoundError):
    pass


def _dirs(root: str, name: str):
    alt = name.replace("-", "_")
    return [os.path.join(root, name), os.path.join(root, alt), root]


def _first(paths):
    for p in paths:
        if os.path.exists(p):
            return p
    return None


def _finish(src, dst, num_nodes, feats, labels, masks, n_classes, multilabel) -> Tuple[Graph, int, bool]:
    g = Graph.from_coo(torch.as_tensor(np.asarray(src), dtype=torch.int64), torch.as_tensor(np.asarray(dst), dtype=torch.int64),
                       int(num_nodes))
    g.ndata["features"] = torch.as_tensor(np.asarray(feats), dtype=torch.float32).contiguous()
    lab = torch.as_tensor(np.asarray(labels))
    g.ndata["labels"] = lab.to(torch.float32) if multilabel else lab.to(torch.int64)     # load_graph.py:69-71
    for key, m in zip(("train_mask", "val_mask", "test_mask"), masks):
        g.ndata[key] = torch.as_tensor(np.asarray(m), dtype=torch.bool)
    g.n_classes, g.multilabel = int(n_classes), bool(multilabel)
    return g, int(n_classes), bool(multilabel)


def _mask(n, idx):
    m = np.zeros(n, dtype=bool)
    m[np.asarray(idx, dtype=np.int64)] = True
    return m


# ---- generic npz --------------------------------------------------------------------------------------
def read_npz(path: str, name: str):
    z = np.load(path, allow_pickle=False)
    if "indptr" in z:
        indptr, indices = z["indptr"].astype(np.int64), z["indices"].astype(np.int64)
        n = indptr.shape[0] - 1
        dst = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
        src = indices
    else:
        src, dst = z["src"].astype(np.int64), z["dst"].astype(np.int64)
        n = int(z["num_nodes"]) if "num_nodes" in z else int(max(src.max(), dst.max())) + 1
    labels = z["labels"]
    multilabel = labels.ndim == 2 and labels.shape[1] > 1
    n_classes = int(z["n_classes"]) if "n_classes" in z else (labels.shape[1] if multilabel else int(labels.max()) + 1)
    return _finish(src, dst, n, z["features"], labels, (z["train_mask"], z["val_mask"], z["test_mask"]), n_classes,
                   multilabel)


# ---- Planetoid (cora / citeseer / pubmed): dgl/data/citation_graph.py ---------------------------------
def read_planetoid(d: str, name: str):
    import scipy.sparse as sp

    def load(suffix):
        with open(os.path.join(d, f"ind.{name}.{suffix}"), "rb") as f:
            return pickle.load(f, encoding="latin1")

    x, y, tx, ty, allx, ally, graph = (load(s) for s in ("x", "y", "tx", "ty", "allx", "ally", "graph"))
    test_idx = np.loadtxt(os.path.join(d, f"ind.{name}.test.index"), dtype=np.int64)
    test_sorted = np.sort(test_idx)
    if name == "citeseer":          # isolated test nodes: pad tx / ty to the full index range (citation_graph.py)
        full = np.arange(test_sorted.min(), test_sorted.max() + 1)
        tx_ext = sp.lil_matrix((len(full), x.shape[1]))
        tx_ext[test_sorted - test_sorted.min(), :] = tx
        ty_ext = np.zeros((len(full), y.shape[1]))
        ty_ext[test_sorted - test_sorted.min(), :] = ty
        tx, ty = tx_ext, ty_ext
    feats = sp.vstack((allx, tx)).tolil()
    feats[test_idx, :] = feats[test_sorted, :]
    onehot = np.vstack((ally, ty))
    onehot[test_idx, :] = onehot[test_sorted, :]
    labels = np.argmax(onehot, axis=1)
    n = feats.shape[0]
    src, dst = [], []
    for u, nbrs in graph.items():                      # adjacency dict -> both directions (networkx from_dict_of_lists)
        for v in nbrs:
            src += [u, v]
            dst += [v, u]
    key = np.unique(np.asarray(src, dtype=np.int64) * n + np.asarray(dst, dtype=np.int64))
    feats = np.asarray(feats.todense(), dtype=np.float32)
    rs = feats.sum(1, keepdims=True)                    # row-normalised features (citation_graph.py _preprocess_features)
    feats = np.divide(feats, rs, out=np.zeros_like(feats), where=rs > 0)
    n_val = min(500, max(0, allx.shape[0] - len(y)))     # citation_graph.py: the 500 nodes after the labelled ones
    masks = (_mask(n, np.arange(len(y))), _mask(n, np.arange(len(y), len(y) + n_val)), _mask(n, test_sorted))
    return _finish(key // n, key % n, n, feats, labels, masks, onehot.shape[1], False)


# ---- Reddit: dgl/data/reddit.py ------------------------------------------------------------------------
def read_reddit(d: str, name: str = "reddit"):
    import scipy.sparse as sp
    data = np.load(os.path.join(d, "reddit_data.npz"))
    adj = sp.load_npz(_first([os.path.join(d, "reddit_graph.npz"), os.path.join(d, "reddit_self_loop_graph.npz")])).tocoo()
    types = data["node_types"]
    labels = data["label"]
    masks = (types == 1, types == 2, types == 3)
    return _finish(adj.row, adj.col, adj.shape[0], data["feature"], labels, masks, int(labels.max()) + 1, False)


# ---- GraphSAINT layout (flickr / yelp): dgl/data/flickr.py, yelp.py -------------------------------------
def read_graphsaint(d: str, name: str):
    import scipy.sparse as sp
    adj = sp.load_npz(os.path.join(d, "adj_full.npz")).tocoo()
    feats = np.load(os.path.join(d, "feats.npy"))
    class_map = json.load(open(os.path.join(d, "class_map.json")))
    role = json.load(open(os.path.join(d, "role.json")))
    n = adj.shape[0]
    first = class_map[str(0)] if "0" in class_map else next(iter(class_map.values()))
    multilabel = isinstance(first, list)
    if multilabel:
        labels = np.zeros((n, len(first)), dtype=np.float32)
        for k, v in class_map.items():
            labels[int(k)] = v
        n_classes = labels.shape[1]
    else:
        labels = np.zeros(n, dtype=np.int64)
        for k, v in class_map.items():
            labels[int(k)] = v
        n_classes = int(labels.max()) + 1
    masks = (_mask(n, role["tr"]), _mask(n, role["va"]), _mask(n, role["te"]))
    return _finish(adj.row, adj.col, n, feats, labels, masks, n_classes, multilabel)


# ---- OGB node property prediction raw layout: ogb/io/read_graph_raw.py ----------------------------------
def _csv(path_no_ext: str, dtype):
    p = _first([path_no_ext + ".csv.gz", path_no_ext + ".csv"])
    if p is None:
        raise DatasetNotFound(path_no_ext + ".csv[.gz]")
    import pandas as pd
    return pd.read_csv(p, header=None, compression="gzip" if p.endswith(".gz") else None).values.astype(dtype)


def read_ogb(d: str, name: str):
    raw = os.path.join(d, "raw")
    edges = _csv(os.path.join(raw, "edge"), np.int64)
    feats = _csv(os.path.join(raw, "node-feat"), np.float32)
    labels = _csv(os.path.join(raw, "node-label"), np.float64)[:, 0]             # load_graph.py:40
    n = feats.shape[0]
    split_root = os.path.join(d, "split")
    split_dir = os.path.join(split_root, sorted(os.listdir(split_root))[0])
    idx = {k: _csv(os.path.join(split_dir, k), np.int64).reshape(-1) for k in ("train", "valid", "test")}
    n_classes = len(np.unique(labels[~np.isnan(labels)]))                        # load_graph.py:43
    masks = (_mask(n, idx["train"]), _mask(n, idx["valid"]), _mask(n, idx["test"]))
    return _finish(edges[:, 0], edges[:, 1], n, feats, np.nan_to_num(labels).astype(np.int64), masks, n_classes, False)


def load_real_dataset(name: str, root: Optional[str] = None) -> Tuple[Graph, int, bool]:
    """``(g, n_classes, multilabel)`` of a real dataset read from local files; raises :class:`DatasetNotFound`
    (with the list of places looked at) when they are not there — it never substitutes a synthetic graph."""
    root = root or os.environ.get("BLISS_DATA")
    looked = []
    if root:
        for d in _dirs(root, name):
            npz = os.path.join(d, f"{name}.npz")
            looked.append(npz)
            if os.path.exists(npz):
                return read_npz(npz, name)
        for d in _dirs(root, name):
            if not os.path.isdir(d):
                continue
            probes = {"planetoid": f"ind.{name}.x", "reddit": "reddit_data.npz", "graphsaint": "adj_full.npz",
                      "ogb": os.path.join("raw", "node-feat.csv.gz"), "ogb2": os.path.join("raw", "node-feat.csv")}
            found = {k for k, f in probes.items() if os.path.exists(os.path.join(d, f))}
            looked += [os.path.join(d, f) for f in probes.values()]
            if "planetoid" in found and name in ("cora", "citeseer", "pubmed"):
                return read_planetoid(d, name)
            if "reddit" in found and name == "reddit":
                return read_reddit(d)
            if "graphsaint" in found and name in ("flickr", "yelp"):
                return read_graphsaint(d, name)
            if found & {"ogb", "ogb2"} and name.startswith("ogbn-"):
                return read_ogb(d, name)
    raise DatasetNotFound(
        f"dataset '{name}': no local files found (BLISS_DATA={root!r}; looked for {looked[:6]} ...). The reference downloads "
        f"it through DGL/OGB (load_graph.py:11-63); this build has no network: put the raw files (or <name>.npz) under "
        f"$BLISS_DATA/{name}/, or ask for a synthetic graph of that shape explicitly with --dataset synthetic:{name}")
