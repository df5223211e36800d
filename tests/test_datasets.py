"""CPU tests of the on-disk dataset readers (``bliss_gnn_b200/datasets.py``; reference ``load_graph.py:5-80``): tiny files
in the layouts DGL / OGB ship are written to a temp directory and read back through ``load_dataset``."""
import gzip
import json
import os
import pickle

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from bliss_gnn_b200.graph import load_dataset


def _toy(n=12, f=5, c=3, seed=0):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n, 40)
    dst = rng.integers(0, n, 40)
    keep = src != dst
    key = np.unique(src[keep] * n + dst[keep])
    src, dst = key // n, key % n
    feats = rng.normal(size=(n, f)).astype(np.float32)
    labels = rng.integers(0, c, n)
    role = rng.permutation(n)
    return n, src, dst, feats, labels, role[:6], role[6:9], role[9:]


def _check(g, n, src, dst, feats, labels, tr, va, te):
    assert g.num_nodes() == n and g.num_edges() == len(src)
    s, d = g.coo()
    assert sorted(zip(s.tolist(), d.tolist())) == sorted(zip(src.tolist(), dst.tolist()))
    assert torch.allclose(g.ndata["features"], torch.from_numpy(feats))
    for mask, idx in (("train_mask", tr), ("val_mask", va), ("test_mask", te)):
        assert sorted(torch.nonzero(g.ndata[mask]).flatten().tolist()) == sorted(np.asarray(idx).tolist())


def test_generic_npz(tmp_path):
    n, src, dst, feats, labels, tr, va, te = _toy()
    m = lambda idx: np.isin(np.arange(n), idx)
    np.savez(tmp_path / "cora.npz", src=src, dst=dst, num_nodes=n, features=feats, labels=labels, train_mask=m(tr),
             val_mask=m(va), test_mask=m(te))
    g, n_classes, multilabel = load_dataset("cora", root=str(tmp_path))
    _check(g, n, src, dst, feats, labels, tr, va, te)
    assert n_classes == int(labels.max()) + 1 and not multilabel and g.ndata["labels"].dtype == torch.int64


def test_reddit_layout(tmp_path):
    n, src, dst, feats, labels, tr, va, te = _toy(seed=1)
    d = tmp_path / "reddit"
    d.mkdir()
    types = np.zeros(n, dtype=np.int64)
    types[tr], types[va], types[te] = 1, 2, 3
    np.savez(d / "reddit_data.npz", feature=feats, label=labels, node_types=types)
    sp.save_npz(d / "reddit_graph.npz", sp.coo_matrix((np.ones(len(src)), (src, dst)), shape=(n, n)))
    g, n_classes, multilabel = load_dataset("reddit", root=str(tmp_path))
    _check(g, n, src, dst, feats, labels, tr, va, te)
    assert not multilabel and torch.equal(g.ndata["labels"], torch.from_numpy(labels))


@pytest.mark.parametrize("name", ["flickr", "yelp"])
def test_graphsaint_layout(tmp_path, name):
    n, src, dst, feats, labels, tr, va, te = _toy(seed=2)
    d = tmp_path / name
    d.mkdir()
    sp.save_npz(d / "adj_full.npz", sp.csr_matrix((np.ones(len(src)), (src, dst)), shape=(n, n)))
    np.save(d / "feats.npy", feats)
    if name == "yelp":       # multi-label: a 0/1 list per node (load_graph.py:69-71 casts to float32)
        lab = (np.random.default_rng(0).random((n, 4)) < 0.4).astype(int)
        class_map = {str(i): lab[i].tolist() for i in range(n)}
    else:
        class_map = {str(i): int(labels[i]) for i in range(n)}
    json.dump(class_map, open(d / "class_map.json", "w"))
    json.dump({"tr": tr.tolist(), "va": va.tolist(), "te": te.tolist()}, open(d / "role.json", "w"))
    g, n_classes, multilabel = load_dataset(name, root=str(tmp_path))
    _check(g, n, src, dst, feats, labels, tr, va, te)
    if name == "yelp":
        assert multilabel and n_classes == 4 and g.ndata["labels"].dtype == torch.float32
        assert torch.equal(g.ndata["labels"], torch.from_numpy(lab).float())
    else:
        assert not multilabel and torch.equal(g.ndata["labels"], torch.from_numpy(labels))


def test_ogb_raw_layout(tmp_path):
    n, src, dst, feats, labels, tr, va, te = _toy(seed=3)
    d = tmp_path / "ogbn_arxiv"
    (d / "raw").mkdir(parents=True)
    (d / "split" / "time").mkdir(parents=True)

    def w(path, arr, fmt):
        with gzip.open(path, "wt") as f:
            np.savetxt(f, arr, delimiter=",", fmt=fmt)

    w(d / "raw" / "edge.csv.gz", np.stack([src, dst], 1), "%d")
    w(d / "raw" / "node-feat.csv.gz", feats, "%.9g")
    w(d / "raw" / "node-label.csv.gz", labels.reshape(-1, 1), "%d")
    for k, idx in (("train", tr), ("valid", va), ("test", te)):
        w(d / "split" / "time" / f"{k}.csv.gz", np.asarray(idx).reshape(-1, 1), "%d")
    g, n_classes, multilabel = load_dataset("ogbn-arxiv", root=str(tmp_path))
    _check(g, n, src, dst, feats, labels, tr, va, te)
    assert n_classes == len(np.unique(labels)) and not multilabel


def test_planetoid_layout(tmp_path):
    """ind.<name>.* pickles as DGL's citation datasets read them: x/y = labelled training nodes, allx/ally = all
    non-test nodes, tx/ty = test nodes in the order of ind.<name>.test.index, graph = adjacency dict."""
    rng = np.random.default_rng(4)
    n, f, c, n_test, n_lab = 20, 6, 3, 5, 4
    feats = (rng.random((n, f)) < 0.4).astype(np.float32)
    feats[:, 0] = 1.0
    labels = rng.integers(0, c, n)
    onehot = np.eye(c)[labels]
    test_idx = rng.permutation(np.arange(n - n_test, n))
    d = tmp_path / "cora"
    d.mkdir()
    graph = {i: [] for i in range(n)}
    for _ in range(30):
        u, v = rng.integers(0, n, 2)
        if u != v and v not in graph[u]:
            graph[u].append(int(v))
    objs = {"x": sp.csr_matrix(feats[:n_lab]), "y": onehot[:n_lab], "allx": sp.csr_matrix(feats[: n - n_test]),
            "ally": onehot[: n - n_test], "tx": sp.csr_matrix(feats[test_idx]), "ty": onehot[test_idx],     # file order
            "graph": graph}
    for k, o in objs.items():
        pickle.dump(o, open(d / f"ind.cora.{k}", "wb"))
    np.savetxt(d / "ind.cora.test.index", test_idx, fmt="%d")
    g, n_classes, multilabel = load_dataset("cora", root=str(tmp_path))
    assert g.num_nodes() == n and n_classes == c and not multilabel
    assert torch.equal(g.ndata["labels"], torch.from_numpy(labels))
    ref = feats / feats.sum(1, keepdims=True)
    assert torch.allclose(g.ndata["features"], torch.from_numpy(ref), atol=1e-6)
    s, dd = g.coo()
    pairs = set(zip(s.tolist(), dd.tolist()))
    for u, nb in graph.items():
        for v in nb:
            assert (u, v) in pairs and (v, u) in pairs                  # undirected, like the DGL datasets
    assert int(g.ndata["train_mask"].sum()) == n_lab and int(g.ndata["test_mask"].sum()) == n_test


def test_missing_real_dataset_raises_and_synthetic_is_explicit(tmp_path):
    with pytest.raises(FileNotFoundError):
        load_dataset("reddit", root=str(tmp_path))
    with pytest.raises(ValueError):
        load_dataset("not-a-dataset")
    g, c, m = load_dataset("synthetic:cora:0.2")
    assert g.num_nodes() == 542 and c == 7 and not m
    gp, _, _ = load_dataset("synthetic:cora:0.5:planted")
    s, d = gp.coo()
    lab = gp.ndata["labels"]
    assert float((lab[s] == lab[d]).float().mean()) > 0.5      # planted partition: edges mostly inside a community
