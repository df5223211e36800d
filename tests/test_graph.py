"""CPU tests of the graph containers and the synthetic dataset shapes (no DGL)."""
import torch

from bliss_gnn_b200.graph import (DATASET_SHAPES, Graph, add_self_loops_and_build, load_dataset, normalized_edata,
                                  synthetic_graph, toy_graph)


def test_toy_graph_matches_reference_fixture():
    g = toy_graph()                                         # load_graph.py:96 + self-loops
    assert g.num_nodes() == 5 and g.num_edges() == 9
    assert g.indptr.tolist() == [0, 3, 6, 7, 8, 9]
    assert g.indices.tolist() == [2, 3, 0, 3, 4, 1, 2, 3, 4]   # self-loop last in every column
    assert g.eid.tolist() == [0, 1, 4, 2, 3, 5, 6, 7, 8]
    w = normalized_edata(g)
    assert torch.allclose(w, torch.tensor([1 / 3] * 6 + [1.0] * 3))
    src, dst = g.coo()
    assert src.tolist() == [2, 3, 3, 4, 0, 1, 2, 3, 4] and dst.tolist() == [0, 0, 1, 1, 0, 1, 2, 3, 4]


def test_csc_edata_permutation_roundtrip():
    g = toy_graph()
    g.edata["x"] = torch.arange(9, dtype=torch.float32)
    assert g.csc_edata("x").tolist() == [float(e) for e in g.eid.tolist()]


def test_self_loops_are_normalised():
    src = torch.tensor([0, 0, 1, 2, 2])
    dst = torch.tensor([0, 1, 2, 2, 0])                     # two existing self-loops are removed, one per node added
    g = add_self_loops_and_build(src, dst, 3)
    assert g.num_edges() == 3 + 3
    for v in range(3):
        col = g.indices[g.indptr[v]:g.indptr[v + 1]].tolist()
        assert col[-1] == v and col.count(v) == 1


def test_synthetic_shapes_small_scale():
    for name in ("cora", "pubmed"):
        g = synthetic_graph(name, seed=0)
        sh = DATASET_SHAPES[name]
        assert g.num_nodes() == sh["nodes"]
        assert abs(g.num_edges() - (sh["edges"] + sh["nodes"])) <= 0.02 * sh["edges"] + 2
        assert g.ndata["features"].shape == (sh["nodes"], sh["feats"])
        assert int(g.ndata["train_mask"].sum()) == sh["split"][0]
        src, dst = g.coo()
        key = src * g.num_nodes() + dst
        assert key.unique().numel() == key.numel()          # simple graph
        rev = dst * g.num_nodes() + src
        assert torch.equal(key.sort().values, rev.sort().values)   # symmetric
        assert int(g.in_degrees().max()) <= 1.5 * sh["max_deg"]      # cap is on the expected degree
    g2 = synthetic_graph("cora", seed=0)
    assert torch.equal(g2.indices, synthetic_graph("cora", seed=0).indices)   # seeded


def test_load_dataset_signature():
    g, n_classes, multilabel = load_dataset("toy")
    assert n_classes == 2 and multilabel is False and g.num_nodes() == 5
    g, n_classes, multilabel = load_dataset("synthetic:yelp:0.002")
    assert multilabel and g.ndata["labels"].dtype == torch.float32 and g.ndata["labels"].shape[1] == n_classes
