"""Drop-in module name of the reference (``load_graph.py``): ``load_dataset(name) -> (g, n_classes, multilabel)``
over the synthetic dataset shapes (``toy`` is the reference's own fixture)."""
from bliss_gnn_b200.graph import load_dataset, toy_graph  # noqa: F401
