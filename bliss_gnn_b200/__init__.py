"""bliss_gnn_b200 — B200-native BLISS sample-and-aggregate hot path."""
from .graph import Graph, Block, NID, EID, normalized_edata, toy_graph, synthetic_graph, load_dataset  # noqa: F401

__all__ = ["Graph", "Block", "NID", "EID", "normalized_edata", "toy_graph", "synthetic_graph", "load_dataset"]
