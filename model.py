"""Drop-in module name of the reference (``model.py``): SAGE / GCN / GATv2 + custom_GATv2Conv."""
from bliss_gnn_b200.model import GCN, SAGE, GATv2, GraphConv, SAGEConv, custom_GATv2Conv  # noqa: F401
