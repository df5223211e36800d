"""Drop-in module name of the reference (``ladies_sampler.py``)."""
from bliss_gnn_b200.graph import normalized_edata  # noqa: F401
from bliss_gnn_b200.sampler import LadiesSampler, PoissonLadiesSampler  # noqa: F401
