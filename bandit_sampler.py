"""Drop-in module name of the reference (``bandit_sampler.py``): same public names, B200-native underneath."""
from bliss_gnn_b200.graph import normalized_edata  # noqa: F401
from bliss_gnn_b200.sampler import BanditLadiesSampler, PoissonBanditLadiesSampler  # noqa: F401
